#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the CPU ORACLE (oracle/, the restatement of the reference -- the Rust reference itself
cannot be built in this image).  Fixtures pin the oracle against drift and give the GPU tests
committed vectors to compare with.  Run from the repo root:  python tests/golden/make_golden.py

Per scene: fixed ray batches (B1 primary rays, B2 seeded random rays, B3 first-bounce shadow / continuation rays)
with the oracle's closest-hit records and any-hit flags, per-sample radiance (S2) on a pixel grid, and a 2-spp film (S1)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import craytracer_b200 as c  # noqa: E402
import oracle_lib as o  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402

SEED = 7
SPP = 2


def scene_table():
    """name -> (host scene factory, random-ray box lo, hi).  Kept in one place: tests/test_golden.py imports it."""
    def dragon_small():
        c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 30001, 0)
        return c.parse_scene(scenes.dragon(width=96, height=64), base_dir="/nonexistent")
    def staircase_small():
        # scenes/staircase.cry over the stand-in assets (assets/objs/staircase: 26 MTL materials, 10 image textures, thin lens)
        c.register_standin_mesh("objs/staircase/staircase.obj", 1, 20000, 0)
        return c.parse_scene(scenes.staircase(width=48, height=80), base_dir=os.path.join(ROOT, "assets"))
    return {
        "simple": (lambda: c.parse_scene(scenes.simple(width=96, height=56)), [-45, -1, -35], [45, 12, 55]),
        "materials": (lambda: c.parse_scene(scenes.materials(width=96, height=64)), [-8, -1, -6], [12, 16, 16]),
        "test": (lambda: c.parse_scene(scenes.test_scene(width=64, height=64)), [-1, -1, -3], [4, 3, 1]),
        "rounding-error": (lambda: c.parse_scene(scenes.rounding_error(width=64, height=64)), [-10, -1, -10], [10, 8, 10]),
        "dragon_small": (dragon_small, [-120, -45, -60], [120, 60, 60]),
        "staircase_small": (staircase_small, [-1.6, -0.1, -2.1], [1.6, 5.6, 3.1]),
        # scenes/cornell.cry over the authored stand-in mesh (assets/objs/local/cornell)
        "cornell": (lambda: c.parse_scene(scenes.cornell(width=64, height=64), base_dir=os.path.join(ROOT, "assets")), [-1.1, -0.1, -1.1], [1.1, 2.1, 1.1]),
    }


def batches(orc, width, height, lo, hi):
    ys, xs = np.mgrid[0:height:4, 0:width:4]
    xs = xs.ravel().astype(np.uint32)
    ys = ys.ravel().astype(np.uint32)
    ss = np.zeros_like(xs)
    rng = np.random.default_rng(1)
    org = rng.uniform(lo, hi, size=(600, 3))
    d = rng.normal(size=(600, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    b3s, b3c = orc.bounce_rays(xs, ys, ss)
    rays = np.concatenate([orc.camera_rays(xs, ys, ss), c.make_rays(org, d), b3s, b3c])
    return xs, ys, rays


def main():
    for name, (factory, lo, hi) in scene_table().items():
        hs = factory()
        orc = o.OracleScene(hs)
        w, h = hs.desc.camera.width, hs.desc.camera.height
        xs, ys, rays = batches(orc, w, h, lo, hi)
        hits = orc.intersect(rays)
        occluded = orc.intersects(rays)
        li, ok = orc.estimate_Li(xs, ys, np.full_like(xs, 3), seed=SEED)
        film, counts = orc.render(w, h, seed=0, sample_begin=0, sample_end=SPP)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, rays=rays.view(np.float64).reshape(-1, 7), prim=hits["prim"], t=hits["t"], u=hits["u"], v=hits["v"],
                            occluded=occluded, xs=xs, ys=ys, li=li, li_ok=ok, film=film.astype(np.float32), counts=np.asarray(counts, dtype=np.uint64))
        print(f"{name}: {len(rays)} rays, {int((hits['prim'] != c.CRAY_NO_HIT).sum())} hits, {int(occluded.sum())} occluded, "
              f"{len(xs)} radiance samples, film {w}x{h} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
