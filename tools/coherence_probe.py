#!/usr/bin/env python3
"""Upper bound on what ray reordering could buy the traversal kernels (GPU box): first-bounce continuation / shadow rays of the
bench workload traced in pixel-tile order (origins spatially coherent) against the same rays in random order."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import craytracer_b200 as c  # noqa: E402
import oracle_lib as o  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402


def main():
    scenes.register_standins()
    hs = c.parse_scene(scenes.dragon(), base_dir=os.path.join(ROOT, "assets"))
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    w, h = gpu.width, gpu.height
    # pixels in 8x4 tiles, 4 samples each
    order = [(x, y) for ty in range(0, h, 4) for tx in range(0, w, 8) for y in range(ty, min(ty + 4, h)) for x in range(tx, min(tx + 8, w))]
    xs = np.array([p[0] for p in order] * 4, dtype=np.uint32)
    ys = np.array([p[1] for p in order] * 4, dtype=np.uint32)
    ss = np.repeat(np.arange(4, dtype=np.uint32), len(order))
    t0 = time.time()
    shadow, cont = orc.bounce_rays(xs, ys, ss)
    print(f"oracle produced {len(cont)} continuation and {len(shadow)} shadow rays in {time.time() - t0:.1f} s")
    rng = np.random.default_rng(0)
    primary = orc.camera_rays(xs, ys, ss)
    mixed = np.concatenate([primary, cont])
    interleaved = np.empty_like(mixed)  # what the wavefront queue looks like: fresh camera rays scattered among continuing paths
    perm = rng.permutation(len(mixed))
    interleaved[:] = mixed[perm]
    for name, rays, closest in (("primary", primary, True), ("primary+cont", mixed, True), ("continuation", cont, True), ("shadow", shadow, False)):
        for label, arr in (("tile order", rays), ("shuffled", rays[rng.permutation(len(rays))])):
            d_rays = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).cuda()
            d_out = torch.empty(len(arr) * 32, dtype=torch.uint8, device="cuda")
            stream = torch.cuda.current_stream()
            best = 1e9
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(stream)
                if closest:
                    gpu.intersect_device(d_rays.data_ptr(), len(arr), d_out.data_ptr(), mode=c.TRAVERSE_FAST, stream=stream.cuda_stream)
                else:
                    gpu.intersects_device(d_rays.data_ptr(), len(arr), d_out.data_ptr(), mode=c.TRAVERSE_FAST, stream=stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{name:13s} {label:10s} {len(arr)} rays  {best:.3f} ms  {len(arr) / best / 1e3:.0f} Mrays/s")


if __name__ == "__main__":
    main()
