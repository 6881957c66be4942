"""Command line of the reference (src/bin/craytracer.rs:321-374) over the B200 path:

    python -m craytracer_b200 --scene scenes/dragon.cry [--output out.exr] [--seed 0] [--spp N] [--mode fast|exact|f32] [--gpus N]

Same flags as the reference's `Cli` (--scene/-s, --output, --seed, --preview); --preview is accepted and ignored (no window:
SURVEY section 2 marks the minifb preview out of scope).  Meshes named by a scene are resolved against the scene file's
directory's parent, the current directory and this repository's assets/; meshes that exist nowhere (the reference does not
ship xyzrgb_dragon.obj / staircase.obj) fall back to the documented procedural stand-ins with a warning.
With --gpus N > 1 the scene is replicated on N GPUs of this process, each renders a slice of the sample range, and the films
are combined by one NCCL reduce (cray_render_multi); bench.py does the same with one process per GPU."""
import argparse
import os
import sys
import time

import numpy as np

from . import BUILD_EXACT, BUILD_F32, BUILD_FAST, Scene, TRAVERSE_EXACT, TRAVERSE_F32, TRAVERSE_FAST, load_scene, render_multi, scenes, write_exr


def find_base_dir(scene_path):
    """The reference resolves mesh paths against the process CWD (it is run from the repository root)."""
    here = os.path.dirname(os.path.abspath(scene_path))
    for cand in (os.getcwd(), os.path.dirname(here), here, scenes.ASSETS):
        if os.path.isdir(os.path.join(cand, "objs")):
            return cand
    return scenes.ASSETS


def main(argv=None):
    ap = argparse.ArgumentParser(prog="craytracer_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--scene", "-s", required=True, help=".cry scene file")
    ap.add_argument("--output", default="out.exr")
    ap.add_argument("--preview", action="store_true", help="accepted for compatibility; ignored")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--spp", type=int, default=None, help="override the scene's num_samples")
    ap.add_argument("--mode", choices=["fast", "exact", "f32"], default="fast")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--base-dir", default=None, help="directory mesh paths are resolved against")
    ap.add_argument("--samples-per-call", type=int, default=0, help="split the sample range into render calls of this many samples (default: as many as one call takes)")
    args = ap.parse_args(argv)

    start = time.time()
    scenes.register_standins()
    base = args.base_dir or find_base_dir(args.scene)
    hs = load_scene(args.scene, base_dir=base)
    for w in hs.warnings:
        print(f"[WARN] {w}", file=sys.stderr)
    devices = list(range(max(1, args.gpus)))
    gpu_scenes = Scene.create_multi(hs, devices, build=BUILD_EXACT | BUILD_FAST | (BUILD_F32 if args.mode == "f32" else 0))
    print(f"[INFO] Scene constructed in {time.time() - start:.3f}s", file=sys.stderr)

    sc0 = gpu_scenes[0]
    spp = args.spp if args.spp is not None else sc0.num_samples
    mode = {"exact": TRAVERSE_EXACT, "fast": TRAVERSE_FAST, "f32": TRAVERSE_F32}[args.mode]
    t0 = time.time()
    # one render call takes at most 2^32 - 1 samples per GPU: a frame with more is rendered as several sample ranges whose sums add up
    per_call = args.samples_per_call if args.samples_per_call > 0 else max(1, 0xFFFFFFFF // (sc0.width * sc0.height)) * len(gpu_scenes)
    film, stats, begin = None, [], 0
    while True:
        end = min(spp, begin + per_call)
        part, st = render_multi(gpu_scenes, seed=args.seed, sample_begin=begin, sample_end=end, mode=mode)
        film = part if film is None else film + part
        stats.append(st)
        begin = end
        if end >= spp:
            break
    film = film / np.float32(max(spp, 1))  # pixels /= num_samples (craytracer.rs:253-259)
    dt = time.time() - t0
    rays = sum(s.closest_rays + s.shadow_rays for s in stats)          # the reference's Scene::intersect + intersects calls
    traced = sum(s.closest_rays + s.shadow_rays_traced for s in stats)  # rays traced here (a light sample that cannot contribute needs none)
    dropped = sum(s.nan_samples for s in stats)
    print(f"[INFO] Rendering finished in {dt:.3f}s ({rays / max(dt, 1e-9) / 1e6:.1f} M reference rays/s, {traced / max(dt, 1e-9) / 1e6:.1f} M traced rays/s, "
          f"{sc0.width * sc0.height * spp / max(dt, 1e-9) / 1e6:.1f} Msamples/s"
          + (f", {dropped} samples dropped where the reference would assert" if dropped else "") + ")", file=sys.stderr)
    write_exr(args.output, film)
    print(f"[INFO] Wrote {args.output}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
