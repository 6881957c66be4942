#!/usr/bin/env python3
"""Scene construction on the bench workload (host threads of the GPU box; CRAY_BUILD_TIMING=1 prints the phases).  The scene is
created four times in one process -- the first call also pays for the CUDA context --: reference tree on the device (twice), then
on the host (twice)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import craytracer_b200 as c  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402

scenes.register_standins()
t = time.time()
hs = c.parse_scene(scenes.dragon(), base_dir=os.path.join(ROOT, "assets"))
print(f"parse + stand-in mesh {time.time() - t:.2f} s, host threads {os.cpu_count()}")
for flag in ("1", "1", "0", "0"):
    os.environ["CRAY_GPU_BUILD"] = flag
    t = time.time()
    scene = c.Scene(hs)
    print(f"CRAY_GPU_BUILD={flag}: cray_scene_create {time.time() - t:.2f} s (build {scene.info.bvh_build_ms:.0f} ms, upload {scene.info.upload_ms:.0f} ms)", flush=True)
    scene.close()
