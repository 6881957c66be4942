#!/usr/bin/env python3
"""North-star check at full size (GPU box): dragon.cry, 600x400, SPP samples per pixel rendered by the CUDA path (wide mode)
and by the CPU oracle (the restatement of the reference, all host threads) with the same sampler; prints one JSON line with
the relative MSE and channel-mean ratios.  usage: python tools/converged_parity.py [spp=1024] [scene=dragon]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import craytracer_b200 as c  # noqa: E402
import oracle_lib as o  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    name = sys.argv[2] if len(sys.argv) > 2 else "dragon"
    scenes.register_standins()
    hs = c.parse_scene(scenes.CONFIGS[name](), base_dir=os.path.join(ROOT, "assets"))
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    t0 = time.time()
    film, st = gpu.render(seed=0, sample_begin=0, sample_end=spp)
    t_gpu = time.time() - t0
    t0 = time.time()
    ref, counts = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=spp, threads=os.cpu_count())
    t_cpu = time.time() - t0
    a, b = film / spp, ref / spp
    rel_mse = float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
    rmse = float(np.sqrt(np.mean((a - b) ** 2)))
    line = {"scene": name, "film": [gpu.width, gpu.height], "spp": spp, "triangles": int(hs.desc.n_triangles), "rel_mse": rel_mse, "rmse": rmse,
            "mean_ratio": [float(a[..., k].mean() / b[..., k].mean()) for k in range(3)], "max_abs_diff": float(np.abs(a - b).max()),
            "gpu_rays": int(st.closest_rays + st.shadow_rays), "cpu_rays": int(counts[0] + counts[1]), "dropped_gpu": int(st.nan_samples), "dropped_cpu": int(counts[2]),
            "gpu_s": t_gpu, "cpu_s": t_cpu, "cpu_threads": os.cpu_count(), "gpu_mrays_s": (st.closest_rays + st.shadow_rays) / (st.render_ms * 1e3),
            "cpu_mrays_s": float(counts[0] + counts[1]) / t_cpu / 1e6}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
