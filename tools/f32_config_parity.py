#!/usr/bin/env python3
"""The opt-in F32 traversal mode against the parity (wide f64) mode on every BASELINE.json configuration at its own film size
(GPU box): same sampler, same spp, both through cray_render; one JSON line per scene with the relative MSE of the F32 film
against the parity film, channel-mean ratios, ray counts and rates.  usage: python tools/f32_config_parity.py [spp=64]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import craytracer_b200 as c  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    scenes.register_standins()
    for name in ("simple", "materials", "cornell", "staircase", "dragon"):
        hs = c.parse_scene(scenes.CONFIGS[name](), base_dir=os.path.join(ROOT, "assets"))
        gpu = c.Scene(hs, build=c.BUILD_EXACT | c.BUILD_FAST | c.BUILD_F32)
        n = spp if name != "staircase" else max(1, spp // 4)
        for warm in (c.TRAVERSE_FAST, c.TRAVERSE_F32):
            gpu.render(seed=1, sample_begin=0, sample_end=1, mode=warm)
        ref, st_ref = gpu.render(seed=0, sample_begin=0, sample_end=n, mode=c.TRAVERSE_FAST)
        film, st = gpu.render(seed=0, sample_begin=0, sample_end=n, mode=c.TRAVERSE_F32)
        a, b = film / n, ref / n
        line = {"scene": name, "film": [gpu.width, gpu.height], "spp": n, "triangles": int(hs.desc.n_triangles),
                "rel_mse_f32_vs_parity": float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2))),
                "mean_ratio": [float(a[..., k].mean() / b[..., k].mean()) for k in range(3)], "max_abs_diff": float(np.abs(a - b).max()),
                "pixels_differing": int((np.abs(a - b).max(axis=-1) > 1e-6).sum()) if a.ndim == 3 else int((np.abs(a - b).reshape(-1, 3).max(axis=1) > 1e-6).sum()),
                "rays_parity": int(st_ref.closest_rays + st_ref.shadow_rays), "rays_f32": int(st.closest_rays + st.shadow_rays),
                "dropped": [int(st_ref.nan_samples), int(st.nan_samples)],
                "parity_mrays_s": (st_ref.closest_rays + st_ref.shadow_rays) / (st_ref.render_ms * 1e3),
                "f32_mrays_s": (st.closest_rays + st.shadow_rays) / (st.render_ms * 1e3)}
        print(json.dumps(line), flush=True)
        gpu.close()


if __name__ == "__main__":
    main()
