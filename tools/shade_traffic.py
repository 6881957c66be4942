#!/usr/bin/env python3
"""Renders one BASELINE configuration a few samples per pixel, for a capture of its shading kernel under ncu:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:k_shade -s 2 -c 6 \\
        python tools/shade_traffic.py staircase full      # the reference's material library + textures of the reference's sizes
        python tools/shade_traffic.py staircase small     # the small stand-in textures of assets/objs/staircase
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import craytracer_b200 as c  # noqa: E402
from craytracer_b200 import scenes  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "staircase"
    which = sys.argv[2] if len(sys.argv) > 2 else "full"
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    scenes.register_standins()
    base = scenes.write_staircase_assets() if (name == "staircase" and which == "full") else scenes.ASSETS
    hs = c.parse_scene(scenes.CONFIGS[name](num_samples=spp), base_dir=base)
    texels = sum(hs.desc.images[i].width * hs.desc.images[i].height * 3 for i in range(hs.desc.n_images))
    scene = c.Scene(hs)
    _, st = scene.render(seed=0, sample_begin=0, sample_end=spp)
    print(f"{name} ({which}): {hs.desc.n_materials} materials, {hs.desc.n_images} images, {texels / 1e6:.1f} MB of texels; {st.closest_rays} path vertices, "
          f"shade {st.shade_ms:.2f} ms of {st.render_ms:.2f} ms, {st.iterations} iterations")


if __name__ == "__main__":
    main()
