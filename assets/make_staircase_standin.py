#!/usr/bin/env python3
"""Authors the stand-in assets of scenes/staircase.cry: objs/staircase/staircase.mtl (26 materials in the same categories as
the reference's file: illum 2 plastics with Ns 250 / 1000 / 0, five illum 4 conductors, one `d 0.1` glass with Ni 1.1, ten
map_Kd textures, every Ke 0) and ten seeded procedural 128 x 128 textures as binary PPM (the C++ host reads PPM itself; the
reference's JPGs are not redistributed).  The mesh is the host's procedural interior (host_scene.cpp:make_interior_standin).
Run from the repo root:  python assets/make_staircase_standin.py"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "objs", "staircase")
TEXTURES = ["wood", "wood_floor", "wood_panel", "wood_chair", "fabric", "wallpaper", "painting1", "painting2", "painting3", "brushed_metal"]


def texture(kind, index, n=128):
    rng = np.random.default_rng(100 + index)
    y, x = np.mgrid[0:n, 0:n] / n
    base = rng.uniform(0.25, 0.9, size=3)
    if kind.startswith("wood"):
        g = 0.5 + 0.5 * np.sin(40 * (x + 0.15 * np.sin(6 * y + index)) + 3 * rng.normal(size=(n, 1)).cumsum(axis=0) / n)
        img = base[None, None, :] * (0.55 + 0.45 * g[..., None]) * np.array([1.0, 0.75, 0.5])
    elif kind == "fabric":
        g = 0.5 + 0.25 * np.sin(80 * x) + 0.25 * np.sin(80 * y)
        img = base[None, None, :] * g[..., None]
    elif kind == "wallpaper":
        g = ((np.floor(8 * x) + np.floor(8 * y)) % 2)
        img = np.where(g[..., None] > 0, base[None, None, :], base[None, None, ::-1] * 0.6)
    elif kind.startswith("painting"):
        img = np.stack([0.5 + 0.5 * np.sin(2 * np.pi * (k + 1) * (x * (index % 3 + 1) + y * k) + index) for k in range(3)], axis=-1)
    else:  # brushed metal
        g = 0.6 + 0.4 * rng.uniform(size=(n, 1)) * np.ones((1, n))
        img = np.repeat(g[..., None], 3, axis=-1) * 0.8
    return np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8)


def main():
    os.makedirs(os.path.join(OUT, "textures"), exist_ok=True)
    for i, name in enumerate(TEXTURES):
        img = texture(name, i)
        with open(os.path.join(OUT, "textures", name + ".ppm"), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
            f.write(img.tobytes())
    rng = np.random.default_rng(7)
    lines = ["# Stand-in for objs/staircase/staircase.mtl (see assets/make_staircase_standin.py)"]
    for i in range(26):
        kd = rng.uniform(0.1, 0.9, size=3)
        ks = rng.uniform(0.0, 0.6, size=3)
        illum, ns, d, ni, tex = 2, 250.0, 1.0, 1.0, None
        if i < 10:
            tex = TEXTURES[i]
        elif i < 15:
            illum = 4                       # conductor: eta = Kd, k = Ks (src/obj.rs:95-99)
            kd, ks = rng.uniform(0.15, 1.5, size=3), rng.uniform(2.0, 4.0, size=3)
        elif i == 15:
            d, ni = 0.1, 1.1                # glass (src/obj.rs:89-94)
        elif i < 18:
            ns = 1000.0
        elif i < 20:
            ns = 0.0                        # sigma = 0: Lambertian diffuse lobe
        if i % 4 == 3 and illum == 2:
            ks = np.zeros(3)                # black specular: BSDF with the diffuse lobe only
        lines += ["", f"newmtl standin_{i:02d}", "Ka 0 0 0", "Kd %.4f %.4f %.4f" % tuple(kd), "Ks %.4f %.4f %.4f" % tuple(ks), "Ke 0 0 0",
                  f"Ns {ns}", f"Ni {ni}", f"d {d}", f"illum {illum}"]
        if tex:
            lines.append(f"map_Kd textures/{tex}.ppm")
    with open(os.path.join(OUT, "staircase.mtl"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
