"""F32 mode against the wide parity mode on one scene's first-vertex rays (GPU box): where do the two disagree?"""
import sys
import numpy as np
sys.path.insert(0, "tests")
import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes

hs = c.parse_scene(scenes.cornell(width=96, height=96), base_dir=scenes.ASSETS)
gpu = c.Scene(hs, build=c.BUILD_EXACT | c.BUILD_FAST | c.BUILD_F32)
orc = o.OracleScene(hs)
ys, xs = np.mgrid[0:96, 0:96]
xs, ys = xs.ravel().astype(np.uint32), ys.ravel().astype(np.uint32)
for sample in range(2):
    shadow, cont = orc.bounce_rays(xs, ys, np.full_like(xs, sample))
    for label, rays in (("shadow", shadow), ("continuation", cont)):
        a, b = gpu.intersects(rays, mode=c.TRAVERSE_FAST), gpu.intersects(rays, mode=c.TRAVERSE_F32)
        ha, hb = gpu.intersect(rays, mode=c.TRAVERSE_FAST), gpu.intersect(rays, mode=c.TRAVERSE_F32)
        bad = np.nonzero(a != b)[0]
        badc = np.nonzero(ha["prim"] != hb["prim"])[0]
        print(f"sample {sample} {label}: {len(rays)} rays, any-hit differs on {len(bad)} (fast occluded / f32 occluded: {int(a[bad].sum())} / {int(b[bad].sum())}), closest differs on {len(badc)}")
        for i in list(bad[:6]):
            r = rays[i]
            print("   any ", r["origin"], r["direction"], r["max_distance"], "fast", a[i], "f32", b[i], "closest fast", ha[i]["prim"], ha[i]["t"], "f32", hb[i]["prim"], hb[i]["t"])
        for i in list(badc[:6]):
            r = rays[i]
            print("   clos", r["origin"], r["direction"], r["max_distance"], "fast", ha[i]["prim"], ha[i]["t"], "f32", hb[i]["prim"], hb[i]["t"])
