"""Ad-hoc GPU bring-up probe (not a test): parity of S3/S2 against the oracle on a few scenes + first timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import craytracer_b200 as c
from craytracer_b200 import scenes
import oracle_lib as o

def rand_rays(bounds_lo, bounds_hi, n, seed=1):
    rng = np.random.default_rng(seed)
    org = rng.uniform(bounds_lo, bounds_hi, size=(n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return c.make_rays(org, d)

def compare(name, hs, lo, hi, n_rand=200000, spp_check=2, perf_spp=0):
    t = time.time(); sc = c.Scene(hs); t_create = time.time() - t
    orc = o.OracleScene(hs)
    W, H = sc.width, sc.height
    print(f"[{name}] prims={hs.desc.n_primitives} create={t_create:.2f}s bvh={sc.info.bvh_build_ms:.0f}ms wide_nodes={sc.info.wide_nodes} depth={sc.info.wide_depth}")
    ys, xs = np.mgrid[0:H:4, 0:W:4]
    xs = xs.ravel().astype(np.uint32); ys = ys.ravel().astype(np.uint32); ss = np.zeros_like(xs)
    batches = {"B1": orc.camera_rays(xs, ys, ss), "B2": rand_rays(lo, hi, n_rand)}
    sh, ct = orc.bounce_rays(xs, ys, ss)
    batches["B3s"] = sh; batches["B3c"] = ct
    for bname, rays in batches.items():
        if len(rays) == 0: continue
        ref, rsurf = orc.intersect(rays, surface=True)
        refo = orc.intersects(rays)
        for mode, mname in ((c.TRAVERSE_EXACT, "exact"), (c.TRAVERSE_FAST, "fast")):
            t = time.time(); got, gsurf = sc.intersect(rays, mode=mode, surface=True); dt = time.time() - t
            occ = sc.intersects(rays, mode=mode)
            prim_bad = int((got["prim"] != ref["prim"]).sum())
            hit = ref["prim"] != c.CRAY_NO_HIT
            both = hit & (got["prim"] == ref["prim"])
            t_bits = int((got["t"][both] != ref["t"][both]).sum())
            uv_bad = int(((got["u"][both] != ref["u"][both]) | (got["v"][both] != ref["v"][both])).sum())
            loc_err = float(np.abs(gsurf["location"][both] - rsurf["location"][both]).max()) if both.any() else 0.0
            nrm_err = float(np.abs(gsurf["normal"][both] - rsurf["normal"][both]).max()) if both.any() else 0.0
            occ_bad = int((occ != refo).sum())
            print(f"  {bname:4s} {mname:5s} n={len(rays)} hits={int(hit.sum())} prim_mismatch={prim_bad} t_bitdiff={t_bits} uv_diff={uv_bad} loc_err={loc_err:.2e} nrm_err={nrm_err:.2e} any_mismatch={occ_bad} ({dt*1e3:.0f} ms e2e)")
    # S2 parity
    n = min(len(xs), 20000)
    for mode, mname in ((c.TRAVERSE_EXACT, "exact"), (c.TRAVERSE_FAST, "fast")):
        ref, ok = orc.estimate_Li(xs[:n], ys[:n], ss[:n] + 1)
        got = sc.estimate_Li(xs[:n], ys[:n], ss[:n] + 1, mode=mode)
        good = ok & np.isfinite(got).all(axis=1)
        err = np.abs(got[good] - ref[good]) / (np.abs(ref[good]) + 1e-3)
        print(f"  S2 {mname}: n={n} ok={int(ok.sum())} exact_equal={int((got[good]==ref[good]).all(axis=1).sum())} rel_err>1e-9: {int((err.max(axis=1)>1e-9).sum())} max={err.max():.3e} mean_ref={ref[good].mean():.5f} mean_got={got[good].mean():.5f}")
    # S1
    for mode, mname in ((c.TRAVERSE_EXACT, "exact"), (c.TRAVERSE_FAST, "fast")):
        film, st = sc.render(sample_begin=0, sample_end=spp_check, mode=mode)
        rays = st.closest_rays + st.shadow_rays
        print(f"  S1 {mname}: spp={spp_check} {st.render_ms:.1f} ms, trace {st.trace_ms:.1f} ms, rays={rays} ({rays/st.render_ms/1e3:.1f} Mrays/s) iters={st.iterations} nan={st.nan_samples} mean={film.mean()/spp_check:.5f}")
    t = time.time(); ofilm, counts = orc.render(W, H, sample_begin=0, sample_end=spp_check); dt = time.time() - t
    print(f"  oracle render: {dt:.2f}s rays={int(counts[0]+counts[1])} ({(counts[0]+counts[1])/dt/1e6:.2f} Mrays/s) mean={ofilm.mean()/spp_check:.5f} nan={counts[2]}")
    film, st = sc.render(sample_begin=0, sample_end=spp_check, mode=c.TRAVERSE_EXACT)
    d = np.abs(film - ofilm); print(f"  film exact vs oracle: max abs diff {d.max():.3e}, rel-mse {np.mean((film-ofilm)**2/(ofilm**2+1e-2)):.3e}; counts gpu=({st.closest_rays},{st.shadow_rays}) oracle=({counts[0]},{counts[1]})")
    if perf_spp:
        for rep in range(2):
            film, st = sc.render(sample_begin=0, sample_end=perf_spp, mode=c.TRAVERSE_FAST)
            rays = st.closest_rays + st.shadow_rays
            print(f"  PERF fast spp={perf_spp}: {st.render_ms:.1f} ms, trace(extend) {st.trace_ms:.1f} ms, rays={rays} -> {rays/st.render_ms/1e3:.1f} Mrays/s, {st.samples/st.render_ms/1e3:.2f} Msamples/s, iters={st.iterations}, closest={st.closest_rays} ({st.closest_rays/st.trace_ms/1e3:.1f} M closest rays/s in k_extend)")
    sc.close(); orc.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["simple", "materials", "test", "dragon_small"]
    for name in which:
        if name == "dragon_small":
            c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 200001, 0)
            hs = c.parse_scene(scenes.dragon(width=300, height=200), base_dir="/nonexistent")
            compare(name, hs, [-120, -45, -60], [120, 60, 60])
        elif name == "dragon":
            scenes.register_standins()
            t = time.time(); hs = c.parse_scene(scenes.dragon(), base_dir="/nonexistent"); print("parse+standin", time.time() - t)
            compare(name, hs, [-120, -45, -60], [120, 60, 60], n_rand=1000000, spp_check=1, perf_spp=32)
        else:
            hs = c.parse_scene(scenes.CONFIGS[name](width=200, height=120))
            compare(name, hs, [-10, -2, -10], [10, 10, 20])
