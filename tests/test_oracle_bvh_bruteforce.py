"""The oracle's SAH tree and its two walks (src/bvh.rs:58-147, :234-336) against no tree at all: for rays that do not start on a
surface, the closest hit and the any-hit answer cannot depend on how the primitives were partitioned, so a brute-force loop over
every primitive -- the triangle test restated a second time in numpy from src/shape.rs:226-260, element-wise in the reference's
operation order -- must name the same primitive at the same distance, bit for bit.  (The reference's own tests pin the walk on
four rays, tests/test_bvh.rs:38-66; `Bvh::intersects` and the SAH build are pinned by none.)  Rays that start ON a surface can
be culled by the reference's box test (bounds.rs:62-88) and are the business of the false-miss tests, not of this one."""
import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes

NO_HIT = 0xFFFFFFFF
EPSILON = 1e-9


def cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2], a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def triangle_distances(ro, rd, rmax, v0, e1, e2):
    """(rays, triangles) -> distance or inf, shape.rs:226-246 + ray.rs:26-28"""
    ro, rd, rmax = ro[:, None, :], rd[:, None, :], rmax[:, None]
    v0, e1, e2 = v0[None, :, :], e1[None, :, :], e2[None, :, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        P = cross(np.broadcast_to(rd, (rd.shape[0], e2.shape[1], 3)), np.broadcast_to(e2, (rd.shape[0], e2.shape[1], 3)))
        den = dot(P, e1)
        T = ro - v0
        u = dot(P, T) / den
        Q = cross(T, np.broadcast_to(e1, T.shape))
        v = dot(Q, rd) / den
        t = dot(Q, e2) / den
    ok = ~((den > -EPSILON) & (den < EPSILON)) & ~((u < 0.0) | (u > 1.0)) & ~((v < 0.0) | (u + v > 1.0)) & (t > EPSILON) & (t < rmax)
    return np.where(ok, t, np.inf)


def brute_force(hs, rays):
    d = hs.desc
    n_rays = len(rays)
    ro, rd, rmax = rays["origin"].astype(np.float64), rays["direction"].astype(np.float64), rays["max_distance"].astype(np.float64)
    tri = np.ctypeslib.as_array(np.ctypeslib.ctypes.cast(d.triangles, np.ctypeslib.ctypes.POINTER(np.ctypeslib.ctypes.c_double)), shape=(d.n_triangles, 24)) if d.n_triangles else np.zeros((0, 24))
    best_t, best_prim = np.full(n_rays, np.inf), np.full(n_rays, NO_HIT, dtype=np.uint32)
    prim_of_triangle = np.zeros(d.n_triangles, dtype=np.uint32)
    analytic = []
    for k in range(d.n_primitives):
        p = d.primitives[k]
        if p.shape_kind == 1:
            prim_of_triangle[p.shape_index] = k
        else:
            analytic.append((k, p.shape_kind, p.shape_index))
    for start in range(0, n_rays, 256):
        sl = slice(start, min(n_rays, start + 256))
        if d.n_triangles:
            t = triangle_distances(ro[sl], rd[sl], rmax[sl], tri[:, 0:3], tri[:, 3:6], tri[:, 6:9])
            arg = np.argmin(t, axis=1)
            tt = t[np.arange(t.shape[0]), arg]
            best_t[sl] = tt
            best_prim[sl] = np.where(np.isfinite(tt), prim_of_triangle[arg], NO_HIT)
    out = np.zeros(9)
    for k, kind, index in analytic:
        if kind == 0:
            s = d.spheres[index]
            params = np.array(list(s.origin) + [s.radius])
        else:
            s = d.disks[index]
            params = np.array(list(s.origin) + [s.rotate_x, s.rotate_y, s.radius, s.inner_radius])
        for r in range(n_rays):
            if o.lib().orc_shape_intersect(kind, params.ctypes.data, rays[r:r + 1].ctypes.data, out.ctypes.data) == 1 and out[8] < best_t[r]:
                best_t[r], best_prim[r] = out[8], k
    return best_prim, best_t


def ray_batch(hs, orc, n, seed):
    rng = np.random.default_rng(seed)
    nodes, _ = orc.bvh()
    lo, hi = np.maximum(nodes[0]["min"], -300.0), np.minimum(nodes[0]["max"], 300.0)      # (the ground sphere's box is huge)
    rays = np.zeros(n, dtype=c._abi.RAY_DTYPE)
    rays["origin"] = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["max_distance"] = np.where(rng.uniform(size=n) < 0.5, np.inf, rng.uniform(1.0, 80.0, size=n))
    w, h = hs.desc.camera.width, hs.desc.camera.height
    cam = orc.camera_rays(rng.integers(0, w, size=n), rng.integers(0, h, size=n), rng.integers(0, 16, size=n), seed=1)
    return np.concatenate([rays, cam])


@pytest.mark.parametrize("name", ["dragon stand-in", "interior stand-in", "test.cry"])
def test_tree_walks_equal_a_loop_over_every_primitive(name):
    if name == "dragon stand-in":
        c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 9_001, 0)
        hs = c.parse_scene(scenes.dragon(width=48, height=32), base_dir="/nonexistent")
    elif name == "interior stand-in":
        scenes.register_standins(interior_triangles=6_000)
        hs = c.parse_scene(scenes.staircase(width=36, height=64), base_dir=scenes.ASSETS)
    else:
        hs = c.parse_scene(scenes.test_scene(width=48, height=48))
    orc = o.OracleScene(hs)
    rays = ray_batch(hs, orc, 700, 4)
    want_prim, want_t = brute_force(hs, rays)
    hits = orc.intersect(rays)
    assert np.array_equal(hits["prim"], want_prim), np.argwhere(hits["prim"] != want_prim)[:5]
    hit = want_prim != NO_HIT
    assert np.array_equal(hits["t"][hit], want_t[hit])
    assert np.array_equal(orc.intersects(rays), hit)
    # the rest of PrimitiveIntersection for the triangle hits (shape.rs:247-258): location, interpolated normal, uv
    d = hs.desc
    hits2, surf = orc.intersect(rays, surface=True)
    assert np.array_equal(hits2["prim"], want_prim)
    tri = np.ctypeslib.as_array(np.ctypeslib.ctypes.cast(d.triangles, np.ctypeslib.ctypes.POINTER(np.ctypeslib.ctypes.c_double)), shape=(d.n_triangles, 24))
    checked = 0
    for r in np.flatnonzero(hit):
        p = d.primitives[int(want_prim[r])]
        if p.shape_kind != 1:
            continue
        v0, e1, e2, n0, n01, n02 = (tri[p.shape_index, 3 * k:3 * k + 3] for k in range(6))
        uv0, uv01, uv02 = (tri[p.shape_index, 18 + 2 * k:20 + 2 * k] for k in range(3))
        ro, rd = rays["origin"][r], rays["direction"][r]
        P = cross(rd, e2)
        den = dot(P, e1)
        T = ro - v0
        u = dot(P, T) / den
        v = dot(cross(T, e1), rd) / den
        n = n0 + n01 * u + n02 * v
        n = n / np.sqrt(dot(n, n))
        assert np.array_equal(surf["location"][r], ro + rd * want_t[r]) and np.array_equal(surf["normal"][r], n), r
        assert np.array_equal(surf["uv"][r], uv0 + uv01 * u + uv02 * v), r
        assert (hits2["u"][r], hits2["v"][r]) == (u, v)
        checked += 1
    assert checked > 30
    assert 0.2 < hit.mean() < 0.98, hit.mean()
    orc.close()
