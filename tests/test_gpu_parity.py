"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Bars (BASELINE.json north_star): closest-hit primitive index bit-exact, t / barycentrics
within 1e-5 relative (they are bit-equal here), images within a stated relative-MSE at equal spp."""
import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import scenes

pytestmark = pytest.mark.gpu

MODES = [(c.TRAVERSE_EXACT, "exact"), (c.TRAVERSE_FAST, "fast")]
SCENES = {
    "simple": (lambda: c.parse_scene(scenes.simple(width=160, height=96)), [-45, -1, -35], [45, 12, 55]),
    "materials": (lambda: c.parse_scene(scenes.materials(width=160, height=104)), [-8, -1, -6], [12, 16, 16]),
    "test": (lambda: c.parse_scene(scenes.test_scene(width=128, height=128)), [-1, -1, -3], [4, 3, 1]),
    "rounding-error": (lambda: c.parse_scene(scenes.rounding_error(width=128, height=128)), [-10, -1, -10], [10, 8, 10]),
    "dragon_small": (None, [-120, -45, -60], [120, 60, 60]),
    # scenes/staircase.cry over the stand-in assets: 26 MTL materials (plastic / conductor / glass), 10 image textures, thin lens
    "staircase_small": (None, [-1.6, -0.1, -2.1], [1.6, 5.6, 3.1]),
    # scenes/cornell.cry over the authored stand-in mesh; planar walls coincide with BVH box faces, so the reference's AABB rule
    # produces false misses here too (SURVEY A-4b): both modes reproduce them
    "cornell": (lambda: c.parse_scene(scenes.cornell(width=96, height=96), base_dir=scenes.ASSETS), [-1.1, -0.1, -1.1], [1.1, 2.1, 1.1]),
}
FALSE_MISS_SCENES = ("rounding-error", "cornell")


def _dragon_small():
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 120001, 0)
    return c.parse_scene(scenes.dragon(width=150, height=100), base_dir="/nonexistent")


_cache = {}


def _staircase_small():
    c.register_standin_mesh("objs/staircase/staircase.obj", 1, 20000, 0)
    return c.parse_scene(scenes.staircase(width=72, height=128), base_dir=scenes.ASSETS)


def get_scene(name):
    if name not in _cache:
        hs = _dragon_small() if name == "dragon_small" else (_staircase_small() if name == "staircase_small" else SCENES[name][0]())
        _cache[name] = (hs, c.Scene(hs), o.OracleScene(hs))
    return _cache[name]


def random_rays(lo, hi, n, seed):
    rng = np.random.default_rng(seed)
    org = rng.uniform(lo, hi, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return c.make_rays(org, d)


def pixel_grid(scene, step):
    ys, xs = np.mgrid[0:scene.height:step, 0:scene.width:step]
    xs = xs.ravel().astype(np.uint32)
    return xs, ys.ravel().astype(np.uint32), np.zeros_like(xs)


def check_closest(gpu, orc, rays, mode):
    ref, rsurf = orc.intersect(rays, surface=True)
    got, gsurf = gpu.intersect(rays, mode=mode, surface=True)
    assert np.array_equal(got["prim"], ref["prim"]), f"{int((got['prim'] != ref['prim']).sum())} primitive mismatches of {len(rays)}"
    hit = ref["prim"] != c.CRAY_NO_HIT
    assert np.array_equal(got["t"][hit], ref["t"][hit])  # bit-equal, the bar is 1e-5 relative
    du = np.abs(got["u"][hit] - ref["u"][hit])
    dv = np.abs(got["v"][hit] - ref["v"][hit])
    du = np.minimum(du, 1.0 - du)  # sphere / disk u wraps at 1 (atan2 differs by ulps between libm and CUDA)
    assert du.max(initial=0) <= 1e-9 and dv.max(initial=0) <= 1e-9
    scale = 1.0 + np.abs(rsurf["location"][hit]).max(initial=0)
    assert np.abs(gsurf["location"][hit] - rsurf["location"][hit]).max(initial=0) <= 1e-12 * scale
    assert np.abs(gsurf["normal"][hit] - rsurf["normal"][hit]).max(initial=0) <= 1e-12
    return int(hit.sum())


@pytest.mark.parametrize("mode,mode_name", MODES)
@pytest.mark.parametrize("name", list(SCENES))
def test_fixed_ray_batches(name, mode, mode_name):
    """SURVEY 8(d) batches: B1 primary rays, B2 uniform random rays in the scene box, B3 the shadow and continuation rays
    of the first path vertex (origins ON surfaces, finite max_distance)."""
    hs, gpu, orc = get_scene(name)
    _, lo, hi = SCENES[name]
    xs, ys, ss = pixel_grid(gpu, 2)
    b1 = orc.camera_rays(xs, ys, ss)
    b2 = random_rays(lo, hi, 100_000, seed=1)
    b3s, b3c = orc.bounce_rays(xs, ys, ss)
    hits = 0
    for rays in (b1, b2, b3s, b3c):
        if len(rays):
            hits += check_closest(gpu, orc, rays, mode)
            assert np.array_equal(gpu.intersects(rays, mode=mode), orc.intersects(rays))
    assert hits > 0


@pytest.mark.parametrize("mode,mode_name", MODES)
def test_edge_case_rays(mode, mode_name):
    hs, gpu, orc = get_scene("test")
    assert len(gpu.intersect(np.empty(0, dtype=c.RAY_DTYPE), mode=mode)) == 0          # empty batch
    assert len(gpu.intersects(np.empty(0, dtype=c.RAY_DTYPE), mode=mode)) == 0
    o_, d_, m_ = [], [], []
    for axis in range(3):                                                                # axis-aligned rays: zero direction components
        for sign in (1.0, -1.0):
            for org in ([0.25, 0.25, -3.0], [2.5, 2.0, 3.0], [1.5, 5.0, -1.0], [0.0, 0.0, 0.0], [0.5, 0.25, -1.0]):
                d = [0.0, 0.0, 0.0]
                d[axis] = sign
                o_.append(org); d_.append(d); m_.append(np.inf)
    for tmax in (1e-12, 1e-9, 2e-9, 0.5, 2.999999, 3.0, 3.000001):                      # finite / tiny max_distance
        o_.append([0.25, 0.25, -3.0]); d_.append([0.0, 0.0, 1.0]); m_.append(tmax)
    o_.append([0.0, 0.5, 0.0]); d_.append([1.0, 0.0, 0.0]); m_.append(np.inf)           # in the plane of triangle1 (parallel)
    o_.append([0.5, 0.0, -2.0]); d_.append([0.0, 0.0, 3.0]); m_.append(np.inf)          # un-normalised direction, hits an edge
    o_.append([1.0, 0.0, -2.0]); d_.append([0.0, 0.0, 1.0]); m_.append(np.inf)          # exactly through a shared vertex position
    rays = c.make_rays(o_, d_, np.array(m_))
    check_closest(gpu, orc, rays, mode)
    assert np.array_equal(gpu.intersects(rays, mode=mode), orc.intersects(rays))


def test_exact_t_ties_follow_the_reference_visit_order():
    """Two coincident triangles: the reference keeps the one its traversal reaches first (strict `<`, src/ray.rs:26);
    the wide traversal must agree although it visits nodes in a different order."""
    tris = []
    for k in range(12):  # pairs of identical triangles spread along x so that the BVH separates the pairs
        x = 3.0 * k
        for dup in range(2):
            tris.append(f"t{k}_{dup}: Triangle {{ v0: Point({x}, 0, 0), v1: Point({x + 1}, 0, 0), v2: Point({x}, 1, {0.25 * dup * 0}) }}")
    prims = ", ".join(f"Shape {{ shape: 't{k}_{dup}', material: 'm' }}" for k in range(12) for dup in range(2))
    text = ("{ camera: Perspective { origin: Point(0,0,-5), target: Point(0,0,0), up: Vector(0,1,0), fov: 60, film: { width: 8, height: 8 } }, "
            "lights: [ Infinite { intensity: Color(1,1,1) } ], materials: { m: Matte { reflectance: Color(1,1,1), sigma: 0 } }, "
            "shapes: { " + ", ".join(tris) + " }, primitives: [ " + prims + " ] }")
    hs = c.parse_scene(text)
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    rng = np.random.default_rng(3)
    org = np.stack([rng.uniform(0, 36, 4000), rng.uniform(-0.2, 1.2, 4000), rng.choice([-2.0, 2.0], 4000)], axis=1)
    d = np.stack([rng.normal(0, 0.2, 4000), rng.normal(0, 0.2, 4000), -np.sign(org[:, 2])], axis=1)
    rays = c.make_rays(org, d)
    for mode, _ in MODES:
        n = check_closest(gpu, orc, rays, mode)
        assert n > 200


@pytest.mark.parametrize("name", FALSE_MISS_SCENES)
def test_reference_false_misses_are_reproduced_by_both_modes(name):
    """scenes/rounding-error.cry documents a shadow ray that the reference's AABB test wrongly culls (SURVEY A-4b): the ray starts
    within 1e-9 OUTSIDE a node box (a ground point under the ball) and ends inside it.  The exact mode restates that rule for
    every ray; the fast mode finds the rays that start in such a shell (planar contact analysis at build time + a per-ray test)
    and traces those in reference order, so both agree with the oracle on every shadow and bounce ray of every pixel -- and the
    false misses are really there (a conservative traversal would call those rays occluded)."""
    hs, gpu, orc = get_scene(name)
    assert gpu.info.contact_nodes > 0 and gpu.info.contact_primitives > 0
    xs, ys, ss = pixel_grid(gpu, 1)
    shadow, bounce = orc.bounce_rays(xs, ys, ss)
    ref = orc.intersects(shadow)
    for mode, _ in MODES:
        assert np.array_equal(gpu.intersects(shadow, mode=mode), ref)
        check_closest(gpu, orc, bounce, mode)
    # the false misses are really there: shadow rays the reference calls unoccluded although the same ray with an infinite
    # max_distance (whose box tests pass: the exit distance is in range) finds an occluder in front of the light
    probe = shadow.copy()
    probe["max_distance"] = np.inf
    far = orc.intersect(probe)
    occluder_ahead = (far["prim"] != c.CRAY_NO_HIT) & (far["t"] < shadow["max_distance"])
    leaks = int((occluder_ahead & ~ref).sum())
    print(f"{name}.cry: {leaks} of {len(shadow)} shadow rays are false misses of the reference's AABB rule (reproduced)")
    assert leaks > 0


@pytest.mark.parametrize("mode,mode_name", MODES)
@pytest.mark.parametrize("name", ["simple", "materials", "test", "dragon_small", "staircase_small"])
def test_radiance_samples_match_oracle(name, mode, mode_name):
    """S2: render_pixel + estimate_Li per (x, y, sample).  Same sampler integers on both sides, f64 shading with the
    reference's operation order: samples agree to ~1e-12; only libm-vs-CUDA ulp differences in sin/cos/atan2 remain."""
    hs, gpu, orc = get_scene(name)
    xs, ys, _ = pixel_grid(gpu, 2)
    for sample in (0, 5):
        ss = np.full_like(xs, sample)
        ref, ok = orc.estimate_Li(xs, ys, ss, seed=7)
        got = gpu.estimate_Li(xs, ys, ss, seed=7, mode=mode)
        assert ok.all()
        err = np.abs(got - ref) / (np.abs(ref) + 1e-3)
        close = (err.max(axis=1) <= 1e-9)
        assert close.mean() >= 0.999, f"{int((~close).sum())} of {len(xs)} samples differ"
        assert abs(got.mean() - ref.mean()) <= 2e-3 * abs(ref.mean()) + 1e-12


@pytest.mark.parametrize("name", ["simple", "materials", "test", "rounding-error", "dragon_small", "staircase_small"])
def test_film_matches_oracle_exact_mode(name):
    """S1 at equal spp: the exact mode renders the oracle's film (f32 sums; accumulation order is the only difference)."""
    hs, gpu, orc = get_scene(name)
    spp = 4
    film, st = gpu.render(seed=0, sample_begin=0, sample_end=spp, mode=c.TRAVERSE_EXACT)
    ref, counts = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=spp)
    assert st.samples == gpu.width * gpu.height * spp
    assert (st.closest_rays, st.shadow_rays, st.nan_samples) == (int(counts[0]), int(counts[1]), int(counts[2]))
    rel_mse = float(np.mean((film - ref) ** 2 / (ref ** 2 + 1e-2)))
    assert rel_mse <= 1e-10, rel_mse            # stated tolerance for equal-sample renders
    assert np.abs(film - ref).max() <= 1e-4 * max(1.0, float(ref.max()))


@pytest.mark.parametrize("name", ["simple", "materials", "dragon_small", "staircase_small"])
def test_film_fast_mode_within_relative_mse(name):
    """Production (wide BVH) mode vs the oracle at equal spp.  Tolerance: relMSE <= 1e-4 and channel means within 0.5 %
    (SURVEY 8d); on these scenes the two agree far better because the sample sets are identical."""
    hs, gpu, orc = get_scene(name)
    spp = 4
    film, _ = gpu.render(seed=0, sample_begin=0, sample_end=spp, mode=c.TRAVERSE_FAST)
    ref, _ = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=spp)
    rel_mse = float(np.mean((film - ref) ** 2 / (ref ** 2 + 1e-2)))
    assert rel_mse <= 1e-4, rel_mse
    for ch in range(3):
        assert abs(film[..., ch].mean() - ref[..., ch].mean()) <= 5e-3 * abs(ref[..., ch].mean()) + 1e-9


def rel_mse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def test_cornell_converged_image():
    """BASELINE.json config 2 (cornell.cry, path integrator with NEE) at equal spp, compared the way SURVEY 8(d) states:
    relMSE(GPU, oracle) <= max(1e-4, 1.5 x relMSE(oracle seed 0, oracle seed 1)) and channel means within 0.5 %, in BOTH modes.
    Per-sample identity is not available on this scene: the floor and walls coincide with BVH box faces, so whether the
    reference's AABB rule culls a bounce ray hinges on the sign of a ~1e-16 coordinate (SURVEY A-4b), which ulp differences
    between libm and CUDA sin/cos flip for a few samples.  The fast mode traces the rays that start in such a contact shell in
    reference order, so it shows the reference's light leaks like the exact mode does."""
    hs, gpu, orc = get_scene("cornell")
    spp = 64
    ref0, _ = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=spp)
    ref1, _ = orc.render(gpu.width, gpu.height, seed=1, sample_begin=0, sample_end=spp)
    noise = rel_mse(ref1 / spp, ref0 / spp)
    bound = max(1e-4, 1.5 * noise)
    for mode, mode_name in MODES:
        film, st = gpu.render(seed=0, sample_begin=0, sample_end=spp, mode=mode)
        err = rel_mse(film / spp, ref0 / spp)
        ratios = [float(film[..., ch].mean() / ref0[..., ch].mean()) for ch in range(3)]
        print(f"cornell {mode_name}: relMSE {err:.3e} (seed-to-seed noise {noise:.3e}), channel mean ratios {ratios}")
        assert st.nan_samples == 0
        assert err <= bound, (err, bound)
        assert all(abs(r - 1.0) <= 5e-3 for r in ratios), ratios


def test_sample_ranges_are_additive_and_deterministic():
    """Size-independent properties of the sample-range API that multi-GPU sharding relies on."""
    hs, gpu, orc = get_scene("materials")
    whole, st = gpu.render(seed=3, sample_begin=0, sample_end=6)
    a, sa = gpu.render(seed=3, sample_begin=0, sample_end=2)
    b, sb = gpu.render(seed=3, sample_begin=2, sample_end=6)
    assert st.closest_rays == sa.closest_rays + sb.closest_rays and st.shadow_rays == sa.shadow_rays + sb.shadow_rays
    assert np.abs(whole - (a + b)).max() <= 1e-5 * max(1.0, float(whole.max()))
    again, _ = gpu.render(seed=3, sample_begin=0, sample_end=6)
    assert np.abs(whole - again).max() <= 1e-5 * max(1.0, float(whole.max()))
    other, _ = gpu.render(seed=4, sample_begin=0, sample_end=6)
    assert np.abs(whole - other).max() > 1e-3      # the seed matters
    empty, se = gpu.render(seed=3, sample_begin=2, sample_end=2)
    assert se.samples == 0 and not empty.any()


def test_device_buffer_entry_points():
    """The *_device variants take raw device pointers (here: torch CUDA tensors) and run on the caller's stream."""
    import torch
    hs, gpu, orc = get_scene("materials")
    rays = random_rays([-8, -1, -6], [12, 16, 16], 50_000, seed=9)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(len(rays) * 32, dtype=torch.uint8, device="cuda")
    d_occ = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    gpu.intersect_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), mode=c.TRAVERSE_FAST, stream=stream)
    gpu.intersects_device(d_rays.data_ptr(), len(rays), d_occ.data_ptr(), mode=c.TRAVERSE_FAST, stream=stream)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy().view(c.HIT_DTYPE)
    ref = orc.intersect(rays)
    assert np.array_equal(hits["prim"], ref["prim"]) and np.array_equal(hits["t"], ref["t"])
    assert np.array_equal(d_occ.cpu().numpy().astype(bool), orc.intersects(rays))
    film = torch.empty(gpu.height * gpu.width * 3, dtype=torch.float32, device="cuda")
    st = gpu.render_device(film.data_ptr(), seed=0, sample_begin=0, sample_end=2, stream=stream)
    host, _ = gpu.render(seed=0, sample_begin=0, sample_end=2)
    assert st.samples == gpu.width * gpu.height * 2
    assert np.abs(film.cpu().numpy().reshape(host.shape) - host).max() <= 1e-5 * max(1.0, float(host.max()))


def test_full_size_dragon_batches():
    """BASELINE.json's full size: the 7 219 045-triangle dragon stand-in.  Oracle on 200 k rays (seconds), and the
    size-independent property exact mode == wide mode on 2 M rays."""
    scenes.register_standins()
    hs = c.parse_scene(scenes.dragon(), base_dir="/nonexistent")
    assert hs.desc.n_triangles == scenes.DRAGON_TRIANGLES
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    rays = random_rays([-120, -45, -60], [120, 60, 60], 200_000, seed=1)
    for mode, _ in MODES:
        check_closest(gpu, orc, rays, mode)
        assert np.array_equal(gpu.intersects(rays, mode=mode), orc.intersects(rays))
    big = random_rays([-120, -45, -60], [120, 60, 60], 2_000_000, seed=2)
    a = gpu.intersect(big, mode=c.TRAVERSE_EXACT)
    b = gpu.intersect(big, mode=c.TRAVERSE_FAST)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["t"], b["t"])
    xs, ys, _ = pixel_grid(gpu, 4)
    ss = np.zeros_like(xs)
    ref, ok = orc.estimate_Li(xs, ys, ss)
    got = gpu.estimate_Li(xs, ys, ss)
    err = np.abs(got - ref) / (np.abs(ref) + 1e-3)
    assert ok.all() and (err.max(axis=1) <= 1e-9).mean() >= 0.999
    gpu.close()


def test_render_multi_single_gpu_is_render():
    """cray_render_multi with one scene is cray_render."""
    hs, gpu, orc = get_scene("materials")
    a, sa = c.render_multi([gpu], seed=2, sample_begin=0, sample_end=3)
    b, sb = gpu.render(seed=2, sample_begin=0, sample_end=3)
    assert np.array_equal(a, b) and (sa.closest_rays, sa.shadow_rays) == (sb.closest_rays, sb.shadow_rays)


def test_render_multi_shards_samples_over_gpus():
    """SURVEY 8(e) in one process: the scene replicated on every visible GPU, sample slices per GPU, ONE ncclReduce of the f32
    films.  Needs >= 2 GPUs (the round-end run has one; run with `gpurun --gpus 2`)."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    hs = c.parse_scene(scenes.materials(width=160, height=104))
    replicas = c.Scene.create_multi(hs, list(range(n)))  # one host-side build, uploaded to every device
    whole, st = replicas[0].render(seed=4, sample_begin=1, sample_end=11)
    multi, sm = c.render_multi(replicas, seed=4, sample_begin=1, sample_end=11)
    assert (sm.samples, sm.closest_rays, sm.shadow_rays, sm.nan_samples) == (st.samples, st.closest_rays, st.shadow_rays, st.nan_samples)
    assert np.abs(multi - whole).max() <= 1e-5 * max(1.0, float(whole.max()))  # f32 partial sums in a different order
    again, _ = c.render_multi(replicas, seed=4, sample_begin=1, sample_end=11)   # the communicators are reused
    assert np.array_equal(again, multi)
    with pytest.raises(c.CrayError):
        c.render_multi([replicas[0], replicas[0]], seed=0, sample_begin=0, sample_end=2)  # two scenes on one device


def test_api_argument_errors():
    """Bad arguments come back as error codes with a message, never as a crash (the reference panics)."""
    hs, gpu, orc = get_scene("test")
    rays = random_rays([-1, -1, -3], [4, 3, 1], 16, seed=0)
    with pytest.raises(c.CrayError) as e:
        gpu.intersect(rays, mode=7)
    assert e.value.code == c._abi.CRAY_E_INVALID
    with pytest.raises(c.CrayError) as e:
        gpu.render(seed=0, sample_begin=0, sample_end=65537)          # sobol_burley has 2^16 points (sampling.rs:197-247)
    assert "2^16" in e.value.message
    with pytest.raises(c.CrayError):
        gpu.render(seed=0, sample_begin=5, sample_end=2)
    exact_only = c.Scene(hs, build=c.BUILD_EXACT)
    with pytest.raises(c.CrayError) as e:
        exact_only.intersect(rays, mode=c.TRAVERSE_FAST)
    assert "CRAY_BUILD_FAST" in e.value.message
    assert len(exact_only.intersect(rays, mode=c.TRAVERSE_EXACT)) == 16
    film, st = gpu.render(seed=0, sample_begin=65535, sample_end=65536)  # the last sample index is valid
    assert st.samples == gpu.width * gpu.height and np.isfinite(film).all()
    with pytest.raises(c.CrayError):
        c.Scene(hs, device=99)


def test_duplicated_mesh_every_hit_is_a_tie():
    """The same mesh listed twice: every triangle has an exact copy under another primitive index, so EVERY hit is an exact-t tie
    and the winner is whichever copy the reference's traversal order reaches first.  Stresses the tie path of the wide
    traversal (several ties per test round, ties against an earlier round's best) at scale."""
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 20001, 0)
    text = scenes.dragon(width=96, height=64).replace(
        "Mesh { file_name: 'objs/xyzrgb_dragon.obj', fallback_material: 'dragon' }",
        "Mesh { file_name: 'objs/xyzrgb_dragon.obj', fallback_material: 'dragon' },\n    Mesh { file_name: 'objs/xyzrgb_dragon.obj', fallback_material: 'dragon' }")
    hs = c.parse_scene(text, base_dir="/nonexistent")
    assert hs.desc.n_triangles == 2 * 20001
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    xs, ys, ss = pixel_grid(gpu, 1)
    rays = np.concatenate([orc.camera_rays(xs, ys, ss), random_rays([-120, -45, -60], [120, 60, 60], 200_000, seed=11)])
    ref = orc.intersect(rays)
    on_mesh = ref["prim"] >= 2  # primitives 0, 1 are the ground sphere and the light
    assert on_mesh.sum() > 5_000
    for mode, _ in MODES:
        check_closest(gpu, orc, rays, mode)
    film, st = gpu.render(seed=0, sample_begin=0, sample_end=2)
    oref, counts = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=2)
    assert rel_mse(film, oref) <= 1e-8


@pytest.mark.parametrize("lens", ["", ", lens_radius: 0.05, focal_distance: 4"])
def test_orthographic_camera_and_thin_lens(lens):
    """Camera::orthographic (camera.rs:105-129) and the square-aperture thin lens (:156-160) through render_pixel."""
    text = scenes.test_scene(width=64, height=64)
    start = text.index("camera: Perspective")
    end = text.index("film:", start)
    text = text[:start] + "camera: Orthographic { origin: Point(1.5, 1, -3), target: Point(1.5, 1, 0), up: Vector(0, 1, 0), " + text[end:]
    text = text.replace("film: { width: 64, height: 64 }", "film: { width: 64, height: 64 }" + lens, 1)
    hs = c.parse_scene(text)
    assert hs.desc.camera.kind == 1 and (hs.desc.camera.lens_radius > 0) == bool(lens)
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    xs, ys, ss = pixel_grid(gpu, 1)
    rays = orc.camera_rays(xs, ys, ss)
    for mode, _ in MODES:
        check_closest(gpu, orc, rays, mode)
    got = gpu.estimate_Li(xs, ys, ss, seed=1)
    ref, ok = orc.estimate_Li(xs, ys, ss, seed=1)
    err = np.abs(got - ref) / (np.abs(ref) + 1e-3)
    assert ok.all() and (err.max(axis=1) <= 1e-9).mean() >= 0.999


def test_checkerboard_textures_colour_and_scalar():
    """Texture::eval for Checkerboard (texture.rs:22-30) as a colour (reflectance) and as a scalar (Oren-Nayar sigma, plastic
    roughness): the texture coordinates of spheres, disks and triangles all feed it."""
    text = scenes.test_scene(width=80, height=80)
    text = text.replace("white: Matte { reflectance: Color(1, 1, 1), sigma: 100 }",
                        "white: Matte { reflectance: Checkerboard { a: Color(1, 1, 1), b: Color(0.2, 0.3, 0.4), scale: 900 }, sigma: Checkerboard { a: 0, b: 60, scale: 450 } }")
    text = text.replace("red: Matte { reflectance: Color(1, 0, 0), sigma: 0 }",
                        "red: Plastic { diffuse: Checkerboard { a: Color(1, 0, 0), b: Color(0, 0, 1), scale: 3 }, specular: Color(0.5, 0.5, 0.5), roughness: Checkerboard { a: 5, b: 0, scale: 2 } }")
    assert text.count("Checkerboard") == 5
    hs = c.parse_scene(text)
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    xs, ys, _ = pixel_grid(gpu, 1)
    for sample in (0, 3):
        ss = np.full_like(xs, sample)
        ref, ok = orc.estimate_Li(xs, ys, ss, seed=2)
        for mode, _ in MODES:
            got = gpu.estimate_Li(xs, ys, ss, seed=2, mode=mode)
            err = np.abs(got - ref) / (np.abs(ref) + 1e-3)
            assert ok.all() and (err.max(axis=1) <= 1e-9).mean() >= 0.999


def test_device_entry_points_on_a_caller_stream():
    """The *_device calls run on whatever stream the caller hands in (here a fresh non-default torch stream), including the very
    first call on a new scene, which allocates the path pool."""
    import torch
    hs = c.parse_scene(scenes.materials(width=96, height=64))
    gpu, orc = c.Scene(hs), o.OracleScene(hs)
    side = torch.cuda.Stream()
    film = torch.empty(gpu.height * gpu.width * 3, dtype=torch.float32, device="cuda")
    with torch.cuda.stream(side):
        st = gpu.render_device(film.data_ptr(), seed=0, sample_begin=0, sample_end=3, stream=side.cuda_stream)
    side.synchronize()
    ref, counts = orc.render(gpu.width, gpu.height, seed=0, sample_begin=0, sample_end=3)
    assert (st.closest_rays, st.shadow_rays) == (int(counts[0]), int(counts[1]))
    assert rel_mse(film.cpu().numpy().reshape(ref.shape), ref) <= 1e-8
    rays = random_rays([-8, -1, -6], [12, 16, 16], 20_000, seed=5)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_hits = torch.empty(len(rays) * 32, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    gpu.intersect_device(d_rays.data_ptr(), len(rays), d_hits.data_ptr(), mode=c.TRAVERSE_FAST, stream=side.cuda_stream)
    side.synchronize()
    hits = d_hits.cpu().numpy().view(c.HIT_DTYPE)
    want = orc.intersect(rays)
    assert np.array_equal(hits["prim"], want["prim"]) and np.array_equal(hits["t"], want["t"])
