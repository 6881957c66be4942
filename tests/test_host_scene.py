"""Host logic (no GPU): the C-ABI library loads and exports what include/cray_b200.h declares, the product's scene
set-up arithmetic (matrices, BVH build, OBJ/MTL ingest) agrees with the oracle bit for bit, the 8-wide BVH is
structurally sound."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import _abi, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cray_b200.h")).read()
    declared = set(re.findall(r"\b(cray_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    L = _abi.lib()
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, missing
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    assert b"sm_100a" in L.cray_version()


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the device entry points fail loudly (CRAY_E_CUDA); on a GPU box this test is vacuous."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    hs = c.parse_scene(scenes.simple(width=16, height=16))
    with pytest.raises(c.CrayError) as e:
        c.Scene(hs)
    assert e.value.code == _abi.CRAY_E_CUDA


def _product_xform(kind, params):
    p = np.array(params, dtype=np.float64)
    m, inv = np.zeros(16), np.zeros(16)
    _abi.lib().cray_debug_transformation(kind, p.ctypes.data, m.ctypes.data, inv.ctypes.data)
    return m, inv


def _oracle_xform(kind, params):
    p = np.array(params, dtype=np.float64)
    m, inv = np.zeros(16), np.zeros(16)
    o.lib().orc_transformation(kind, p.ctypes.data, m.ctypes.data, inv.ctypes.data)
    return m, inv


def test_transformations_bit_identical_to_oracle():
    rng = np.random.default_rng(11)
    cases = [(0, rng.normal(size=3) * 10), (1, rng.uniform(0.1, 5, 3)), (2, [0.3]), (3, [-1.7]), (4, [2.9]), (2, [np.deg2rad(90.0)]),
             (5, [150, 70, 150, 30, -50, 0, 0, 1, 0]), (5, list(rng.normal(size=6) * 5) + [0.1, 1, -0.2]), (6, [60.0, 1e-2, 1000.0]), (6, [12.5, 1e-2, 1000.0]),
             (7, [0.0, 1.0])]
    for kind, params in cases:
        pm, pi = _product_xform(kind, params)
        om, oi = _oracle_xform(kind, params)
        assert pm.tobytes() == om.tobytes() and pi.tobytes() == oi.tobytes(), (kind, params)


def test_matrix_inverse_bit_identical_to_oracle():
    rng = np.random.default_rng(5)
    for _ in range(200):
        a = rng.normal(size=16) * rng.choice([1e-3, 1.0, 1e3])
        pi, oi = np.zeros(16), np.zeros(16)
        assert _abi.lib().cray_debug_matrix_inverse(a.ctypes.data, pi.ctypes.data) == o.lib().orc_matrix_inverse(a.ctypes.data, oi.ctypes.data) == 1
        assert pi.tobytes() == oi.tobytes()


@pytest.mark.parametrize("name", ["simple", "materials", "dragon", "staircase", "test"])
def test_camera_matrices_bit_identical_to_oracle(name):
    scenes.register_standins(dragon_triangles=2000, interior_triangles=2000)
    hs = c.parse_scene(scenes.CONFIGS[name](), base_dir=os.path.join(ROOT, "assets"))
    out = np.zeros(32)
    _abi.lib().cray_debug_camera_matrices(C.byref(hs.desc.camera), out.ctypes.data)
    cfr, wfc = o.OracleScene(hs).camera_matrices()
    assert out[:16].tobytes() == cfr.tobytes() and out[16:].tobytes() == wfc.tobytes()


def _same_bvh(hs):
    n1, o1 = o.OracleScene(hs).bvh()
    n2, o2 = o.product_bvh(hs)
    assert len(n1) == len(n2)
    assert np.array_equal(o1, o2), "leaf order differs"
    for f in ("min", "max", "axis", "a", "b"):
        assert np.array_equal(n1[f], n2[f]), f
    return n1


@pytest.mark.parametrize("name", ["simple", "materials", "test", "rounding-error"])
def test_reference_bvh_matches_oracle_small(name):
    _same_bvh(c.parse_scene(scenes.CONFIGS[name]()))


def test_reference_bvh_matches_oracle_mesh_and_parallel_build():
    # > 65 536 primitives takes the multi-threaded sub-tree path of the product builder
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 150001, 0)
    hs = c.parse_scene(scenes.dragon(), base_dir="/nonexistent")
    assert hs.desc.n_triangles == 150001 and any("stand-in" in w for w in hs.warnings)
    nodes = _same_bvh(hs)
    leaves = nodes[nodes["axis"] == 3]
    assert leaves["b"].max() <= 4 and leaves["b"].sum() == hs.desc.n_primitives  # MAX_LEAF_PRIMITIVES, bvh.rs:237


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "objs", "anthropic.obj")), reason="reference assets not mounted")
def test_reference_bvh_matches_oracle_on_the_reference_mesh():
    hs = c.parse_scene(scenes.anthropic(), base_dir=REFERENCE)
    assert hs.desc.n_triangles == 20060
    _same_bvh(hs)


@pytest.mark.parametrize("name", ["simple", "materials", "test", "dragon"])
def test_wide_bvh_structure(name):
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 60001, 0)
    hs = c.parse_scene(scenes.CONFIGS[name](), base_dir="/nonexistent")
    out = np.zeros(4, dtype=np.uint64)
    rc = _abi.lib().cray_debug_check_wide_bvh(hs.desc_ptr, out.ctypes.data)
    assert rc == 0, _abi.lib().cray_last_error()
    assert out[0] >= 1 and out[1] >= 1 and out[1] < 30


def _contacts(hs):
    out = (C.c_uint64 * 3)()
    flags = np.zeros(hs.desc.n_primitives, dtype=np.uint8)
    assert _abi.lib().cray_debug_find_contacts(hs.desc_ptr, out, flags.ctypes.data) == 0
    return int(out[0]), int(out[1]), flags


def test_planar_contact_analysis_marks_the_false_miss_configurations():
    """bvh_build.hpp "planar contact": the node boxes a planar primitive touches and the primitives whose outgoing rays can start in
    the 1e-9 outer shell of such a box -- the only place where the reference's box test (bounds.rs:62-88 with bvh.rs:70,:117) can
    cull what a conservative traversal enters (SURVEY A-4b)."""
    # scenes/rounding-error.cry: the ground disk (primitive 0, y = -6e-17 |z|) under the ball's box (min.y = 0 exactly)
    nodes, prims, flags = _contacts(c.parse_scene(scenes.rounding_error()))
    assert nodes > 0 and flags.tolist() == [1, 0, 0]
    # cornell stand-in: floor, ceiling and walls are axis-aligned planes that coincide with faces of the root box and of inner boxes
    nodes, prims, flags = _contacts(c.parse_scene(scenes.cornell(), base_dir=scenes.ASSETS))
    assert nodes > 0 and prims >= 8
    # curved grounds (r = 1e5 spheres) touch boxes at points only: no primitive is marked
    assert _contacts(c.parse_scene(scenes.materials()))[1] == 0
    # the dragon stand-in: no planar primitive but the light disk, and no path ray ever leaves an area light
    c.register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 30001, 0)
    nodes, prims, flags = _contacts(c.parse_scene(scenes.dragon(), base_dir="/nonexistent"))
    assert prims == 0 and not flags.any()
    # a box resting on a ground quad: both ground triangles are marked (they span the box's footprint), the box's own faces are
    # planar too but only its bottom face lies in a plane that another, thicker node box shares
    text = """{ camera: Perspective { origin: Point(0, 3, -6), target: Point(0, 0, 0), up: Vector(0, 1, 0), fov: 50, film: { width: 8, height: 8 } },
      lights: [ Point { origin: Point(0, 5, 0), intensity: Color(10, 10, 10) } ],
      materials: { m: Matte { reflectance: Color(1, 1, 1), sigma: 0 } },
      shapes: { g0: Triangle { v0: Point(-4, 0, -4), v1: Point(4, 0, -4), v2: Point(4, 0, 4) }, g1: Triangle { v0: Point(-4, 0, -4), v1: Point(4, 0, 4), v2: Point(-4, 0, 4) },
                ball: Sphere { origin: Point(0, 1, 0), radius: 1 } },
      primitives: [ Shape { shape: 'g0', material: 'm' }, Shape { shape: 'g1', material: 'm' }, Shape { shape: 'ball', material: 'm' } ] }"""
    nodes, prims, flags = _contacts(c.parse_scene(text))
    assert flags.tolist() == [1, 1, 0] and nodes >= 1


def test_scene_description_limits_are_errors():
    """max_depth beyond the sampler's 256 dimensions (8 per bounce) is refused before any device is touched."""
    hs = c.parse_scene(scenes.simple(width=16, height=16).replace("num_samples:", "max_depth: 32, num_samples:", 1))
    assert hs.desc.max_depth == 32
    with pytest.raises(c.CrayError) as e:
        c.Scene(hs)
    assert e.value.code == _abi.CRAY_E_UNSUPPORTED and "max_depth" in e.value.message


def test_bvh_degenerate_input_is_an_error_not_a_crash():
    # > 4 primitives sharing one centroid: the reference asserts (bvh.rs:327-328); the ABI reports CRAY_E_BVH
    prims = ", ".join("Shape { shape: 's', material: 'm' }" for _ in range(6))
    hs = c.parse_scene(_minimal(prims))
    nodes_p, order_p = C.POINTER(_abi.BvhNodeDump)(), C.POINTER(C.c_uint32)()
    n1, n2 = C.c_uint64(), C.c_uint64()
    rc = _abi.lib().cray_build_reference_bvh(hs.desc_ptr, C.byref(nodes_p), C.byref(n1), C.byref(order_p), C.byref(n2))
    assert rc == _abi.CRAY_E_BVH
    with pytest.raises(RuntimeError):
        o.OracleScene(hs)


def _minimal(prims="Shape { shape: 's', material: 'm' }"):
    return ("{ camera: Perspective { origin: Point(0,0,-5), target: Point(0,0,0), up: Vector(0,1,0), fov: 60, film: { width: 8, height: 8 } }, "
            "lights: [ Infinite { intensity: Color(1,1,1) } ], materials: { m: Matte { reflectance: Color(1,1,1), sigma: 0 } }, "
            "shapes: { s: Sphere { origin: Point(0,0,3), radius: 1 } }, primitives: [ " + prims + " ] }")


OBJ = """# quad + triangle with normals and uvs, negative indices, two materials
mtllib test.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 1
v 3 0 1
v 3 1 1
v 5 0 2
v 6 0 2
v 6 1 3
v 8 0 2
v 9 0 2
v 9 1 5
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
o first
usemtl shiny
f 1/1/1 2/2/1 3/3/1 4/4/1
usemtl lamp
f -9/1/1 -8/2/1 -7/3/1
o second
usemtl glassy
f 8 9 10
f 1 1 2
"""
MTL = """newmtl shiny
Kd 0.5 0.6 0.7
Ks 3 2 1
Ns 250
illum 4
newmtl lamp
Kd 1 1 1
Ke 2 3 4
newmtl glassy
Kd 0.9 0.9 0.9
d 0.5
Ni 1.45
newmtl textured
Kd 1 1 1
map_Kd tex.ppm
Ns 100
illum 2
"""


def test_obj_mtl_ingest(tmp_path):  # src/obj.rs:26-220
    (tmp_path / "test.obj").write_text(OBJ + "usemtl textured\nf 11/1 12/2 13/3\n")
    (tmp_path / "test.mtl").write_text(MTL)
    with open(tmp_path / "tex.ppm", "wb") as f:
        f.write(b"P6\n2 2\n255\n" + bytes([255, 0, 0, 0, 255, 0, 0, 0, 255, 128, 128, 128]))
    text = _minimal("Mesh { file_name: 'test.obj', fallback_material: 'm' }")
    hs = c.parse_scene(text, base_dir=tmp_path)
    d = hs.desc
    # quad -> 2 triangles (fan), lamp triangle, one good + one degenerate (dropped, shape.rs:110-116), textured triangle
    assert d.n_triangles == 5 and d.n_primitives == 5
    t0 = d.triangles[0]
    assert list(t0.v0) == [0, 0, -0.0] and list(t0.e1) == [1, 0, 0] and list(t0.e2) == [1, 1, 0]
    assert list(t0.n0) == [0, 0, -1.0] and list(t0.n01) == [0, 0, 0]          # z flipped (obj.rs:135-142)
    assert list(t0.uv0) == [0, 1.0] and list(t0.uv01) == [1.0, 0.0] and list(t0.uv02) == [1.0, -1.0]  # v -> 1 - v (obj.rs:144-151)
    shiny = d.materials[d.primitives[0].material]
    assert shiny.kind == _abi.MAT_METAL and list(shiny.t0.a) == [0.5, 0.6, 0.7] and list(shiny.t1.a) == [3, 2, 1]
    lamp = d.primitives[2]
    assert lamp.area_light == 1 and d.lights[1].kind == _abi.LIGHT_AREA and list(d.lights[1].color) == [2, 3, 4] and d.lights[1].primitive == 2
    glassy = d.materials[d.primitives[3].material]
    assert glassy.kind == _abi.MAT_GLASS and glassy.eta == 1.45
    flat = d.triangles[3]   # no vn in this model: geometric normal of (vk - vi) x (vj - vi), uv defaults
    assert list(flat.uv01) == [1.0, 0.0] and list(flat.uv02) == [1.0, 1.0] and abs(np.linalg.norm(list(flat.n0)) - 1.0) < 1e-15 and list(flat.n01) == [0, 0, 0]
    tex = d.materials[d.primitives[4].material]
    assert tex.kind == _abi.MAT_PLASTIC and tex.t0.kind == _abi.TEX_IMAGE and d.n_images == 1 and d.images[0].width == 2
    assert abs(tex.t2.a[0] - 180.0 * (1.0 - np.e ** -1.0)) < 1e-12          # roughness from Ns (obj.rs:84)
    # the oracle accepts the same description
    o.OracleScene(hs)


def test_missing_mesh_is_an_io_error():
    with pytest.raises(c.CrayError) as e:
        c.parse_scene(_minimal("Mesh { file_name: 'nope.obj', fallback_material: 'm' }"), base_dir="/nonexistent")
    assert e.value.code == _abi.CRAY_E_IO


def test_shard_samples_tile_the_range():
    from craytracer_b200.distributed import shard_samples
    for spp in (1, 7, 10, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_samples(spp, r, world, sample_begin=5) for r in range(world)]
            assert ranges[0][0] == 5 and ranges[-1][1] == 5 + spp
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_ctypes_mirror_matches_the_header_layout():
    """The structs of craytracer_b200/_abi.py are the same size as the C structs they mirror (cray_abi_struct_sizes)."""
    out = np.zeros(32, dtype=np.uint32)
    n = _abi.lib().cray_abi_struct_sizes(out.ctypes.data, 32)
    mirror = [C.sizeof(_abi.SphereDesc), C.sizeof(_abi.TriangleDesc), C.sizeof(_abi.DiskDesc), C.sizeof(_abi.PrimitiveDesc), C.sizeof(_abi.TextureDesc),
              C.sizeof(_abi.ImageDesc), C.sizeof(_abi.MaterialDesc), C.sizeof(_abi.LightDesc), C.sizeof(_abi.CameraDesc), C.sizeof(_abi.SceneDesc),
              _abi.RAY_DTYPE.itemsize, _abi.HIT_DTYPE.itemsize, _abi.SURFACE_DTYPE.itemsize, C.sizeof(_abi.RenderStats), C.sizeof(_abi.SceneInfo),
              C.sizeof(_abi.BvhNodeDump)]
    assert n == len(mirror) and list(out[:n]) == mirror


def test_ctypes_constants_match_the_header():
    """Build flags, traversal modes and error codes of craytracer_b200/_abi.py are the header's."""
    import re
    header = open(os.path.join(ROOT, "include", "cray_b200.h")).read()
    defines = {m.group(1): m.group(2) for m in re.finditer(r"#define\s+(CRAY_\w+)\s+\(?(-?\d+)u?\)?", header)}
    assert (int(defines["CRAY_BUILD_EXACT"]), int(defines["CRAY_BUILD_FAST"]), int(defines["CRAY_BUILD_F32"])) == (_abi.BUILD_EXACT, _abi.BUILD_FAST, _abi.BUILD_F32)
    modes = re.search(r"enum\s*\{\s*CRAY_TRAVERSE_EXACT\s*=\s*(\d+),\s*CRAY_TRAVERSE_FAST\s*=\s*(\d+),\s*CRAY_TRAVERSE_F32\s*=\s*(\d+)\s*\}", header)
    assert tuple(int(g) for g in modes.groups()) == (_abi.TRAVERSE_EXACT, _abi.TRAVERSE_FAST, _abi.TRAVERSE_F32)
    assert int(defines["CRAY_E_PARSE"]) == _abi.CRAY_E_PARSE


def _staircase_library(base_dir):
    c.register_standin_mesh("objs/staircase/staircase.obj", 1, 3000, 0)
    hs = c.parse_scene(scenes.staircase(width=72, height=128), base_dir=base_dir)
    d = hs.desc
    mats = bytes(C.string_at(d.materials, C.sizeof(_abi.MaterialDesc) * d.n_materials))
    dims = [(d.images[i].width, d.images[i].height) for i in range(d.n_images)]
    return hs, mats, dims


def test_staircase_material_library_and_jpeg_textures_load_in_the_compiled_host():
    """objs/staircase/staircase.mtl as the reference ships it (26 materials: illum 4 -> Metal, d < 1 -> Glass, otherwise Plastic
    with map_Kd, src/obj.rs:61-105) with ten JPEG textures of the reference's sizes, decoded by the C++ host itself."""
    base = scenes.write_staircase_assets()
    hs, mats, dims = _staircase_library(base)
    d = hs.desc
    assert d.n_materials == 1 + len(scenes.STAIRCASE_MATERIALS)  # the scene's `default` + the library
    assert sorted(dims) == sorted((w, h) for (w, h, *_rest) in scenes.STAIRCASE_TEXTURES.values())
    assert sum(w * h for w, h in dims) == 27_642_338
    kinds = [d.materials[1 + i].kind for i in range(len(scenes.STAIRCASE_MATERIALS))]
    for (name, _ns, _kd, _ks, _ni, dissolve, illum, tex), kind, i in zip(scenes.STAIRCASE_MATERIALS, kinds, range(len(kinds))):
        want = _abi.CRAY_MAT_METAL if illum == 4 else (_abi.CRAY_MAT_GLASS if dissolve < 1.0 else _abi.CRAY_MAT_PLASTIC)
        assert kind == want, name
        if tex and kind == _abi.CRAY_MAT_PLASTIC:
            t0 = d.materials[1 + i].t0
            assert t0.kind == _abi.CRAY_TEX_IMAGE and (d.images[t0.image].width, d.images[t0.image].height) == scenes.STAIRCASE_TEXTURES[tex][:2], name
    # a texel of a decoded texture is what PIL reads from the same file
    Image = pytest.importorskip("PIL.Image")
    tex_dir = os.path.join(base, "objs", "staircase", "textures")
    by_dims = {}
    for n in os.listdir(tex_dir):
        im = np.asarray(Image.open(os.path.join(tex_dir, n)).convert("RGB"))
        by_dims[(im.shape[1], im.shape[0])] = im
    for i in range(d.n_images):
        w, h = d.images[i].width, d.images[i].height
        got = np.ctypeslib.as_array(d.images[i].rgb, shape=(h, w, 3))
        assert np.array_equal(got, by_dims[(w, h)])


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "objs", "staircase", "staircase.mtl")), reason="reference assets not mounted")
def test_generated_staircase_library_equals_the_reference_file():
    """The material table of craytracer_b200/scenes.py against the reference's own staircase.mtl + JPEGs: same materials byte for
    byte (kinds, colours, roughness, eta, texture bindings) and the same texture dimensions."""
    _, ref_mats, ref_dims = _staircase_library(REFERENCE)
    _, gen_mats, gen_dims = _staircase_library(scenes.write_staircase_assets())
    assert ref_dims == gen_dims
    assert ref_mats == gen_mats
