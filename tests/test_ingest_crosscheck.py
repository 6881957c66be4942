"""An independent restatement of the reference's mesh ingest, held against the compiled host's (VERDICT r1 "weak" #8: the oracle
and the GPU both consume the scene description the repo's own C++ reader produces, so an ingest error -- winding, the z / v flips,
the MTL -> material mapping, tobj's model splitting -- would be invisible to every parity check).

`load_obj_py` below is written from src/obj.rs:26-220 and the documented behaviour of tobj 4.0 `load_obj(path, &GPU_LOAD_OPTIONS)`
(triangulate = fan from the first corner, single_index = one index per distinct v/vt/vn triple, one model per object / group /
material change) in plain Python floats; it shares no code with craytracer_b200/csrc/host_scene.cpp.  Every triangle record of
the C++ host must equal it bit for bit, in order, with the same material and light bindings."""
import math
import os

import numpy as np
import pytest

import craytracer_b200 as c
from craytracer_b200 import _abi, scenes

REFERENCE = "/root/reference"


def parse_mtl(path):
    mats, cur = [], None
    if not os.path.exists(path):
        return mats
    for raw in open(path):
        parts = raw.split("#")[0].split()
        if not parts:
            continue
        key, args = parts[0], parts[1:]
        if key == "newmtl":
            cur = {"name": " ".join(args), "Kd": None, "Ks": None, "Ke": None, "Ns": None, "Ni": None, "d": None, "illum": None, "map_Kd": None, "map_Ks": None}
            mats.append(cur)
        elif cur is not None and key in ("Kd", "Ks", "Ke"):
            cur[key] = [float(a) for a in args[:3]]
        elif cur is not None and key in ("Ns", "Ni", "d"):
            cur[key] = float(args[0])
        elif cur is not None and key == "illum":
            cur[key] = int(args[0])
        elif cur is not None and key in ("map_Kd", "map_Ks"):
            cur[key] = args[-1]
    return mats


def load_obj_py(path):
    """-> (triangles [(v0, e1, e2, n0, n01, n02, uv0, uv01, uv02, material name or None)], MTL materials)"""
    positions, normals, texcoords = [], [], []
    models, cur = [], None   # model = {"material": name, "corners": [(vi, ti, ni)] three per triangle}
    mtl = []
    material = None

    def start_model():
        nonlocal cur
        cur = {"material": material, "corners": []}
        models.append(cur)

    def index(tok, n):
        i = int(tok)
        return i - 1 if i > 0 else n + i

    for raw in open(path):
        parts = raw.split("#")[0].split()
        if not parts:
            continue
        key, args = parts[0], parts[1:]
        if key == "v":
            positions.append([float(a) for a in args[:3]])
        elif key == "vn":
            normals.append([float(a) for a in args[:3]])
        elif key == "vt":
            texcoords.append([float(a) for a in args[:2]])
        elif key == "mtllib":
            mtl = parse_mtl(os.path.join(os.path.dirname(path), " ".join(args)))
        elif key in ("o", "g"):
            cur = None                      # the next face opens a new model
        elif key == "usemtl":
            name = " ".join(args)
            if name != material:
                material = name
                cur = None                  # tobj closes the running model when the material changes
        elif key == "f":
            if cur is None:
                start_model()
            corners = []
            for tok in args:
                f = tok.split("/")
                vi = index(f[0], len(positions))
                ti = index(f[1], len(texcoords)) if len(f) > 1 and f[1] else None
                ni = index(f[2], len(normals)) if len(f) > 2 and f[2] else None
                corners.append((vi, ti, ni))
            for k in range(1, len(corners) - 1):   # triangulate: fan
                cur["corners"] += [corners[0], corners[k], corners[k + 1]]

    sub = lambda a, b: [a[0] - b[0], a[1] - b[1], a[2] - b[2]]  # noqa: E731
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]  # noqa: E731
    mag2 = lambda a: a[0] * a[0] + a[1] * a[1] + a[2] * a[2]  # noqa: E731
    out = []
    for m in models:
        cs = m["corners"]
        # single_index: a model has normals / texture coordinates only if every corner of it names one
        has_n = bool(cs) and all(c3[2] is not None for c3 in cs)
        has_t = bool(cs) and all(c3[1] is not None for c3 in cs)
        for t in range(0, len(cs), 3):
            (vi, ti, ni), (vj, tj, nj), (vk, tk, nk) = cs[t:t + 3]
            flip = lambda p: [p[0], p[1], -p[2]]  # noqa: E731  right-handed -> left-handed (obj.rs:127-142)
            pi, pj, pk = flip(positions[vi]), flip(positions[vj]), flip(positions[vk])
            e1, e2 = sub(pj, pi), sub(pk, pi)
            if has_n:
                n0, n1, n2 = flip(normals[ni]), flip(normals[nj]), flip(normals[nk])
            else:
                g = cross(sub(pk, pi), sub(pj, pi))     # obj.rs:159
                length = math.sqrt(mag2(g))
                n0 = n1 = n2 = [g[0] / length, g[1] / length, g[2] / length] if length != 0.0 else [math.nan] * 3
            if has_t:
                uv0, uv1, uv2 = [[texcoords[q][0], 1.0 - texcoords[q][1]] for q in (ti, tj, tk)]   # obj.rs:144-151
            else:
                uv0, uv1, uv2 = [0.0, 0.0], [1.0, 0.0], [1.0, 1.0]
            # Shape::new_triangle_with_normals_and_texture_coordinates shape.rs:96-132 drops degenerate triangles
            if mag2(cross(e2, e1)) == 0.0 or mag2(n0) == 0.0 or mag2(n1) == 0.0 or mag2(n2) == 0.0 or any(math.isnan(x) for x in n0):
                continue
            out.append((pi, e1, e2, n0, sub(n1, n0), sub(n2, n0), uv0, [uv1[0] - uv0[0], uv1[1] - uv0[1]], [uv2[0] - uv0[0], uv2[1] - uv0[1]], m["material"]))
    return out, mtl


def expected_material(m):
    """obj.rs:61-105 -> ('light', Ke) | ('glass', eta) | ('metal',) | ('plastic', roughness)"""
    ke = m["Ke"] or [0.0, 0.0, 0.0]
    if any(x != 0.0 for x in ke):
        return ("light", ke)
    if (m["d"] if m["d"] is not None else 1.0) < 1.0:
        return ("glass", m["Ni"] if m["Ni"] is not None else 1.0)
    if m["illum"] in (3, 4, 5, 6, 7, 8, 9):
        return ("metal",)
    ns = m["Ns"] if m["Ns"] is not None else 0.0
    return ("plastic", 180.0 * (1.0 - math.e ** (-ns / 100.0)))


def check_mesh(obj_path, base_dir, mesh_name):
    text = f"""{{ num_samples: 1, camera: Perspective {{ origin: Point(0, 0, -5), target: Point(0, 0, 0), up: Vector(0, 1, 0), fov: 60,
      film: {{ width: 8, height: 8 }} }}, lights: [ Point {{ origin: Point(0, 5, 0), intensity: Color(1, 1, 1) }} ], materials: {{ m: Matte {{ reflectance: Color(1, 1, 1), sigma: 0 }} }}, shapes: {{}},
      primitives: [ Mesh {{ file_name: '{mesh_name}', fallback_material: 'm' }} ] }}"""
    hs = c.parse_scene(text, base_dir=base_dir)
    d = hs.desc
    want, mtl = load_obj_py(obj_path)
    by_name = {m["name"]: m for m in mtl}
    assert d.n_triangles == len(want) == d.n_primitives
    got = np.frombuffer(bytes(np.ctypeslib.as_array(np.ctypeslib.ctypes.cast(d.triangles, np.ctypeslib.ctypes.POINTER(np.ctypeslib.ctypes.c_double)),
                                                     shape=(len(want) * 24,))), dtype=np.float64).reshape(len(want), 24)
    ref = np.array([sum([list(x) for x in w[:9]], []) for w in want], dtype=np.float64)
    # -0.0 against +0.0 is a difference too: compare the bit patterns
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), np.argwhere(got.view(np.uint64) != ref.view(np.uint64))[:5]
    n_lights = 0
    for k, w in enumerate(want):
        p = d.primitives[k]
        assert p.shape_kind == 1 and p.shape_index == k
        m = by_name.get(w[9])
        if m is None:
            assert p.material == 0 and p.area_light < 0          # the fallback material (the scene's only one)
            continue
        kind = expected_material(m)
        if kind[0] == "light":
            n_lights += 1
            assert p.area_light >= 0 and list(d.lights[p.area_light].color) == kind[1] and d.lights[p.area_light].primitive == k
            continue
        mat = d.materials[p.material]
        assert p.area_light < 0
        if kind[0] == "glass":
            assert mat.kind == _abi.CRAY_MAT_GLASS and mat.eta == kind[1]
        elif kind[0] == "metal":
            assert mat.kind == _abi.CRAY_MAT_METAL
        else:
            assert mat.kind == _abi.CRAY_MAT_PLASTIC and mat.t2.a[0] == kind[1]
        if m["map_Kd"] is None:
            assert mat.t0.kind == _abi.CRAY_TEX_CONSTANT and list(mat.t0.a) == (m["Kd"] or [0.0, 0.0, 0.0])
        else:
            assert mat.t0.kind == _abi.CRAY_TEX_IMAGE
    assert d.n_lights == n_lights + 1   # + the scene's own point light
    return len(want)


def test_cornell_box_mesh_with_its_material_library():
    path = os.path.join(scenes.ASSETS, "objs", "local", "cornell", "CornellBox-Original.obj")
    assert check_mesh(path, scenes.ASSETS, "objs/local/cornell/CornellBox-Original.obj") == 32


def test_generated_mesh_with_every_obj_feature(tmp_path):
    rng = np.random.default_rng(4)
    lines = ["mtllib m.mtl", "o first"]
    for _ in range(40):
        lines.append("v " + " ".join(f"{x:.6f}" for x in rng.normal(size=3)))
    for _ in range(12):
        n = rng.normal(size=3)
        lines.append("vn " + " ".join(f"{x:.4f}" for x in n / np.linalg.norm(n)))
    for _ in range(9):
        lines.append("vt " + " ".join(f"{x:.5f}" for x in rng.uniform(-1, 2, size=2)))
    lines += ["usemtl shiny", "f 1/1/1 2/2/2 3/3/3 4/4/4 5/5/5", "f 6//6 7//7 8//8", "usemtl lamp", "f 9 10 11", "f -1 -2 -3", "g second", "usemtl seethrough",
              "f 12/1 13/2 14/3", "f 15/4/9 16/5/10 17/6/11 18/7/12", "usemtl shiny", "f 20 21 22 23 24 25", "f 1 1 2", "o third", "f 30/9/1 31/8/2 32/7/3"]
    (tmp_path / "g.obj").write_text("\n".join(lines) + "\n")
    (tmp_path / "m.mtl").write_text("newmtl shiny\nKd 0.5 0.6 0.7\nKs 3 2 1\nNs 40\nillum 5\n\nnewmtl lamp\nKd 1 1 1\nKe 2 3 4\n\n"
                                    "newmtl seethrough\nKd 0.9 0.9 0.9\nd 0.4\nNi 1.45\nillum 2\n")
    assert check_mesh(str(tmp_path / "g.obj"), str(tmp_path), "g.obj") >= 12


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "objs", "anthropic.obj")), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("name", ["anthropic.obj", "triangle.obj"])
def test_the_reference_meshes(name):
    n = check_mesh(os.path.join(REFERENCE, "objs", name), REFERENCE, "objs/" + name)
    assert n == (20060 if name == "anthropic.obj" else n)
