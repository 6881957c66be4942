"""Second, independent restatements (plain Python / numpy, from the Rust text, no code shared with oracle/) of three more parts of
the path that no test of the reference pins, held against the C oracle:

* camera rays      src/camera.rs:25-162 with src/transformation.rs:267-389 (look_at, perspective, orthographic, the raster -> camera
                   matrix with the reference's `film_height = film.width` slip, the 2-pixel box filter, the square-aperture thin lens)
* disk hits        src/shape.rs:263-311 (annulus, uv, the 1e-9 distance rule of src/ray.rs:26-37)
* textures         src/texture.rs:19-47 + src/color.rs:39-46 (checkerboard with saturating casts, nearest-texel image lookup with
                   fract wrap, the 2.2 gamma per lookup), read through a Lambertian's f = texture / pi"""
import math
import struct
import zlib

import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import _abi, scenes
from test_oracle_integrator_crosscheck import Sampler


# ---- src/transformation.rs:267-389 as 4x4 numpy matrices -------------------------------------------------------------------
def translate(x, y, z):
    m = np.eye(4)
    m[:3, 3] = [x, y, z]
    return m


def scale(x, y, z):
    return np.diag([x, y, z, 1.0])


def rotate_x(rad):
    s, co = math.sin(rad), math.cos(rad)
    return np.array([[1, 0, 0, 0], [0, co, -s, 0], [0, s, co, 0], [0, 0, 0, 1.0]])


def rotate_y(rad):
    s, co = math.sin(rad), math.cos(rad)
    return np.array([[co, 0, s, 0], [0, 1, 0, 0], [-s, 0, co, 0], [0, 0, 0, 1.0]])


def unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(v @ v))


def look_at(origin, target, up):
    z = unit(np.asarray(target) - np.asarray(origin))
    x = unit(np.cross(unit(up), z))
    y = unit(np.cross(z, x))
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, origin
    return m


def perspective(fov, near, far):
    m = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, far / (far - near), -far * near / (far - near)], [0, 0, 1.0, 0]])
    inv_tan = 1.0 / math.tan(fov * (math.pi / 180.0) * 0.5)
    return m @ scale(inv_tan, inv_tan, 1.0)


def orthographic(near, far):
    return scale(1.0, 1.0, 1.0 / (far - near)) @ translate(0.0, 0.0, -near)


def point(m, p):
    q = m @ np.array([p[0], p[1], p[2], 1.0])
    return q[:3] / q[3]


def vector(m, v):
    return m[:3, :3] @ np.asarray(v, dtype=np.float64)


def camera_ray(cam, film_sample, lens_sample, x, y):
    """Camera::sample (camera.rs:131-145) -> world-space origin, direction"""
    film_width = float(cam.width)
    film_height = float(cam.width)        # (sic) camera.rs:30
    screen_w, screen_h = (film_width / film_height, 1.0) if film_width > film_height else (1.0, film_height / film_width)
    screen_from_raster = scale(2.0 * screen_w / film_width, -2.0 * screen_h / film_height, 1.0) @ translate(-film_width / 2.0, -film_height / 2.0, 0.0)
    screen_from_camera = perspective(cam.fov, 1e-2, 1000.0) if cam.kind == 0 else orthographic(0.0, 1.0)
    camera_from_raster = np.linalg.inv(screen_from_camera) @ screen_from_raster
    world_from_camera = look_at(list(cam.origin), list(cam.target), list(cam.up))
    dx, dy = 2.0 * film_sample[0] - 1.0, 2.0 * film_sample[1] - 1.0
    p_camera = point(camera_from_raster, [x + dx, y + dy, 0.0])
    origin, direction = p_camera, (unit(p_camera) if cam.kind == 0 else np.array([0.0, 0.0, 1.0]))
    if cam.lens_radius != 0.0:
        lx, ly = 2.0 * lens_sample[0] - 1.0, 2.0 * lens_sample[1] - 1.0
        p_lens = np.array([lx * cam.lens_radius, ly * cam.lens_radius, 0.0])
        p_focal = origin + direction * (cam.focal_distance / direction[2])
        origin, direction = p_lens, unit(p_focal - p_lens)
    return point(world_from_camera, origin), vector(world_from_camera, direction)


CAMERAS = {
    "perspective": "camera: Perspective { origin: Point(3, 2, -7), target: Point(0.5, 0, 1), up: Vector(0.1, 1, 0), fov: 37, film: { width: 90, height: 50 } }",
    "perspective tall film": "camera: Perspective { origin: Point(0, 1, -4), target: Point(0, 0, 0), up: Vector(0, 1, 0), fov: 70, film: { width: 40, height: 100 } }",
    "thin lens": "camera: Perspective { origin: Point(3, 2, -7), target: Point(0.5, 0, 1), up: Vector(0, 1, 0), fov: 37, film: { width: 64, height: 48 }, lens_radius: 0.2, focal_distance: 6 }",
    "orthographic": "camera: Orthographic { origin: Point(1.5, 1, -3), target: Point(1.5, 1, 0), up: Vector(0, 1, 0), film: { width: 64, height: 64 } }",
    "orthographic lens": "camera: Orthographic { origin: Point(1, 1, -3), target: Point(1.5, 0, 0), up: Vector(0, 1, 0), film: { width: 64, height: 32 }, lens_radius: 0.05, focal_distance: 4 }",
}


@pytest.mark.parametrize("name", list(CAMERAS))
def test_camera_rays_equal_the_second_restatement(name):
    text = f"""{{ num_samples: 1, {CAMERAS[name]}, lights: [ Infinite {{ intensity: Color(1, 1, 1) }} ],
      materials: {{ m: Matte {{ reflectance: Color(1, 1, 1), sigma: 0 }} }}, shapes: {{ s: Sphere {{ origin: Point(0, 0, 0), radius: 1 }} }},
      primitives: [ Shape {{ shape: 's', material: 'm' }} ] }}"""
    hs = c.parse_scene(text)
    orc = o.OracleScene(hs)
    cam = hs.desc.camera
    rng = np.random.default_rng(5)
    n, seed = 300, 9
    xs, ys, ss = rng.integers(0, cam.width, size=n), rng.integers(0, cam.height, size=n), rng.integers(0, 4096, size=n)
    rays = orc.camera_rays(xs, ys, ss, seed=seed)
    for k in range(n):
        sampler = Sampler(seed, int(xs[k]), int(ys[k]), int(ss[k]), 0)
        film, lens = sampler.sample_2d(), sampler.sample_2d()      # render_pixel (craytracer.rs:152-153): both drawn whatever the lens
        origin, direction = camera_ray(cam, film, lens, int(xs[k]), int(ys[k]))
        assert np.allclose(rays["origin"][k], origin, rtol=0, atol=1e-11), (name, k, rays["origin"][k], origin)
        assert np.allclose(rays["direction"][k], direction, rtol=0, atol=1e-11), (name, k, rays["direction"][k], direction)
        assert rays["max_distance"][k] == np.inf
    orc.close()


# ---- src/shape.rs:263-311 ----------------------------------------------------------------------------------------------------
def disk_intersect(origin, rx, ry, radius, inner, ray_o, ray_d, max_distance):
    o2w = translate(*origin) @ rotate_x(rx * (math.pi / 180.0)) @ rotate_y(ry * (math.pi / 180.0))
    w2o = np.linalg.inv(o2w)
    oo, od = point(w2o, ray_o), vector(w2o, ray_d)
    if od[2] == 0.0:
        return None
    t = -oo[2] / od[2]
    if not (t > 1e-9 and t < max_distance):
        return None
    lx, ly = oo[0] + od[0] * t, oo[1] + od[1] * t
    d2 = lx ** 2.0 + ly ** 2.0
    if d2 < inner ** 2.0 or d2 > radius ** 2.0:
        return None
    theta = math.atan2(ly, lx)
    if theta < 0.0:
        theta += math.pi * 2.0
    normal = np.linalg.inv(o2w).T[:3, :3] @ np.array([0.0, 0.0, 1.0])       # inverse transpose
    return t, point(o2w, [lx, ly, 0.0]), normal, (theta / (math.pi * 2.0), math.sqrt(d2) / radius)


def test_disk_hits_equal_the_second_restatement():
    L = o.lib()
    rng = np.random.default_rng(8)
    out = np.zeros(9)
    hits = misses = 0
    for trial in range(1500):
        origin = list(rng.uniform(-2, 2, size=3))
        rx, ry = float(rng.uniform(-180, 180)), float(rng.uniform(-180, 180))
        radius = float(rng.uniform(0.5, 3.0))
        inner = float(rng.uniform(0.0, 0.9) * radius) if trial % 3 else 0.0
        params = np.array(origin + [rx, ry, radius, inner])
        ray_o = rng.uniform(-4, 4, size=3)
        aim = np.array(origin) + rng.normal(size=3) * radius * 0.8                # mostly towards the disk
        ray_d = (aim - ray_o) * float(rng.uniform(0.3, 3.0))                        # un-normalised directions too
        max_distance = float(rng.choice([np.inf, 0.6, 2.0]))
        ray = o.ray(ray_o, ray_d, max_distance)
        got = L.orc_shape_intersect(2, params.ctypes.data, ray.ctypes.data, out.ctypes.data)
        want = disk_intersect(origin, rx, ry, radius, inner, ray_o, ray_d, max_distance)
        assert (got == 1) == (want is not None), (trial, got, want)
        if want is None:
            misses += 1
            assert out[8] == max_distance
            continue
        hits += 1
        t, location, normal, uv = want
        assert abs(out[8] - t) <= 1e-11 * max(1.0, t)
        assert np.allclose(out[0:3], location, rtol=0, atol=1e-10) and np.allclose(out[3:6], normal, rtol=0, atol=1e-11)
        assert abs(out[7] - uv[1]) <= 1e-10 and min(abs(out[6] - uv[0]), 1.0 - abs(out[6] - uv[0])) <= 1e-9
    assert hits > 300 and misses > 300, (hits, misses)


# ---- src/texture.rs:19-47, src/color.rs:39-46 ----------------------------------------------------------------------------------
def as_usize(x):
    """f64 as usize: saturating, NaN -> 0"""
    if math.isnan(x) or x <= 0.0:
        return 0
    return min(int(x), 2 ** 64 - 1)


def checkerboard(a, b, scale_, u, v):
    return a if ((as_usize(u * scale_ * 2.0) & 1) ^ (as_usize(v * scale_ * 2.0) & 1)) == 0 else b


def image_lookup(texels, u, v):
    h, w, _ = texels.shape
    u, v = math.fmod(u, 1.0), math.fmod(v, 1.0)      # f64::fract keeps the sign
    if u < 0.0:
        u += 1.0
    if v < 0.0:
        v += 1.0
    x, y = min(as_usize((w - 1) * u), 2 ** 32 - 1), min(as_usize((h - 1) * v), 2 ** 32 - 1)
    return [(float(ch) / 255.0) ** 2.2 for ch in texels[y, x]]


def write_png(path, rgb):
    h, w, _ = rgb.shape
    raw = b"".join(b"\0" + rgb[y].tobytes() for y in range(h))

    def chunk(kind, data):
        return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def lambert_f(orc, material, uv):
    """Material::f of a Lambertian for directions on the same side: texture(uv) / pi"""
    w_o, w_i, n = np.array([0.0, 0.6, 0.8]), np.array([0.6, 0.0, 0.8]), np.array([0.0, 0.0, 1.0])
    uv = np.array(uv, dtype=np.float64)
    out = np.zeros(5)
    o.lib().orc_material_f_pdf(orc._h, material, w_o.ctypes.data, w_i.ctypes.data, n.ctypes.data, uv.ctypes.data, out.ctypes.data)
    return out[:3] * math.pi


def uv_samples(rng, n):
    uvs = [tuple(p) for p in rng.uniform(-3.0, 3.0, size=(n, 2))] + [tuple(p) for p in rng.uniform(0.0, 1.0, size=(n, 2))]
    return uvs + [(0.0, 0.0), (1.0, 1.0), (0.999999999, 0.5), (-0.0, 2.0), (-1.0, -1.0), (0.25, 0.125), (1e-300, -1e-300), (12345.678, -9876.5)]


def test_checkerboard_equals_the_second_restatement():
    text = scenes.test_scene(width=16, height=16)
    hs = c.parse_scene(text)
    orc = o.OracleScene(hs)
    d = hs.desc
    index = next(k for k in range(d.n_materials) if d.materials[k].t0.kind == _abi.CRAY_TEX_CHECKERBOARD)
    t = d.materials[index].t0
    assert d.materials[index].kind == _abi.CRAY_MAT_MATTE and d.materials[index].t2.a[0] == 0.0
    rng = np.random.default_rng(2)
    for u, v in uv_samples(rng, 400):
        want = checkerboard(list(t.a), list(t.b), t.scale, u, v)
        assert np.allclose(lambert_f(orc, index, (u, v)), want, rtol=1e-14, atol=0.0), (u, v)
    orc.close()


def test_image_texture_lookup_equals_the_second_restatement(tmp_path):
    rng = np.random.default_rng(3)
    texels = rng.integers(0, 256, size=(7, 13, 3), dtype=np.uint8)
    texels[0, 0] = [0, 255, 1]
    write_png(tmp_path / "tex.png", texels)
    (tmp_path / "m.mtl").write_text("newmtl textured\nKd 1 1 1\nKs 0 0 0\nNs 0\nillum 2\nmap_Kd tex.png\n")
    (tmp_path / "q.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nusemtl textured\nf 1/1 2/2 3/3 4/4\n")
    text = """{ num_samples: 1, camera: Perspective { origin: Point(0, 0, -5), target: Point(0, 0, 0), up: Vector(0, 1, 0), fov: 60, film: { width: 8, height: 8 } },
      lights: [ Point { origin: Point(0, 5, -5), intensity: Color(1, 1, 1) } ], materials: { fallback: Matte { reflectance: Color(1, 1, 1), sigma: 0 } }, shapes: {},
      primitives: [ Mesh { file_name: 'q.obj', fallback_material: 'fallback' } ] }"""
    hs = c.parse_scene(text, base_dir=str(tmp_path))
    orc = o.OracleScene(hs)
    d = hs.desc
    index = next(k for k in range(d.n_materials) if d.materials[k].t0.kind == _abi.CRAY_TEX_IMAGE)
    m = d.materials[index]
    # MTL -> plastic with a black specular and roughness 180 (1 - e^0) = 0: a single Lambertian lobe (obj.rs:61-105, material.rs:39-64)
    assert m.kind == _abi.CRAY_MAT_PLASTIC and list(m.t1.a) == [0.0, 0.0, 0.0] and m.t2.a[0] == 0.0
    for u, v in uv_samples(rng, 400):
        assert np.allclose(lambert_f(orc, index, (u, v)), image_lookup(texels, u, v), rtol=1e-14, atol=0.0), (u, v)
    orc.close()
