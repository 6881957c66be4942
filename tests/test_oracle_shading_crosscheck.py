"""A second, independent restatement of the reference's shading arithmetic, held against the C oracle (oracle/material.hpp,
oracle/sampling.hpp).

No test of the reference pins a BxDF / BSDF / Fresnel value, a sampling warp or a shape area ("the source text is the
only specification"), so a transcription slip in the oracle would go unnoticed -- the GPU is compared against the oracle, and would
inherit it.  The functions below are written from the Rust text alone (src/bxdf.rs:83-382, src/bsdf.rs:15-98,
src/material.rs:20-95, src/sampling.rs:11-65, src/geometry.rs:397-417, src/shape.rs:504-514) in plain Python floats and share no
code with oracle/; both go through the same libm, so they have to agree to the last few bits."""
import ctypes as C
import math

import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import _abi

TOL = 1e-13


# ---- vectors and colours as 3-lists --------------------------------------------------------------------------------
def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def scale(a, s):
    return [a[0] * s, a[1] * s, a[2] * s]


def add(a, b):
    return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]


def sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def neg(a):
    return [-a[0], -a[1], -a[2]]


def cmul(a, b):
    return [a[0] * b[0], a[1] * b[1], a[2] * b[2]]


def cdiv(a, b):
    return [a[0] / b[0], a[1] / b[1], a[2] / b[2]]


def fdiv(a, b):
    """IEEE division (Python raises where Rust returns inf / NaN)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(a) / np.float64(b))


def normalized(a):  # geometry.rs:54-57
    m = math.sqrt(dot(a, a))
    return [a[0] / m, a[1] / m, a[2] / m]


WHITE, BLACK = [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]
FRAC_1_PI = 1.0 / math.pi


# ---- src/geometry.rs:403-417 -----------------------------------------------------------------------------------------
def same_hemisphere(n, v1, v2):
    return dot(n, v1) * dot(n, v2) > 0.0


def generate_tangents(n):
    v = normalized(n)
    sign = math.copysign(1.0, v[2])          # f64::signum: +-1 for +-0 as well
    a = -1.0 / (sign + v[2])
    b = v[0] * v[1] * a
    return [1.0 + sign * v[0] * v[0] * a, sign * b, -sign * v[0]], [b, sign + v[1] * v[1] * a, -v[1]]


# ---- src/sampling.rs:11-65 -------------------------------------------------------------------------------------------
def sample_disk(u, v):
    if u == 0.0 or v == 0.0:
        return 0.0, 0.0
    u, v = 2.0 * u - 1.0, 2.0 * v - 1.0
    if abs(u) > abs(v):
        r, theta = u, math.pi / 4 * v / u
    else:
        r, theta = v, math.pi / 2 - fdiv(math.pi / 4 * u, v)   # (0.5, 0.5) -> 0 / 0: the reference returns NaNs for the exact centre
    if math.isnan(theta):
        return math.nan, math.nan
    return math.cos(theta) * r, math.sin(theta) * r


def sample_sphere(u, v):
    z = 1.0 - 2.0 * u
    r = math.sqrt(max(1.0 - z ** 2.0, 0.0))
    phi = 2.0 * math.pi * v
    return [r * math.cos(phi), r * math.sin(phi), z]


def sample_hemisphere(u, v, n):
    r = sample_sphere(u, v)
    return r if dot(r, n) > 0.0 else neg(r)


def sample_triangle(u, v):
    su = math.sqrt(u)
    return 1.0 - su, v * su


def cosine_sample_hemisphere(u, v, n):
    t, bt = generate_tangents(n)
    x, y = sample_disk(u, v)
    z = math.sqrt(max(1.0 - x * x - y * y, 0.0))
    return normalized(add(add(scale(t, x), scale(bt, y)), scale(n, z)))


def power_heuristic(n_f, pdf_f, n_g, pdf_g):
    f, g = n_f * pdf_f, n_g * pdf_g
    return (f * f) / (f * f + g * g)


# ---- src/bxdf.rs:287-382 ---------------------------------------------------------------------------------------------
def reflect(d, n):
    return sub(scale(n, dot(n, d) * 2.0), d)


def refract(d, n, cos_theta_i, eta_i, eta_t):
    if math.copysign(1.0, cos_theta_i) < 0.0:
        n, eta_rel, cos_theta = neg(n), eta_i / eta_t, -cos_theta_i
    else:
        eta_rel, cos_theta = eta_t / eta_i, cos_theta_i
    sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
    if sin_theta > eta_rel:
        return None
    perp = [x / eta_rel for x in sub(scale(n, cos_theta), d)]
    par = scale(n, -math.sqrt(1.0 - dot(perp, perp)))
    return add(perp, par)


def fresnel_dielectric(eta_i, eta_t, cos_theta_i):
    if math.copysign(1.0, cos_theta_i) < 0.0:
        cos_theta_i, eta_i, eta_t = -cos_theta_i, eta_t, eta_i
    sin_i = math.sqrt(1.0 - cos_theta_i * cos_theta_i)
    sin_t = eta_i / eta_t * sin_i
    if sin_t >= 1.0:
        return 1.0
    cos_t = math.sqrt(1.0 - sin_t * sin_t)
    r_par = (eta_t * cos_theta_i - eta_i * cos_t) / (eta_t * cos_theta_i + eta_i * cos_t)
    r_perp = (eta_i * cos_theta_i - eta_t * cos_t) / (eta_i * cos_theta_i + eta_t * cos_t)
    return (r_par * r_par + r_perp * r_perp) * 0.5


def fresnel_conductor(eta_i, eta_t, k, cos_theta_i):
    out = []
    for ch in range(3):   # Color arithmetic is per channel
        eta_rel = eta_t[ch] / eta_i[ch]
        eta_rel_2 = eta_rel * eta_rel
        k_rel = k[ch] / eta_i[ch]
        k_rel_2 = k_rel * k_rel
        cos2 = cos_theta_i * cos_theta_i
        sin2 = 1.0 - cos2
        t0 = eta_rel_2 - k_rel_2 - 1.0 * sin2
        a2b2 = (t0 * t0 + eta_rel_2 * k_rel_2 * 4.0) ** 0.5
        a = ((a2b2 + t0) * 0.5) ** 0.5
        t1 = a2b2 + 1.0 * cos2
        t2 = a * cos_theta_i * 2.0
        r_perp = (t1 - t2) / (t1 + t2)
        t3 = a2b2 * cos2 + 1.0 * sin2 * sin2
        t4 = a * cos_theta_i * sin2 * 2.0
        r_par = r_perp * (t3 - t4) / (t3 + t4)
        out.append((r_par * r_par + r_perp * r_perp) * 0.5)
    return out


# ---- src/bxdf.rs:83-284: BxDF as (kind, parameters) ------------------------------------------------------------------------
def bxdf_has_reflection(b):
    return b[0] != "btdf"


def bxdf_has_transmission(b):
    return b[0] in ("btdf", "fresnel_specular")


def bxdf_f(b, w_o, w_i, n):
    if b[0] == "lambert":
        return scale(b[1], FRAC_1_PI) if same_hemisphere(n, w_o, w_i) else BLACK
    if b[0] == "oren_nayar":
        if not same_hemisphere(n, w_o, w_i):
            return BLACK
        cos_i, cos_o = abs(dot(w_i, n)), abs(dot(w_o, n))
        sin_i, sin_o = math.sqrt(max(1.0 - cos_i * cos_i, 0.0)), math.sqrt(max(1.0 - cos_o * cos_o, 0.0))
        if sin_i > 1e-4 and sin_o > 1e-4:
            tangent = generate_tangents(n)[0]
            cpi, cpo = abs(dot(w_i, tangent)), abs(dot(w_o, tangent))
            spi, spo = math.sqrt(1.0 - cpi * cpi), math.sqrt(1.0 - cpo * cpo)
            max_cos = max(cpi * cpo + spi * spo, 0.0)
        else:
            max_cos = 0.0
        if cos_i > cos_o:
            sin_alpha, tan_beta = sin_o, sin_i / cos_i
        else:
            sin_alpha, tan_beta = sin_i, sin_o / cos_o
        sigma = b[2] * (math.pi / 180.0)      # f64::to_radians
        s2 = sigma * sigma
        A = 1.0 - s2 / (2.0 * (s2 + 0.33))
        B = 0.45 * s2 / (s2 + 0.09)
        return scale(scale(b[1], A + B * max_cos * sin_alpha * tan_beta), FRAC_1_PI)
    return BLACK


def bxdf_pdf(b, w_o, w_i, n):
    """-> value, or None for Pdf::Delta"""
    if b[0] in ("lambert", "oren_nayar"):
        return FRAC_1_PI * abs(dot(w_i, n))
    return None


def bxdf_sample(b, s2, w_o, n):
    """-> (w_i, f, pdf or None, is_specular) or None"""
    if b[0] in ("lambert", "oren_nayar"):
        w_i = cosine_sample_hemisphere(s2[0], s2[1], n)
        if dot(n, w_o) < 0.0:
            w_i = neg(w_i)
        return w_i, bxdf_f(b, w_o, w_i, n), bxdf_pdf(b, w_o, w_i, n), False
    if b[0] == "conductor":
        w_i = reflect(w_o, n)
        cos_i = abs(dot(w_o, n))
        fr = fresnel_conductor(WHITE, b[1], b[2], cos_i)
        return w_i, [x / cos_i for x in fr], None, True
    if b[0] == "specular":          # dielectric Fresnel 1 -> 1.5 (material.rs:55-61)
        w_i = reflect(w_o, n)
        cos_i = abs(dot(w_o, n))
        fr = scale(WHITE, fresnel_dielectric(b[2], b[3], cos_i))
        return w_i, [x / abs(cos_i) for x in cmul(b[1], fr)], None, True
    if b[0] == "fresnel_specular":
        _, refl, trans, eta_i, eta_t = b
        cos_i = dot(w_o, n)
        fr = fresnel_dielectric(eta_i, eta_t, cos_i)
        if s2[0] < fr:
            return reflect(w_o, n), [x / abs(cos_i) for x in scale(refl, fr)], fr, True
        w_i = refract(w_o, n, cos_i, eta_i, eta_t)
        if w_i is None:
            return None
        return w_i, [x / abs(cos_i) for x in scale(trans, 1.0 - fr)], 1.0 - fr, True
    raise AssertionError(b[0])


# ---- src/bsdf.rs:15-98, src/material.rs:20-95 --------------------------------------------------------------------------
def relevant(bxdfs, w_o, w_i, n):
    reflecting = dot(w_o, n) * dot(w_i, n) > 0.0
    return [(k, b) for k, b in enumerate(bxdfs) if (bxdf_has_reflection(b) if reflecting else bxdf_has_transmission(b))]


def material_new(desc):
    """Material::new_* from the flat description -> ('bxdf', b) | ('bsdf', [b...])"""
    const = lambda t: [t.a[0], t.a[1], t.a[2]]  # noqa: E731
    is_black = lambda col: col == [0.0, 0.0, 0.0]  # noqa: E731
    if desc.kind == _abi.CRAY_MAT_MATTE:
        sigma = desc.t2.a[0]
        return ("bxdf", ("lambert", const(desc.t0)) if sigma == 0.0 else ("oren_nayar", const(desc.t0), sigma))
    if desc.kind == _abi.CRAY_MAT_GLASS:
        return ("bxdf", ("fresnel_specular", const(desc.t0), const(desc.t1), 1.0, desc.eta))
    if desc.kind == _abi.CRAY_MAT_PLASTIC:
        bxdfs = []
        if not is_black(const(desc.t0)):
            bxdfs.append(("oren_nayar", const(desc.t0), desc.t2.a[0]) if desc.t2.a[0] != 0.0 else ("lambert", const(desc.t0)))
        if not is_black(const(desc.t1)):
            bxdfs.append(("specular", const(desc.t1), 1.0, 1.5))
        return ("bsdf", bxdfs)
    return ("bsdf", [("conductor", const(desc.t0), const(desc.t1))])


def material_sample(m, s1, s2, w_o, n):
    if m[0] == "bxdf":
        return bxdf_sample(m[1], s2, w_o, n)
    bxdfs = m[1]
    if not bxdfs:
        return None
    index = int(s1 * len(bxdfs))
    got = bxdf_sample(bxdfs[index], s2, w_o, n)
    if got is None:
        return None
    w_i, f, pdf, spec = got
    if pdf is None:
        return got
    for k, other in relevant(bxdfs, w_o, w_i, n):
        if k != index:
            f = add(f, bxdf_f(other, w_o, w_i, n))
            p = bxdf_pdf(other, w_o, w_i, n)
            if p is not None:
                pdf += p
    return w_i, f, pdf / len(bxdfs), spec


def material_f(m, w_o, w_i, n):
    if m[0] == "bxdf":
        return bxdf_f(m[1], w_o, w_i, n)
    f = BLACK
    for _, b in relevant(m[1], w_o, w_i, n):
        f = add(f, bxdf_f(b, w_o, w_i, n))
    return f


def material_pdf(m, w_o, w_i, n):
    if m[0] == "bxdf":
        return bxdf_pdf(m[1], w_o, w_i, n)
    total, count = 0.0, 0
    for _, b in relevant(m[1], w_o, w_i, n):
        p = bxdf_pdf(b, w_o, w_i, n)
        if p is not None:
            total += p
            count += 1
    return total / count if count else None


# ---- the checks ------------------------------------------------------------------------------------------------------------
MATERIALS = """
  lambert: Matte { reflectance: Color(0.7, 0.4, 0.1), sigma: 0 },
  rough: Matte { reflectance: Color(0.2, 0.9, 0.5), sigma: 20 },
  very_rough: Matte { reflectance: Color(1, 1, 1), sigma: 75 },
  glass: Glass { reflectance: Color(1, 1, 1), transmittance: Color(1, 1, 1), eta: 1.5 },
  dense_glass: Glass { eta: 2.4, reflectance: Color(0.9, 0.8, 0.7), transmittance: Color(0.6, 0.7, 0.8) },
  plastic: Plastic { diffuse: Color(0.3, 0.2, 0.7), specular: Color(0.5, 0.5, 0.6), roughness: 40 },
  smooth_plastic: Plastic { diffuse: Color(0.3, 0.2, 0.7), specular: Color(0.5, 0.5, 0.6), roughness: 0 },
  diffuse_only: Plastic { diffuse: Color(0.9, 0.9, 0.1), specular: Color(0, 0, 0), roughness: 10 },
  specular_only: Plastic { diffuse: Color(0, 0, 0), specular: Color(1, 0.9, 0.8), roughness: 10 },
  no_lobes: Plastic { diffuse: Color(0, 0, 0), specular: Color(0, 0, 0), roughness: 10 },
  gold: Metal { eta: Color(0.143, 0.375, 1.442), k: Color(3.983, 2.386, 1.603) },
  steel: Metal { eta: Color(2.5, 2.5, 2.5), k: Color(3, 3, 3) }
"""


@pytest.fixture(scope="module")
def scene():
    names = [ln.split(":")[0].strip() for ln in MATERIALS.strip().splitlines()]
    text = f"""{{ num_samples: 1, camera: Perspective {{ origin: Point(0, 0, -5), target: Point(0, 0, 0), up: Vector(0, 1, 0), fov: 60, film: {{ width: 8, height: 8 }} }},
      lights: [ Infinite {{ intensity: Color(1, 1, 1) }} ], materials: {{ {MATERIALS} }},
      shapes: {{ {", ".join(f"s{k}: Sphere {{ origin: Point({3 * k}, 0, 0), radius: 1 }}" for k in range(len(names)))} }},
      primitives: [ {", ".join(f"Shape {{ shape: 's{k}', material: '{n}' }}" for k, n in enumerate(names))} ] }}"""
    hs = c.parse_scene(text)
    orc = o.OracleScene(hs)
    # materials are deduplicated / ordered by the host: bind by primitive
    mats = {n: hs.desc.primitives[k].material for k, n in enumerate(names)}
    yield hs, orc, mats
    orc.close()


def unit(rng):
    v = rng.normal(size=3)
    return list(v / np.linalg.norm(v))


def arr(v):
    return np.array(v, dtype=np.float64)


def close(a, b):
    a, b = np.atleast_1d(np.asarray(a, dtype=np.float64)), np.atleast_1d(np.asarray(b, dtype=np.float64))
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    ok = ~np.isnan(b)
    return bool(np.all(np.abs(a[ok] - b[ok]) <= TOL * np.maximum(1.0, np.abs(b[ok]))))


def test_fresnel_terms():
    L = o.lib()
    rng = np.random.default_rng(1)
    for _ in range(500):
        eta_i, eta_t = rng.uniform(1.0, 2.5, size=2)
        cos = float(rng.uniform(-1.0, 1.0))
        assert abs(L.orc_fresnel_dielectric(eta_i, eta_t, cos) - fresnel_dielectric(eta_i, eta_t, cos)) <= TOL
    assert L.orc_fresnel_dielectric(1.5, 1.0, 0.2) == 1.0 == fresnel_dielectric(1.5, 1.0, 0.2)   # total internal reflection
    assert L.orc_fresnel_dielectric(1.0, 1.5, -0.2) == 1.0 == fresnel_dielectric(1.0, 1.5, -0.2)  # ... seen from inside
    out = np.zeros(3)
    for _ in range(500):
        eta_i, eta_t, k = rng.uniform(0.5, 2.0, size=3), rng.uniform(0.1, 3.0, size=3), rng.uniform(0.0, 4.0, size=3)
        cos = float(rng.uniform(0.0, 1.0))
        L.orc_fresnel_conductor(eta_i.ctypes.data, eta_t.ctypes.data, k.ctypes.data, cos, out.ctypes.data)
        assert close(out, fresnel_conductor(list(eta_i), list(eta_t), list(k), cos))


def test_reflect_and_refract():
    L = o.lib()
    rng = np.random.default_rng(2)
    out = np.zeros(3)
    for _ in range(500):
        d, n = arr(unit(rng)), arr(unit(rng))
        L.orc_reflect(d.ctypes.data, n.ctypes.data, out.ctypes.data)
        assert close(out, reflect(list(d), list(n)))
        eta_i, eta_t = rng.uniform(1.0, 2.0, size=2)
        cos = float(np.dot(d, n))
        some = L.orc_refract(d.ctypes.data, n.ctypes.data, cos, eta_i, eta_t, out.ctypes.data)
        want = refract(list(d), list(n), cos, eta_i, eta_t)
        assert (some != 0) == (want is not None)
        if want is not None:
            assert close(out, want)


@pytest.mark.parametrize("which", [0, 1, 2, 3, 4])
def test_sampling_warps(which):
    L = o.lib()
    rng = np.random.default_rng(3 + which)
    out = np.zeros(3)
    samples = [(float(u), float(v)) for u, v in rng.uniform(0.0, 1.0, size=(400, 2))] + [(0.0, 0.3), (0.3, 0.0), (0.5, 0.5), (0.999999, 1e-9)]
    for u, v in samples:
        n = arr(unit(rng) if rng.uniform() < 0.8 else [0.0, 0.0, float(rng.choice([-1.0, 1.0]))])   # the poles of generate_tangents too
        L.orc_sampling_fn(which, u, v, n.ctypes.data, out.ctypes.data)
        if which == 0:
            want = list(sample_disk(u, v)) + [0.0]
        elif which == 1:
            want = sample_sphere(u, v)
        elif which == 2:
            want = sample_hemisphere(u, v, list(n))
        elif which == 3:
            want = list(sample_triangle(u, v)) + [0.0]
        else:
            want = cosine_sample_hemisphere(u, v, list(n))
        assert close(out, want), (which, u, v, n, out, want)


def test_shape_areas():
    L = o.lib()
    sphere = arr([1.0, 2.0, 3.0, 2.5])
    assert abs(L.orc_shape_area(0, sphere.ctypes.data) - math.pi * 2.5 ** 2.0) <= 1e-12    # (sic) shape.rs:506: pi r^2, not 4 pi r^2
    tri = arr([0, 0, 0, 2, 0, 0, 0, 3, 0])
    kinds = {"triangle": 1, "disk": 2}
    assert abs(L.orc_shape_area(kinds["triangle"], tri.ctypes.data) - 3.0) <= 1e-12
    disk = arr([0.0, 1.0, 0.0, 30.0, 60.0, 2.0, 0.5])
    assert abs(L.orc_shape_area(kinds["disk"], disk.ctypes.data) - math.pi * (2.0 ** 2.0 - 0.5 ** 2.0)) <= 1e-12


def test_material_sample_f_pdf_against_the_second_restatement(scene):
    hs, orc, mats = scene
    L = o.lib()
    rng = np.random.default_rng(7)
    uv = arr([0.25, 0.75])
    out8, out5 = np.zeros(8), np.zeros(5)
    counts = {"none": 0, "delta": 0, "specular": 0, "diffuse": 0}
    for name, index in mats.items():
        m = material_new(hs.desc.materials[index])
        for trial in range(300):
            n, w_o = arr(unit(rng)), arr(unit(rng))     # w_o on either side of the surface
            if trial % 50 == 0:
                w_o = arr(normalized(add(list(n), scale(unit(rng), 1e-3))))   # near-normal incidence: sin_theta < 1e-4 in Oren-Nayar
            s = arr(rng.uniform(0.0, 1.0, size=3))
            L.orc_material_sample(orc._h, index, s.ctypes.data, w_o.ctypes.data, n.ctypes.data, uv.ctypes.data, out8.ctypes.data)
            want = material_sample(m, float(s[0]), (float(s[1]), float(s[2])), list(w_o), list(n))
            if want is None:
                assert out8[7] == 0.0, (name, trial)
                counts["none"] += 1
            else:
                w_i, f, pdf, spec = want
                assert out8[7] == 1.0 + (2.0 if pdf is None else 0.0) + (4.0 if spec else 0.0), (name, trial, out8, want)
                assert close(out8[0:3], w_i) and close(out8[3:6], f), (name, trial, out8, want)
                assert close(out8[6], 0.0 if pdf is None else pdf), (name, trial, out8, want)
                counts["delta" if pdf is None else ("specular" if spec else "diffuse")] += 1
            w_i = arr(unit(rng))
            L.orc_material_f_pdf(orc._h, index, w_o.ctypes.data, w_i.ctypes.data, n.ctypes.data, uv.ctypes.data, out5.ctypes.data)
            p = material_pdf(m, list(w_o), list(w_i), list(n))
            assert close(out5[0:3], material_f(m, list(w_o), list(w_i), list(n))), (name, trial)
            assert out5[4] == (1.0 if p is None else 0.0) and close(out5[3], 0.0 if p is None else p), (name, trial)
    assert counts["delta"] > 300 and counts["specular"] > 300 and counts["diffuse"] > 900 and counts["none"] > 0, counts


def test_power_heuristic_is_the_squared_ratio():
    assert power_heuristic(1, 0.5, 1, 0.5) == 0.5 and power_heuristic(1, 3.0, 1, 0.0) == 1.0
