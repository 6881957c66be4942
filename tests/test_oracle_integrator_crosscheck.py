"""A second, independent restatement of the reference's path integrator, lights and light sampler, held against the C oracle's
`estimate_Li` sample by sample.

No test of the reference pins a radiance value, and the GPU path is compared against the oracle: a transcription
slip in oracle/scene.hpp's integrator would be inherited silently.  `estimate_li` below is written from the Rust text alone
(src/path_integrator.rs:19-215, src/light.rs:59-219, src/shape.rs:445-514, src/intersection.rs:25-34, src/primitive.rs:32-48,
src/scene.rs:41-42) in plain Python floats; the BxDF / sampling-warp layer is the second restatement of
tests/test_oracle_shading_crosscheck.py.  What it takes from the oracle is what the reference's own tests pin (ray / shape and
ray / scene intersection) plus the sampler values (third-party sobol_burley, unpinned either way) and the camera ray."""
import ctypes as C
import math

import numpy as np
import pytest

import craytracer_b200 as c
import oracle_lib as o
from craytracer_b200 import _abi, scenes
from test_oracle_shading_crosscheck import (BLACK, WHITE, FRAC_1_PI, add, cmul, dot, fdiv, material_f, material_new, material_pdf, material_sample, neg,
                                            normalized, power_heuristic, sample_disk, sample_hemisphere, sample_sphere, scale, sub)

EPSILON = 1e-9   # src/constants.rs
SPHERE, TRIANGLE, DISK = 0, 1, 2


def is_black(col):
    return col[0] == 0.0 and col[1] == 0.0 and col[2] == 0.0


class Sampler:
    """SobolSampler (src/sampling.rs:197-247): value of dimension d = sobol_burley::sample(sample_index, d, hash) widened to f64."""

    def __init__(self, seed, x, y, sample_index, first_dimension):
        self.hash = o.lib().orc_pixel_hash(seed, x, y)
        self.index, self.dim = sample_index, first_dimension

    def sample_1d(self):
        v = float(o.lib().orc_sobol_sample(self.index, self.dim, self.hash))
        self.dim += 1
        return v

    def sample_2d(self):
        return self.sample_1d(), self.sample_1d()


class PyScene:
    def __init__(self, hs, orc):
        d = hs.desc
        self.orc, self.max_depth = orc, d.max_depth
        self.prims = [(d.primitives[k].shape_kind, d.primitives[k].shape_index, d.primitives[k].material, d.primitives[k].area_light) for k in range(d.n_primitives)]
        self.materials = [material_new(d.materials[k]) for k in range(d.n_materials)]
        self.spheres = [([d.spheres[k].origin[0], d.spheres[k].origin[1], d.spheres[k].origin[2]], d.spheres[k].radius) for k in range(d.n_spheres)]
        self.disks = [([d.disks[k].origin[0], d.disks[k].origin[1], d.disks[k].origin[2]], d.disks[k].rotate_x, d.disks[k].rotate_y, d.disks[k].radius,
                       d.disks[k].inner_radius) for k in range(d.n_disks)]
        self.lights = [(d.lights[k].kind, d.lights[k].primitive, list(d.lights[k].v), list(d.lights[k].color)) for k in range(d.n_lights)]
        # Scene::new (scene.rs:41-42): world radius from the root box of the tree, LightSampler::new (light.rs:187-201)
        nodes, _ = orc.bvh()
        diagonal = sub(list(nodes[0]["max"]), list(nodes[0]["min"]))
        world_radius = math.sqrt(dot(diagonal, diagonal)) * 0.5
        total, cdfs = 0.0, []
        for light in self.lights:
            p = self.power(light, world_radius)
            total += (p[0] + p[1] + p[2]) / 3.0
            cdfs.append(total)
        self.cdfs = [x / total for x in cdfs]

    # ---- shapes (shape.rs:445-514); ray / shape intersection is the oracle's, pinned by the reference's own tests ----
    def shape_params(self, prim):
        kind, index = self.prims[prim][0], self.prims[prim][1]
        if kind == SPHERE:
            return kind, np.array(self.spheres[index][0] + [self.spheres[index][1]], dtype=np.float64)
        assert kind == DISK, "only analytic lights in these scenes"
        origin, rx, ry, radius, inner = self.disks[index]
        return kind, np.array(origin + [rx, ry, radius, inner], dtype=np.float64)

    def shape_area(self, prim):
        kind, index = self.prims[prim][0], self.prims[prim][1]
        if kind == SPHERE:
            return math.pi * self.spheres[index][1] ** 2.0          # (sic) shape.rs:506
        return math.pi * (self.disks[index][3] ** 2.0 - self.disks[index][4] ** 2.0)

    def shape_sample(self, prim, s2):
        kind, index = self.prims[prim][0], self.prims[prim][1]
        if kind == SPHERE:
            origin, radius = self.spheres[index]
            return add(scale(sample_sphere(*s2), radius), origin)   # translate(origin) applied to O + sample * radius
        origin, rx, ry, radius, _ = self.disks[index]
        x, y = sample_disk(*s2)
        p = np.array([x * radius, y * radius, 0.0, 1.0])
        # translate * rotate_x(rx) * rotate_y(ry)   (shape.rs:140-144, transformation.rs:311-340)
        a, b = rx * (math.pi / 180.0), ry * (math.pi / 180.0)
        RX = np.array([[1, 0, 0, 0], [0, math.cos(a), -math.sin(a), 0], [0, math.sin(a), math.cos(a), 0], [0, 0, 0, 1.0]])
        RY = np.array([[math.cos(b), 0, math.sin(b), 0], [0, 1, 0, 0], [-math.sin(b), 0, math.cos(b), 0], [0, 0, 0, 1.0]])
        T = np.eye(4)
        T[:3, 3] = origin
        q = (T @ RX @ RY) @ p
        return [float(q[0]), float(q[1]), float(q[2])]

    def shape_pdf_from(self, prim, location, normal, w_i):
        kind, params = self.shape_params(prim)
        ray = o.ray(location, w_i)
        out = np.zeros(9)
        if o.lib().orc_shape_intersect(kind, params.ctypes.data, ray.ctypes.data, out.ctypes.data) != 1:
            return 0.0
        delta = sub([out[0], out[1], out[2]], location)
        return fdiv(dot(delta, delta), abs(dot(w_i, normal)) * self.shape_area(prim))

    # ---- lights (light.rs:59-177) ----
    def power(self, light, world_radius):
        kind, prim, _, color = light
        if kind == _abi.LIGHT_POINT:
            return scale(scale(color, 4.0), math.pi)
        if kind in (_abi.LIGHT_DISTANT, _abi.LIGHT_INFINITE):
            return scale(scale(scale(color, math.pi), world_radius), world_radius)
        return scale(scale(color, math.pi), self.shape_area(prim))

    def light_pick_pdf(self, k):
        return self.cdfs[k] - self.cdfs[k - 1] if k > 0 else self.cdfs[k]

    def light_pick(self, u):
        # binary_search_by(total_cmp).unwrap_or_else(|i| i): an exact match returns its index, otherwise the insertion point
        for k, cdf in enumerate(self.cdfs):
            if cdf == u:
                return k, self.light_pick_pdf(k)
        k = sum(1 for cdf in self.cdfs if cdf < u)
        return k, self.light_pick_pdf(k)

    def light_pdf_li(self, light, location, normal, w_i):
        """-> value, or None for Pdf::Delta"""
        kind, prim = light[0], light[1]
        if kind in (_abi.LIGHT_POINT, _abi.LIGHT_DISTANT):
            return None
        if kind == _abi.LIGHT_INFINITE:
            return FRAC_1_PI / 4.0
        return self.shape_pdf_from(prim, location, normal, w_i)

    def light_le(self, light):
        return light[3] if light[0] == _abi.LIGHT_INFINITE else BLACK

    def light_sample_li(self, light, s1, s2, location, normal):
        """-> Li, w_i, pdf (None = delta), shadow ray (origin, direction, max_distance)"""
        kind, prim, v, color = light
        if kind == _abi.LIGHT_POINT:
            op = sub(v, location)
            d2 = dot(op, op)
            dist = math.sqrt(d2)
            w_i = [op[0] / dist, op[1] / dist, op[2] / dist]
            return [color[0] / d2, color[1] / d2, color[2] / d2], w_i, None, (location, w_i, dist)
        if kind == _abi.LIGHT_DISTANT:
            return color, v, None, (location, v, math.inf)
        if kind == _abi.LIGHT_INFINITE:
            w_i = sample_hemisphere(s2[0], s2[1], [1.0, 0.0, 0.0] if s1 < 0.5 else [-1.0, 0.0, 0.0])
            return color, w_i, FRAC_1_PI / 4.0, (location, w_i, math.inf)
        point = self.shape_sample(prim, s2)
        w_i = normalized(sub(point, location))
        pdf = self.shape_pdf_from(prim, location, normal, w_i)
        delta = sub(point, location)
        return color, w_i, pdf, (location, w_i, math.sqrt(dot(delta, delta)) - EPSILON)

    # ---- Scene::intersect / intersects: the oracle's ----
    def intersect(self, origin, direction):
        hits, surf = self.orc.intersect(o.ray(origin, direction), surface=True)
        if hits["prim"][0] == 0xFFFFFFFF:
            return None
        return int(hits["prim"][0]), list(surf["location"][0]), list(surf["normal"][0]), (float(surf["uv"][0][0]), float(surf["uv"][0][1]))

    def intersects(self, shadow_ray):
        return bool(self.orc.intersects(o.ray(shadow_ray[0], shadow_ray[1], shadow_ray[2]))[0])


def estimate_li(scene, sampler, origin, direction):
    """path_integrator.rs:41-215"""
    L, beta = list(BLACK), list(WHITE)
    bounces, is_specular_bounce, prev_bsdf_pdf, prev = 0, True, 0.0, None
    while bounces < scene.max_depth and not is_black(beta):
        w_o = neg(direction)
        hit = scene.intersect(origin, direction)
        if hit is None:
            for k, light in enumerate(scene.lights):
                Le = scene.light_le(light)
                if is_specular_bounce:
                    L = add(L, cmul(beta, Le))
                elif not is_black(Le):
                    light_pdf = scene.light_pdf_li(light, prev[0], prev[1], w_o) * scene.light_pick_pdf(k)
                    L = add(L, scale(cmul(beta, Le), power_heuristic(1, light_pdf, 1, prev_bsdf_pdf)))
            break
        prim, location, normal, uv = hit
        _, _, material_index, area_light = scene.prims[prim]
        # an emitting primitive carries a black Lambertian (primitive.rs:43-46)
        material = scene.materials[material_index] if area_light < 0 else ("bxdf", ("lambert", list(BLACK)))
        m1, m2 = sampler.sample_1d(), sampler.sample_2d()
        u_light_index = sampler.sample_1d()
        l1, l2 = sampler.sample_1d(), sampler.sample_2d()
        u_rr = sampler.sample_1d()

        if area_light >= 0:
            Le = scene.lights[area_light][3]
            if not is_black(Le):
                if is_specular_bounce:
                    L = add(L, cmul(beta, Le))
                else:
                    light_pdf = scene.light_pdf_li(scene.lights[area_light], location, normal, w_o) * scene.light_pick_pdf(area_light)
                    L = add(L, scale(cmul(beta, Le), power_heuristic(1, light_pdf, 1, prev_bsdf_pdf)))

        k, pick_pdf = scene.light_pick(u_light_index)
        Li, w_i, light_pdf, shadow_ray = scene.light_sample_li(scene.lights[k], l1, l2, location, normal)
        if not scene.intersects(shadow_ray):
            f = material_f(material, w_o, w_i, normal)
            cos_theta = abs(dot(w_i, normal))
            if light_pdf is None:
                L = add(L, [x / pick_pdf for x in scale(cmul(cmul(beta, Li), f), cos_theta)])
            elif light_pdf > 0.0:
                light_pdf = light_pdf * pick_pdf
                bsdf_pdf = material_pdf(material, w_o, w_i, normal)
                weight = power_heuristic(1, light_pdf, 1, 0.0 if bsdf_pdf is None else bsdf_pdf)
                L = add(L, [x / light_pdf for x in scale(scale(cmul(cmul(beta, Li), f), cos_theta), weight)])

        got = material_sample(material, m1, m2, w_o, normal)
        if got is None:
            break
        w_i, f, bsdf_pdf, is_specular = got
        if is_black(f):
            break
        cos_theta = abs(dot(w_i, normal))
        bsdf_pdf = 1.0 if bsdf_pdf is None else bsdf_pdf
        if bsdf_pdf == 0.0:
            break
        beta = [x / bsdf_pdf for x in scale(cmul(beta, f), cos_theta)]
        origin, direction = location, w_i
        is_specular_bounce, prev_bsdf_pdf, prev = is_specular, bsdf_pdf, (location, normal)

        if bounces > 0:
            max_beta = max(beta[0], max(beta[1], beta[2]))
            if max_beta < 1.0:
                q = 1.0 - max_beta
                if u_rr < q:
                    break
                beta = [x / (1.0 - q) for x in beta]
        bounces += 1
    return L


LIGHTS_SCENE = """{ num_samples: 4, max_depth: 6,
  camera: Perspective { origin: Point(0, 2, -9), target: Point(0, 0.5, 0), up: Vector(0, 1, 0), fov: 50, film: { width: 24, height: 16 } },
  lights: [ Point { origin: Point(-3, 4, -2), intensity: Color(9, 7, 5) }, Distant { direction: Vector(0.3, 1, -0.2), intensity: Color(0.4, 0.5, 0.6) } ],
  materials: { floor: Matte { reflectance: Color(0.8, 0.8, 0.8), sigma: 25 }, ball: Plastic { diffuse: Color(0.7, 0.2, 0.2), specular: Color(0.6, 0.6, 0.6), roughness: 30 },
               mirror: Metal { eta: Color(0.2, 0.4, 1.4), k: Color(3.9, 2.4, 1.6) }, glass: Glass { reflectance: Color(1, 1, 1), transmittance: Color(0.9, 0.9, 0.9), eta: 1.5 } },
  shapes: { floor: Disk { origin: Point(0, 0, 0), rotate_x: 90, radius: 30 }, ball: Sphere { origin: Point(-1.5, 1, 0), radius: 1 },
            mirror: Sphere { origin: Point(1.5, 1, 0.5), radius: 1 }, glass: Sphere { origin: Point(0, 0.6, -2.5), radius: 0.6 },
            lamp: Sphere { origin: Point(2, 4, -1), radius: 0.5 }, panel: Disk { origin: Point(-2, 5, 1), rotate_x: 70, rotate_y: 20, radius: 1.2, inner_radius: 0.3 } },
  primitives: [ Shape { shape: 'floor', material: 'floor' }, Shape { shape: 'ball', material: 'ball' }, Shape { shape: 'mirror', material: 'mirror' },
                Shape { shape: 'glass', material: 'glass' }, Shape { shape: 'lamp', emittance: Color(6, 6, 5) }, Shape { shape: 'panel', emittance: Color(2, 3, 4) } ] }"""


def cases():
    return {"simple": (scenes.simple(num_samples=4, width=28, height=16), scenes.ASSETS),
            "materials": (scenes.materials(num_samples=4, width=32, height=20), scenes.ASSETS),
            "lights": (LIGHTS_SCENE, scenes.ASSETS)}


@pytest.mark.parametrize("name", ["simple", "materials", "lights"])
def test_radiance_samples_equal_the_second_restatement(name):
    text, base = cases()[name]
    hs = c.parse_scene(text, base_dir=base)
    orc = o.OracleScene(hs)
    scene = PyScene(hs, orc)
    assert np.allclose(scene.cdfs, orc.light_cdf(), rtol=1e-14, atol=0.0)          # LightSampler::new
    w, h = hs.desc.camera.width, hs.desc.camera.height
    rng = np.random.default_rng(11)
    n = 500
    xs, ys, ss = rng.integers(0, w, size=n), rng.integers(0, h, size=n), rng.integers(0, 64, size=n)
    seed = 3
    want, ok = orc.estimate_Li(xs, ys, ss, seed=seed)
    rays = orc.camera_rays(xs, ys, ss, seed=seed)
    lit, deep = 0, 0
    for k in range(n):
        sampler = Sampler(seed, int(xs[k]), int(ys[k]), int(ss[k]), 4)     # render_pixel drew dimensions 0-3 for the camera (craytracer.rs:148-162)
        got = estimate_li(scene, sampler, list(rays["origin"][k]), list(rays["direction"][k]))
        if not ok[k]:
            assert not all(math.isfinite(x) for x in got), (name, k)          # where the reference would assert (path_integrator.rs:208)
            continue
        assert np.all(np.abs(np.array(got) - want[k]) <= 1e-11 * np.maximum(1.0, np.abs(want[k]))), (name, k, got, want[k])
        lit += 1 if any(x > 0.0 for x in got) else 0
        deep += 1 if sampler.dim >= 4 + 3 * 8 else 0
    assert lit > n // 3 and deep >= 5, (lit, deep)      # the comparison saw light and paths of three vertices and more
    orc.close()
