"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may use it; the product never does."""
import ctypes as C
import os
import subprocess

import numpy as np

from craytracer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_P = C.c_void_p
_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        L = C.CDLL(ORACLE_LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_create.restype = _P
        L.orc_scene_create.argtypes = [C.POINTER(_abi.SceneDesc), C.c_int]
        L.orc_scene_destroy.argtypes = [_P]
        L.orc_intersect.argtypes = [_P, _P, C.c_uint64, _P, _P, C.c_int]
        L.orc_intersects.argtypes = [_P, _P, C.c_uint64, _P, C.c_int]
        L.orc_camera_rays.argtypes = [_P, C.c_uint64, _P, _P, _P, C.c_uint64, _P]
        L.orc_estimate_li.argtypes = [_P, C.c_uint64, _P, _P, _P, C.c_uint64, _P, _P, C.c_int]
        L.orc_bounce_rays.argtypes = [_P, C.c_uint64, _P, _P, _P, C.c_uint64, _P, _P, _P]
        L.orc_render.argtypes = [_P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _P, _P]
        L.orc_bvh_num_nodes.restype = C.c_uint64
        L.orc_bvh_num_nodes.argtypes = [_P]
        L.orc_bvh_dump.argtypes = [_P, _P, _P]
        L.orc_light_cdf.argtypes = [_P, _P]
        L.orc_camera_matrices.argtypes = [_P, _P]
        L.orc_shape_intersect.argtypes = [C.c_int, _P, _P, _P]
        L.orc_shape_intersects.argtypes = [C.c_int, _P, _P]
        L.orc_shape_bounds.argtypes = [C.c_int, _P, _P]
        L.orc_shape_area.restype = C.c_double
        L.orc_shape_area.argtypes = [C.c_int, _P]
        L.orc_bounds_intersects.argtypes = [_P, _P, _P]
        L.orc_bounds_union.argtypes = [_P, _P, _P]
        L.orc_vector_op.restype = C.c_double
        L.orc_vector_op.argtypes = [C.c_int, _P, _P, C.c_double, _P]
        L.orc_reflect.argtypes = [_P, _P, _P]
        L.orc_refract.argtypes = [_P, _P, C.c_double, C.c_double, C.c_double, _P]
        L.orc_fresnel_dielectric.restype = C.c_double
        L.orc_fresnel_dielectric.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_fresnel_conductor.argtypes = [_P, _P, _P, C.c_double, _P]
        L.orc_matrix_mul.argtypes = [_P, _P, _P]
        L.orc_matrix_inverse.argtypes = [_P, _P]
        L.orc_transformation.argtypes = [C.c_int, _P, _P, _P]
        L.orc_transform_apply.argtypes = [_P, _P, C.c_int, _P, _P]
        L.orc_transform_bounds.argtypes = [_P, _P, _P, _P]
        L.orc_from_rgb.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, _P]
        L.orc_to_rgb.argtypes = [_P, _P]
        L.orc_partition_by.restype = C.c_uint64
        L.orc_partition_by.argtypes = [_P, C.c_uint64, C.c_int, C.c_uint32]
        L.orc_siphash.restype = C.c_uint64
        L.orc_siphash.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_char_p, C.c_uint64]
        L.orc_pixel_hash.restype = C.c_uint32
        L.orc_pixel_hash.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_sobol_sample.restype = C.c_float
        L.orc_sobol_sample.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_sampling_fn.argtypes = [C.c_int, C.c_double, C.c_double, _P, _P]
        L.orc_material_sample.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P]
        L.orc_material_f_pdf.argtypes = [_P, C.c_int, _P, _P, _P, _P, _P]
        _lib = L
    return _lib


def f64(*vals):
    return np.array(vals, dtype=np.float64)


def ray(origin, direction, max_distance=np.inf):
    r = np.zeros(1, dtype=_abi.RAY_DTYPE)
    r["origin"][0] = origin
    r["direction"][0] = direction
    r["max_distance"][0] = max_distance
    return r


class OracleScene:
    """The oracle's Scene::new on the same flat description the product consumes."""

    def __init__(self, host_scene, sah=True):
        self._keep = host_scene
        ptr = host_scene.desc_ptr if hasattr(host_scene, "desc_ptr") else host_scene
        self._h = lib().orc_scene_create(ptr, 1 if sah else 0)
        if not self._h:
            raise RuntimeError(lib().orc_last_error().decode())

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def intersect(self, rays, surface=False, threads=8):
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY_DTYPE)
        hits = np.empty(len(rays), dtype=_abi.HIT_DTYPE)
        surf = np.empty(len(rays), dtype=_abi.SURFACE_DTYPE) if surface else None
        lib().orc_intersect(self._h, rays.ctypes.data, len(rays), hits.ctypes.data, surf.ctypes.data if surface else None, threads)
        return (hits, surf) if surface else hits

    def intersects(self, rays, threads=8):
        rays = np.ascontiguousarray(rays, dtype=_abi.RAY_DTYPE)
        occ = np.empty(len(rays), dtype=np.uint8)
        lib().orc_intersects(self._h, rays.ctypes.data, len(rays), occ.ctypes.data, threads)
        return occ.astype(bool)

    @staticmethod
    def _xys(x, y, s):
        return (np.ascontiguousarray(x, dtype=np.uint32), np.ascontiguousarray(y, dtype=np.uint32), np.ascontiguousarray(s, dtype=np.uint32))

    def camera_rays(self, x, y, s, seed=0):
        x, y, s = self._xys(x, y, s)
        out = np.empty(len(x), dtype=_abi.RAY_DTYPE)
        lib().orc_camera_rays(self._h, seed, x.ctypes.data, y.ctypes.data, s.ctypes.data, len(x), out.ctypes.data)
        return out

    def estimate_Li(self, x, y, s, seed=0, threads=8):
        x, y, s = self._xys(x, y, s)
        rgb = np.empty((len(x), 3), dtype=np.float64)
        ok = np.empty(len(x), dtype=np.uint8)
        lib().orc_estimate_li(self._h, seed, x.ctypes.data, y.ctypes.data, s.ctypes.data, len(x), rgb.ctypes.data, ok.ctypes.data, threads)
        return rgb, ok.astype(bool)

    def bounce_rays(self, x, y, s, seed=0):
        x, y, s = self._xys(x, y, s)
        shadow = np.zeros(len(x), dtype=_abi.RAY_DTYPE)
        cont = np.zeros(len(x), dtype=_abi.RAY_DTYPE)
        valid = np.zeros(len(x), dtype=np.uint8)
        lib().orc_bounce_rays(self._h, seed, x.ctypes.data, y.ctypes.data, s.ctypes.data, len(x), shadow.ctypes.data, cont.ctypes.data, valid.ctypes.data)
        return shadow[(valid & 1) != 0], cont[(valid & 2) != 0]

    def render(self, width, height, seed=0, sample_begin=0, sample_end=1, threads=0):
        film = np.zeros((height, width, 3), dtype=np.float32)
        counts = np.zeros(3, dtype=np.uint64)
        lib().orc_render(self._h, seed, sample_begin, sample_end, threads, film.ctypes.data, counts.ctypes.data)
        return film, counts

    def bvh(self):
        n = lib().orc_bvh_num_nodes(self._h)
        nodes = np.empty(n, dtype=_abi.BVH_NODE_DTYPE)
        order = np.empty(int(self._keep.desc.n_primitives), dtype=np.uint32)
        lib().orc_bvh_dump(self._h, nodes.ctypes.data, order.ctypes.data)
        return nodes, order

    def light_cdf(self):
        cdf = np.empty(int(self._keep.desc.n_lights), dtype=np.float64)
        lib().orc_light_cdf(self._h, cdf.ctypes.data)
        return cdf

    def camera_matrices(self):
        m = np.empty(32, dtype=np.float64)
        lib().orc_camera_matrices(self._h, m.ctypes.data)
        return m[:16].reshape(4, 4), m[16:].reshape(4, 4)


def product_bvh(host_scene, device=-1):
    """cray_build_reference_bvh: the product's BVH builder (host; `device` >= 0: the GPU builder), dumped in the oracle's format."""
    L = _abi.lib()
    nodes_p = C.POINTER(_abi.BvhNodeDump)()
    order_p = C.POINTER(C.c_uint32)()
    n_nodes, n_prims = C.c_uint64(), C.c_uint64()
    rc = L.cray_debug_build_reference_bvh_on(host_scene.desc_ptr, device, C.byref(nodes_p), C.byref(n_nodes), C.byref(order_p), C.byref(n_prims))
    if rc != 0:
        raise RuntimeError(L.cray_last_error().decode())
    nodes = np.frombuffer(C.string_at(nodes_p, n_nodes.value * 64), dtype=_abi.BVH_NODE_DTYPE).copy()
    order = np.frombuffer(C.string_at(order_p, n_prims.value * 4), dtype=np.uint32).copy()
    L.cray_free(nodes_p)
    L.cray_free(order_p)
    return nodes, order
