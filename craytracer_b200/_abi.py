"""ctypes mirror of include/cray_b200.h and loader of the in-tree CUDA library.

The library is the product: there is no Python or CPU fallback.  If ``libcray_b200.so`` has not been
built (``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C craytracer_b200/csrc``) the
import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcray_b200.so")

CRAY_OK = 0
CRAY_E_INVALID, CRAY_E_CUDA, CRAY_E_PARSE, CRAY_E_IO, CRAY_E_BVH, CRAY_E_UNSUPPORTED = -1, -2, -3, -4, -5, -6
CRAY_NO_HIT = 0xFFFFFFFF
CRAY_TEX_CONSTANT, CRAY_TEX_CHECKERBOARD, CRAY_TEX_IMAGE = 0, 1, 2
CRAY_MAT_MATTE, CRAY_MAT_GLASS, CRAY_MAT_PLASTIC, CRAY_MAT_METAL = 0, 1, 2, 3
SHAPE_SPHERE, SHAPE_TRIANGLE, SHAPE_DISK = 0, 1, 2
TEX_CONSTANT, TEX_CHECKERBOARD, TEX_IMAGE = 0, 1, 2
MAT_MATTE, MAT_GLASS, MAT_PLASTIC, MAT_METAL = 0, 1, 2, 3
LIGHT_POINT, LIGHT_DISTANT, LIGHT_INFINITE, LIGHT_AREA = 0, 1, 2, 3
CAMERA_PERSPECTIVE, CAMERA_ORTHOGRAPHIC = 0, 1
BUILD_EXACT, BUILD_FAST, BUILD_F32 = 1, 2, 4
TRAVERSE_EXACT, TRAVERSE_FAST, TRAVERSE_F32 = 0, 1, 2

D3 = C.c_double * 3
D2 = C.c_double * 2


class SphereDesc(C.Structure):
    _fields_ = [("origin", D3), ("radius", C.c_double)]


class TriangleDesc(C.Structure):
    _fields_ = [("v0", D3), ("e1", D3), ("e2", D3), ("n0", D3), ("n01", D3), ("n02", D3), ("uv0", D2), ("uv01", D2), ("uv02", D2)]


class DiskDesc(C.Structure):
    _fields_ = [("origin", D3), ("rotate_x", C.c_double), ("rotate_y", C.c_double), ("radius", C.c_double), ("inner_radius", C.c_double)]


class PrimitiveDesc(C.Structure):
    _fields_ = [("shape_kind", C.c_uint32), ("shape_index", C.c_uint32), ("material", C.c_int32), ("area_light", C.c_int32)]


class TextureDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("image", C.c_int32), ("a", D3), ("b", D3), ("scale", C.c_double)]


class ImageDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb", C.POINTER(C.c_uint8))]


class MaterialDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("_pad", C.c_uint32), ("t0", TextureDesc), ("t1", TextureDesc), ("t2", TextureDesc), ("eta", C.c_double)]


class LightDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("primitive", C.c_int32), ("v", D3), ("color", D3)]


class CameraDesc(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("_pad", C.c_uint32), ("origin", D3), ("target", D3), ("up", D3),
                ("fov", C.c_double), ("lens_radius", C.c_double), ("focal_distance", C.c_double)]


class SceneDesc(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("num_samples", C.c_uint32), ("camera", CameraDesc),
                ("n_spheres", C.c_uint64), ("n_triangles", C.c_uint64), ("n_disks", C.c_uint64), ("n_primitives", C.c_uint64),
                ("n_materials", C.c_uint64), ("n_lights", C.c_uint64), ("n_images", C.c_uint64),
                ("spheres", C.POINTER(SphereDesc)), ("triangles", C.POINTER(TriangleDesc)), ("disks", C.POINTER(DiskDesc)),
                ("primitives", C.POINTER(PrimitiveDesc)), ("materials", C.POINTER(MaterialDesc)), ("lights", C.POINTER(LightDesc)),
                ("images", C.POINTER(ImageDesc))]


class RenderStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nan_samples", C.c_uint64),
                ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64), ("render_ms", C.c_double), ("trace_ms", C.c_double),
                ("shadow_ms", C.c_double), ("shade_ms", C.c_double), ("generate_ms", C.c_double), ("shadow_rays_traced", C.c_uint64), ("contact_rays", C.c_uint64)]


class SceneInfo(C.Structure):
    _fields_ = [("n_primitives", C.c_uint64), ("n_lights", C.c_uint64), ("exact_nodes", C.c_uint64), ("exact_bytes", C.c_uint64),
                ("wide_nodes", C.c_uint64), ("wide_bytes", C.c_uint64), ("leaf_prim_bytes", C.c_uint64), ("wide_depth", C.c_uint64),
                ("width", C.c_uint32), ("height", C.c_uint32), ("max_depth", C.c_uint32), ("num_samples", C.c_uint32),
                ("bvh_build_ms", C.c_double), ("upload_ms", C.c_double), ("contact_nodes", C.c_uint64), ("contact_primitives", C.c_uint64)]


class BvhNodeDump(C.Structure):
    _fields_ = [("min", D3), ("max", D3), ("axis", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("_pad", C.c_uint32)]


# numpy views of the plain-data ray / hit records
RAY_DTYPE = np.dtype([("origin", "<f8", 3), ("direction", "<f8", 3), ("max_distance", "<f8")])
HIT_DTYPE = np.dtype([("prim", "<u4"), ("_pad", "<u4"), ("t", "<f8"), ("u", "<f8"), ("v", "<f8")])
SURFACE_DTYPE = np.dtype([("location", "<f8", 3), ("normal", "<f8", 3), ("uv", "<f8", 2)])
BVH_NODE_DTYPE = np.dtype([("min", "<f8", 3), ("max", "<f8", 3), ("axis", "<u4"), ("a", "<u4"), ("b", "<u4"), ("_pad", "<u4")])
assert RAY_DTYPE.itemsize == 56 and HIT_DTYPE.itemsize == 32 and SURFACE_DTYPE.itemsize == 64 and BVH_NODE_DTYPE.itemsize == 64

IMAGE_DECODER = C.CFUNCTYPE(C.c_int, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_uint8)))

# every symbol include/cray_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "cray_host_scene_parse": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "cray_host_scene_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "cray_host_scene_desc": (C.POINTER(SceneDesc), [_P]),
    "cray_host_scene_destroy": (None, [_P]),
    "cray_last_error_location": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "cray_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.c_int, C.c_uint32, C.POINTER(_P)]),
    "cray_scene_destroy": (None, [_P]),
    "cray_trace_closest": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P, _P]),
    "cray_trace_any": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P]),
    "cray_trace_closest_device": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P, _P, _P]),
    "cray_trace_any_device": (C.c_int, [_P, C.c_int, _P, C.c_uint64, _P, _P]),
    "cray_estimate_li": (C.c_int, [_P, C.c_int, C.c_uint64, _P, _P, _P, C.c_uint64, _P]),
    "cray_scene_create_multi": (C.c_int, [C.POINTER(SceneDesc), _P, C.c_int, C.c_uint32, _P]),
    "cray_render_multi": (C.c_int, [_P, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, _P, C.POINTER(RenderStats)]),
    "cray_abi_struct_sizes": (C.c_int, [_P, C.c_int]),
    "cray_write_exr": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _P]),
    "cray_render": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, _P, C.POINTER(RenderStats)]),
    "cray_render_device": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, _P, _P, C.POINTER(RenderStats)]),
    "cray_scene_get_info": (C.c_int, [_P, C.POINTER(SceneInfo)]),
    "cray_build_reference_bvh": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(C.POINTER(BvhNodeDump)), C.POINTER(C.c_uint64), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]),
    "cray_free": (None, [_P]),
    "cray_last_error": (C.c_char_p, []),
    "cray_version": (C.c_char_p, []),
    "cray_host_scene_num_warnings": (C.c_uint64, [_P]),
    "cray_host_scene_warning": (C.c_char_p, [_P, C.c_uint64]),
    "cray_register_standin_mesh": (None, [C.c_char_p, C.c_int, C.c_uint64, C.c_uint64]),
}
# host-side helpers that are not part of the reference-facing header
EXTRA_SIGNATURES = {
    "cray_set_image_decoder": (None, [IMAGE_DECODER]),
    "cray_clear_standin_meshes": (None, []),
    "cray_debug_decode_image": (C.c_int, [_P, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(_P)]),
    "cray_debug_transformation": (None, [C.c_int, _P, _P, _P]),
    "cray_debug_matrix_inverse": (C.c_int, [_P, _P]),
    "cray_debug_camera_matrices": (None, [C.POINTER(CameraDesc), _P]),
    "cray_debug_check_wide_bvh": (C.c_int, [C.POINTER(SceneDesc), _P]),
    "cray_debug_wide_stats": (C.c_int, [_P]),
    "cray_debug_find_contacts": (C.c_int, [C.POINTER(SceneDesc), _P, _P]),
    "cray_debug_build_reference_bvh_on": (C.c_int, [C.POINTER(SceneDesc), C.c_int, C.POINTER(C.POINTER(BvhNodeDump)), C.POINTER(C.c_uint64), C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]),
    "cray_debug_tokenize": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "cray_debug_parse_raw_value": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
}

_lib = None


def lib():
    """Load libcray_b200.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        path = os.environ.get("CRAY_B200_LIB", LIB_PATH)  # tuning builds (make VARIANT=...) sit beside the default library
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build the CUDA library first (__graft_entry__.build() or make -C craytracer_b200/csrc)")
        handle = C.CDLL(path)
        for table in (SIGNATURES, EXTRA_SIGNATURES):
            for name, (restype, argtypes) in table.items():
                fn = getattr(handle, name)
                fn.restype = restype
                fn.argtypes = argtypes
        _lib = handle
    return _lib
