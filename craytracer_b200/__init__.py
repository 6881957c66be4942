"""craytracer_b200 -- host-side mirror of the reference's scene / integrator seams over the sm_100a library.

Names follow the reference (banga/craytracer): ``parse_scene`` (src/scene_parser.rs:1078), ``Scene.intersect`` /
``Scene.intersects`` (src/scene.rs:55,:59), ``estimate_Li`` (src/path_integrator.rs:41), ``render``
(src/bin/craytracer.rs:224).  Everything numeric happens in ``libcray_b200.so`` (CUDA, sm_100a); this module only
marshals buffers.  There is no CPU fallback.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _abi
from ._abi import (BUILD_EXACT, BUILD_F32, BUILD_FAST, CRAY_NO_HIT, HIT_DTYPE, RAY_DTYPE, SURFACE_DTYPE, TRAVERSE_EXACT, TRAVERSE_F32, TRAVERSE_FAST, RenderStats, SceneDesc,
                   SceneInfo)

__all__ = ["ParserError", "CrayError", "HostScene", "Scene", "parse_scene", "load_scene", "make_rays", "tokenize", "parse_raw_value",
           "register_standin_mesh", "write_exr", "render_multi", "TRAVERSE_EXACT", "TRAVERSE_FAST", "TRAVERSE_F32", "BUILD_EXACT", "BUILD_FAST", "BUILD_F32", "CRAY_NO_HIT"]


class CrayError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


class ParserError(CrayError):
    """ParserError { message, location: Option<Location { line, column }> } (src/scene_parser.rs:72-90)."""

    def __init__(self, message, location):
        CrayError.__init__(self, _abi.CRAY_E_PARSE, f"{message}" + (f" at {location[0]}:{location[1]}" if location else ""))
        self.message = message
        self.location = location


def _check(rc):
    if rc == _abi.CRAY_OK:
        return
    L = _abi.lib()
    msg = (L.cray_last_error() or b"").decode("utf-8", "replace")
    if rc == _abi.CRAY_E_PARSE:
        line, col = C.c_uint32(0), C.c_uint32(0)
        L.cray_last_error_location(C.byref(line), C.byref(col))
        # the reference reports "No lights in the scene." with Location { 0, 0 }; an absent location is (0, 0) with no line info either
        loc = (line.value, col.value) if (line.value or col.value or msg == "No lights in the scene.") else None
        raise ParserError(msg, loc)
    raise CrayError(rc, msg)


_decoder_ref = None


def _install_pil_decoder():
    """Texture files other than binary PPM are decoded by PIL (the C++ host has no JPEG/PNG decoder)."""
    global _decoder_ref
    if _decoder_ref is not None:
        return
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def decode(path, w, h, out):
        try:
            from PIL import Image
            img = Image.open(path.decode()).convert("RGB")  # DynamicImage::to_rgb8 (src/texture.rs:57-59)
            data = img.tobytes()
            buf = libc.malloc(len(data))
            C.memmove(buf, data, len(data))
            w[0], h[0] = img.size
            out[0] = C.cast(buf, C.POINTER(C.c_uint8))
            return 0
        except Exception:
            return 1

    _decoder_ref = _abi.IMAGE_DECODER(decode)
    _abi.lib().cray_set_image_decoder(_decoder_ref)


def register_standin_mesh(file_name, kind, triangles, seed=0):
    """Procedural stand-in used when `file_name` (as written in a .cry file) does not exist.  kind 0: dragon, 1: interior."""
    _abi.lib().cray_register_standin_mesh(file_name.encode(), int(kind), int(triangles), int(seed))


class HostScene:
    """Owner of a flat ``cray_scene_desc`` produced by the C++ scene reader (parse_scene + load_obj)."""

    def __init__(self, handle):
        self._h = handle
        self.desc_ptr = _abi.lib().cray_host_scene_desc(handle)
        self.desc = self.desc_ptr.contents

    @property
    def warnings(self):
        L = _abi.lib()
        return [L.cray_host_scene_warning(self._h, i).decode() for i in range(L.cray_host_scene_num_warnings(self._h))]

    def close(self):
        if self._h:
            _abi.lib().cray_host_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_scene(text, base_dir="."):
    """parse_scene(input: &str) -> Result<Scene, ParserError>; mesh paths resolve against ``base_dir``."""
    _install_pil_decoder()
    h = C.c_void_p()
    _check(_abi.lib().cray_host_scene_parse(text.encode("utf-8"), os.fspath(base_dir).encode(), C.byref(h)))
    return HostScene(h)


def load_scene(path, base_dir=None):
    with open(path, "r", encoding="utf-8") as f:
        return parse_scene(f.read(), base_dir if base_dir is not None else os.getcwd())


def _json_call(fn, text):
    out = C.c_void_p()
    _check(fn(text.encode("utf-8"), C.byref(out)))
    try:
        return json.loads(C.string_at(out).decode("utf-8"))
    finally:
        _abi.lib().cray_free(out)


def tokenize(text):
    """tokenize(input) of src/scene_parser.rs:174 as a list of {kind, value?, line, column}."""
    return _json_call(_abi.lib().cray_debug_tokenize, text)


def parse_raw_value(text):
    """RawValue::from_tokens(tokenize(input)) of src/scene_parser.rs:336."""
    return _json_call(_abi.lib().cray_debug_parse_raw_value, text)


def render_multi(scenes_, seed=0, sample_begin=0, sample_end=None, mode=TRAVERSE_FAST):
    """render() on several GPUs of this process: one Scene per device (same description), sample slices per GPU, one NCCL
    reduce of the films (cray_render_multi).  Returns the film SUM (H, W, 3) f32 and the aggregated statistics."""
    first = scenes_[0]
    if sample_end is None:
        sample_end = first.num_samples
    film = np.empty((first.height, first.width, 3), dtype=np.float32)
    handles = (C.c_void_p * len(scenes_))(*[sc._h for sc in scenes_])
    stats = RenderStats()
    _check(_abi.lib().cray_render_multi(handles, len(scenes_), mode, seed, sample_begin, sample_end, film.ctypes.data, C.byref(stats)))
    return film, stats


def write_exr(path, film):
    """The EXR save of src/bin/craytracer.rs:367-369: (H, W, 3) f32 linear RGB -> 32-bit float OpenEXR."""
    film = np.ascontiguousarray(film, dtype=np.float32)
    h, w, ch = film.shape
    assert ch == 3
    _check(_abi.lib().cray_write_exr(os.fspath(path).encode(), w, h, film.ctypes.data))


def make_rays(origins, directions, max_distance=np.inf):
    """Pack (n,3) origins / directions (+ scalar or (n,) max_distance) into the 56-byte Ray records of src/ray.rs:7-11."""
    o = np.asarray(origins, dtype=np.float64).reshape(-1, 3)
    d = np.asarray(directions, dtype=np.float64).reshape(-1, 3)
    rays = np.empty(len(o), dtype=RAY_DTYPE)
    rays["origin"] = o
    rays["direction"] = d
    rays["max_distance"] = max_distance
    return rays


class Scene:
    """GPU-resident scene: Scene::new (src/scene.rs:25) + the S1/S2/S3 entry points."""

    def __init__(self, host_scene, device=0, build=BUILD_EXACT | BUILD_FAST, _handle=None):
        desc_ptr = host_scene.desc_ptr if isinstance(host_scene, HostScene) else host_scene
        self._keep = host_scene
        if _handle is None:
            h = C.c_void_p()
            _check(_abi.lib().cray_scene_create(desc_ptr, int(device), int(build), C.byref(h)))
        else:
            h = _handle
        self._h = h
        info = SceneInfo()
        _check(_abi.lib().cray_scene_get_info(self._h, C.byref(info)))
        self.info = info
        self.width, self.height = info.width, info.height
        self.max_depth, self.num_samples = info.max_depth, info.num_samples
        self.device = device

    @classmethod
    def create_multi(cls, host_scene, devices, build=BUILD_EXACT | BUILD_FAST):
        """One Scene per device from a single host-side build (cray_scene_create_multi); feed the list to render_multi."""
        desc_ptr = host_scene.desc_ptr if isinstance(host_scene, HostScene) else host_scene
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        handles = (C.c_void_p * len(devices))()
        _check(_abi.lib().cray_scene_create_multi(desc_ptr, devs, len(devices), int(build), handles))
        return [cls(host_scene, device=int(d), build=build, _handle=C.c_void_p(handles[k])) for k, d in enumerate(devices)]

    def close(self):
        if getattr(self, "_h", None):
            _abi.lib().cray_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def film_bounds(self):  # Scene::film_bounds src/scene.rs:63
        return self.width, self.height

    # -- S3 -------------------------------------------------------------------------------------------------------
    def intersect(self, rays, mode=TRAVERSE_FAST, surface=False):
        """Scene::intersect on a batch of rays.  Returns HIT_DTYPE records (prim == CRAY_NO_HIT on a miss)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        surf = np.empty(len(rays), dtype=SURFACE_DTYPE) if surface else None
        _check(_abi.lib().cray_trace_closest(self._h, mode, rays.ctypes.data, len(rays), hits.ctypes.data, surf.ctypes.data if surface else None))
        return (hits, surf) if surface else hits

    def intersects(self, rays, mode=TRAVERSE_FAST):
        """Scene::intersects on a batch of rays -> bool array."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.empty(len(rays), dtype=np.uint8)
        _check(_abi.lib().cray_trace_any(self._h, mode, rays.ctypes.data, len(rays), occ.ctypes.data))
        return occ.astype(bool)

    def intersect_device(self, d_rays, n, d_hits, d_surf=0, mode=TRAVERSE_FAST, stream=0):
        _check(_abi.lib().cray_trace_closest_device(self._h, mode, d_rays, n, d_hits, d_surf or None, stream or None))

    def intersects_device(self, d_rays, n, d_occluded, mode=TRAVERSE_FAST, stream=0):
        _check(_abi.lib().cray_trace_any_device(self._h, mode, d_rays, n, d_occluded, stream or None))

    # -- S2 -------------------------------------------------------------------------------------------------------
    def estimate_Li(self, x, y, sample_index, seed=0, mode=TRAVERSE_FAST):
        """render_pixel + estimate_Li for each (x, y, sample_index) with SobolSampler::new(seed, _).  Returns (n,3) f64."""
        x = np.ascontiguousarray(x, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint32)
        s = np.ascontiguousarray(sample_index, dtype=np.uint32)
        rgb = np.empty((len(x), 3), dtype=np.float64)
        _check(_abi.lib().cray_estimate_li(self._h, mode, seed, x.ctypes.data, y.ctypes.data, s.ctypes.data, len(x), rgb.ctypes.data))
        return rgb

    # -- S1 -------------------------------------------------------------------------------------------------------
    def render(self, seed=0, sample_begin=0, sample_end=None, mode=TRAVERSE_FAST):
        """render(): film SUM (H, W, 3) f32 over samples [sample_begin, sample_end) and the run's statistics."""
        if sample_end is None:
            sample_end = self.num_samples
        film = np.empty((self.height, self.width, 3), dtype=np.float32)
        stats = RenderStats()
        _check(_abi.lib().cray_render(self._h, mode, seed, sample_begin, sample_end, film.ctypes.data, C.byref(stats)))
        return film, stats

    def render_into(self, film, seed=0, sample_begin=0, sample_end=None, mode=TRAVERSE_FAST):
        """render() into a caller-owned host array of W*H*3 f32 (e.g. pinned memory); returns the statistics."""
        if sample_end is None:
            sample_end = self.num_samples
        assert film.dtype == np.float32 and film.size == self.height * self.width * 3 and film.flags["C_CONTIGUOUS"]
        stats = RenderStats()
        _check(_abi.lib().cray_render(self._h, mode, seed, sample_begin, sample_end, film.ctypes.data, C.byref(stats)))
        return stats

    def render_device(self, d_film, seed=0, sample_begin=0, sample_end=None, mode=TRAVERSE_FAST, stream=0):
        """Same, into a device buffer of W*H*3 f32 (e.g. a torch CUDA tensor's data_ptr())."""
        if sample_end is None:
            sample_end = self.num_samples
        stats = RenderStats()
        _check(_abi.lib().cray_render_device(self._h, mode, seed, sample_begin, sample_end, d_film, stream or None, C.byref(stats)))
        return stats
