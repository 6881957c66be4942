"""INTEGRATION.md shows the Rust binding a maintainer of the reference would add.  No rustc in this image, so the structs it
declares are checked here the way `check_layout()` in that document would at run time: their repr(C) sizes, computed from the
declarations as written, against the library's own sizeof (cray_abi_struct_sizes), and every prototype of include/cray_b200.h
must have an `extern "C"` line."""
import ctypes as C
import os
import re

from craytracer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ["cray_sphere_desc", "cray_triangle_desc", "cray_disk_desc", "cray_primitive_desc", "cray_texture_desc", "cray_image_desc", "cray_material_desc",
         "cray_light_desc", "cray_camera_desc", "cray_scene_desc", "cray_ray", "cray_hit", "cray_surface", "cray_render_stats", "cray_scene_info",
         "cray_bvh_node_dump"]
SCALARS = {"u8": (1, 1), "u32": (4, 4), "i32": (4, 4), "c_int": (4, 4), "f32": (4, 4), "u64": (8, 8), "i64": (8, 8), "f64": (8, 8)}


def rust_structs(text):
    out = {}
    for m in re.finditer(r"pub struct (\w+)\s*\{(.*?)\}", text, re.S):
        fields = []
        body = re.sub(r"//[^\n]*", "", m.group(2))
        for f in re.finditer(r"pub (\w+):\s*([^,{}]+?)\s*(?:,|$)", body.strip() + ",", re.S):
            fields.append((f.group(1), " ".join(f.group(2).split())))
        # `[f64; 3]` contains no comma, but `[T; N]` was split at ';' nowhere: fields are separated by commas only
        out[m.group(1)] = fields
    return out


def layout(ty, structs):
    ty = ty.strip()
    if ty.startswith("*const") or ty.startswith("*mut"):
        return 8, 8
    m = re.fullmatch(r"\[(.+);\s*(\d+)\]", ty)
    if m:
        size, align = layout(m.group(1), structs)
        return size * int(m.group(2)), align
    if ty in SCALARS:
        return SCALARS[ty]
    offset, align = 0, 1
    for _, fty in structs[ty]:
        fs, fa = layout(fty, structs)
        offset = (offset + fa - 1) // fa * fa + fs
        align = max(align, fa)
    return (offset + align - 1) // align * align, align


def test_rust_struct_declarations_have_the_library_layout():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    structs = rust_structs(text)
    sizes = (C.c_uint32 * 32)()
    n = _abi.lib().cray_abi_struct_sizes(sizes, 32)
    assert n == len(ORDER)
    for name, want in zip(ORDER, sizes[:n]):
        assert name in structs, f"INTEGRATION.md does not declare {name}"
        got, _ = layout(name, structs)
        assert got == want, f"{name}: the Rust declaration is {got} bytes, the library's struct {want}"


def test_every_header_prototype_has_a_rust_declaration():
    header = open(os.path.join(ROOT, "include", "cray_b200.h")).read()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"^(?:int|void|uint64_t|const [\w ]+\*?)\s*\*?\s*(cray_\w+)\(", header, re.M)))
    assert len(names) >= 25
    missing = [n for n in names if not re.search(r"pub fn " + n + r"\(", text)]
    assert not missing, missing
