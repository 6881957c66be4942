// Host-only entry points of include/cray_b200.h: scene ingest (S0) and the parser introspection used by the
// host-logic tests that restate the reference's tests/test_parser.rs.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/cray_b200.h"
#include "cry_parser.hpp"
#include "host_scene.hpp"
#include "image_decode.hpp"
#include "host_math.hpp"
#include "bvh_build.hpp"

namespace cray {
void set_error(const std::string& msg);
}

namespace {
thread_local uint32_t g_line = 0, g_column = 0;

int fail_parse(const cray::ParserError& e) {
    cray::set_error(e.message);
    g_line = e.has_location ? e.location.line : 0;
    g_column = e.has_location ? e.location.column : 0;
    return CRAY_E_PARSE;
}
char* dup_string(const std::string& s) {
    char* p = (char*)std::malloc(s.size() + 1);
    std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
}  // namespace

struct cray_host_scene {
    cray::HostScene* scene;
};

extern "C" {

void cray_last_error_location(uint32_t* line, uint32_t* column) {
    if (line) *line = g_line;
    if (column) *column = g_column;
}

int cray_host_scene_parse(const char* cry_text, const char* base_dir, cray_host_scene** out) {
    if (!cry_text || !out) { cray::set_error("bad arguments"); return CRAY_E_INVALID; }
    *out = nullptr;
    g_line = g_column = 0;
    try {
        cray::HostScene* hs = cray::build_host_scene(cry_text, base_dir ? base_dir : "");
        *out = new cray_host_scene{hs};
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    } catch (const cray::IoError& e) {
        cray::set_error(e.message);
        return CRAY_E_IO;
    } catch (const cray::UnsupportedError& e) {
        cray::set_error(e.message);
        return CRAY_E_UNSUPPORTED;
    } catch (const std::exception& e) {
        cray::set_error(e.what());
        return CRAY_E_INVALID;
    }
}

int cray_host_scene_load(const char* cry_path, const char* base_dir, cray_host_scene** out) {
    if (!cry_path || !out) { cray::set_error("bad arguments"); return CRAY_E_INVALID; }
    std::ifstream f(cry_path, std::ios::binary);
    if (!f) { cray::set_error(std::string("Error reading scene file ") + cry_path); return CRAY_E_IO; }
    std::stringstream ss;
    ss << f.rdbuf();
    return cray_host_scene_parse(ss.str().c_str(), base_dir, out);
}

// sizeof of the ABI structs of include/cray_b200.h, in declaration order: lets a binding (the ctypes mirror, a Rust `repr(C)`
// block) verify its layout at start-up.  Returns the number of entries written (at most `capacity`).
int cray_abi_struct_sizes(uint32_t* out, int capacity) {
    const uint32_t sizes[] = {sizeof(cray_sphere_desc), sizeof(cray_triangle_desc), sizeof(cray_disk_desc), sizeof(cray_primitive_desc),
                              sizeof(cray_texture_desc), sizeof(cray_image_desc), sizeof(cray_material_desc), sizeof(cray_light_desc),
                              sizeof(cray_camera_desc), sizeof(cray_scene_desc), sizeof(cray_ray), sizeof(cray_hit), sizeof(cray_surface),
                              sizeof(cray_render_stats), sizeof(cray_scene_info), sizeof(cray_bvh_node_dump)};
    const int n = (int)(sizeof(sizes) / sizeof(sizes[0]));
    for (int i = 0; i < n && i < capacity; ++i) out[i] = sizes[i];
    return n < capacity ? n : capacity;
}

// OpenEXR 2.0 single-part scan-line file, channels B, G, R as 32-bit float, NO_COMPRESSION, increasing-y line order.
int cray_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgb) {
    if (!path || !rgb || width == 0 || height == 0) { cray::set_error("bad arguments"); return CRAY_E_INVALID; }
    std::string head;
    auto put = [&](const void* p, size_t n) { head.append(static_cast<const char*>(p), n); };
    auto put_i32 = [&](int32_t v) { put(&v, 4); };
    auto put_f32 = [&](float v) { put(&v, 4); };
    auto put_str = [&](const char* z) { head.append(z); head.push_back('\0'); };
    auto attr = [&](const char* name, const char* type, int32_t size) { put_str(name); put_str(type); put_i32(size); };
    put_i32(20000630);  // magic
    put_i32(2);         // version 2, no flags: single-part scan-line
    attr("channels", "chlist", 3 * (2 + 4 + 4 + 4 + 4) + 1);
    for (const char* ch : {"B", "G", "R"}) {
        put_str(ch);
        put_i32(2);  // FLOAT
        put_i32(0);  // pLinear + 3 reserved bytes
        put_i32(1); put_i32(1);  // x / y sampling
    }
    head.push_back('\0');
    attr("compression", "compression", 1); head.push_back('\0');
    attr("dataWindow", "box2i", 16); put_i32(0); put_i32(0); put_i32((int32_t)width - 1); put_i32((int32_t)height - 1);
    attr("displayWindow", "box2i", 16); put_i32(0); put_i32(0); put_i32((int32_t)width - 1); put_i32((int32_t)height - 1);
    attr("lineOrder", "lineOrder", 1); head.push_back('\0');
    attr("pixelAspectRatio", "float", 4); put_f32(1.0f);
    attr("screenWindowCenter", "v2f", 8); put_f32(0.0f); put_f32(0.0f);
    attr("screenWindowWidth", "float", 4); put_f32(1.0f);
    head.push_back('\0');  // end of header
    const uint64_t row_bytes = (uint64_t)width * 3 * 4, chunk = 8 + row_bytes;
    const uint64_t first = head.size() + 8ull * height;
    std::ofstream f(path, std::ios::binary);
    if (!f) { cray::set_error(std::string("cannot open ") + path + " for writing"); return CRAY_E_IO; }
    f.write(head.data(), (std::streamsize)head.size());
    for (uint32_t y = 0; y < height; ++y) {
        const uint64_t off = first + chunk * y;
        f.write(reinterpret_cast<const char*>(&off), 8);
    }
    std::vector<float> line((size_t)width * 3);
    for (uint32_t y = 0; y < height; ++y) {
        const int32_t yy = (int32_t)y, size = (int32_t)row_bytes;
        f.write(reinterpret_cast<const char*>(&yy), 4);
        f.write(reinterpret_cast<const char*>(&size), 4);
        const float* src = rgb + (size_t)y * width * 3;
        for (uint32_t x = 0; x < width; ++x) {  // channel-planar within a scan line, channels in alphabetical order
            line[x] = src[3 * x + 2];
            line[width + x] = src[3 * x + 1];
            line[2 * (size_t)width + x] = src[3 * x];
        }
        f.write(reinterpret_cast<const char*>(line.data()), (std::streamsize)row_bytes);
    }
    if (!f) { cray::set_error(std::string("error writing ") + path); return CRAY_E_IO; }
    return CRAY_OK;
}

const cray_scene_desc* cray_host_scene_desc(const cray_host_scene* hs) { return hs ? &hs->scene->desc : nullptr; }

void cray_host_scene_destroy(cray_host_scene* hs) {
    if (!hs) return;
    delete hs->scene;
    delete hs;
}

// number of warnings / i-th warning (unused keys, stand-in meshes, MTL problems)
uint64_t cray_host_scene_num_warnings(const cray_host_scene* hs) { return hs ? hs->scene->warnings.size() : 0; }
const char* cray_host_scene_warning(const cray_host_scene* hs, uint64_t i) { return hs && i < hs->scene->warnings.size() ? hs->scene->warnings[i].c_str() : ""; }

void cray_set_image_decoder(cray::image_decoder_fn fn) { cray::set_image_decoder(fn); }

// The built-in texture decoder on its own (tests/test_image_decode.py): RGB8 rows in a buffer to release with cray_free.
int cray_debug_decode_image(const uint8_t* bytes, uint64_t n, uint32_t* width, uint32_t* height, uint8_t** rgb) {
    std::vector<uint8_t> out;
    std::string why;
    if (!bytes || !width || !height || !rgb) { cray::set_error("null argument"); return CRAY_E_INVALID; }
    if (!cray::decode_image(bytes, (size_t)n, *width, *height, out, why)) { cray::set_error(why); return CRAY_E_UNSUPPORTED; }
    *rgb = static_cast<uint8_t*>(std::malloc(out.size() ? out.size() : 1));
    if (!*rgb) { cray::set_error("out of memory"); return CRAY_E_INVALID; }
    std::memcpy(*rgb, out.data(), out.size());
    return CRAY_OK;
}
void cray_register_standin_mesh(const char* file_name, int kind, uint64_t triangles, uint64_t seed) { cray::register_standin_mesh(file_name, kind, triangles, seed); }
void cray_clear_standin_meshes(void) { cray::clear_standin_meshes(); }

// Host math introspection for the parity tests: the matrices the scene builder derives, in the oracle's layout.
// kind: 0 translate 1 scale 2 rotate_x 3 rotate_y 4 rotate_z 5 look_at 6 perspective 7 orthographic
void cray_debug_transformation(int kind, const double* p, double* matrix16, double* inverse16) {
    using namespace cray;
    Xform t;
    switch (kind) {
        case 0: t = xf_translate(p[0], p[1], p[2]); break;
        case 1: t = xf_scale(p[0], p[1], p[2]); break;
        case 2: t = xf_rotate(0, p[0]); break;
        case 3: t = xf_rotate(1, p[0]); break;
        case 4: t = xf_rotate(2, p[0]); break;
        case 5: t = xf_look_at(mk(p[0], p[1], p[2]), mk(p[3], p[4], p[5]), mk(p[6], p[7], p[8])); break;
        case 6: t = xf_perspective(p[0], p[1], p[2]); break;
        default: t = xf_orthographic(p[0], p[1]); break;
    }
    std::memcpy(matrix16, t.fwd.m, 16 * sizeof(double));
    std::memcpy(inverse16, t.inv.m, 16 * sizeof(double));
}
int cray_debug_matrix_inverse(const double* a16, double* out16) {
    cray::Mat4 a, r;
    std::memcpy(a.m, a16, sizeof(a.m));
    if (!cray::invert(a, r)) return 0;
    std::memcpy(out16, r.m, sizeof(r.m));
    return 1;
}
// camera_from_raster then world_from_camera (4x4 row-major each) as uploaded to the GPU
void cray_debug_camera_matrices(const cray_camera_desc* c, double* out32) {
    using namespace cray;
    const Xform wfc = xf_look_at(mk(c->origin[0], c->origin[1], c->origin[2]), mk(c->target[0], c->target[1], c->target[2]), mk(c->up[0], c->up[1], c->up[2]));
    const Xform sfc = c->kind == CRAY_CAMERA_PERSPECTIVE ? xf_perspective(c->fov, 1e-2, 1000.0) : xf_orthographic(0.0, 1.0);
    const Xform cfr = camera_from_raster(sfc, c->width);
    std::memcpy(out32, cfr.fwd.m, 16 * sizeof(double));
    std::memcpy(out32 + 16, wfc.fwd.m, 16 * sizeof(double));
}

// The planar-contact analysis of bvh_build.hpp on the host only (no device): out[0] = marked node boxes, out[1] = marked primitives,
// out[2] = binary nodes; prim_flags (optional, n_primitives bytes) receives the per-primitive marks.
int cray_debug_find_contacts(const cray_scene_desc* desc, uint64_t* out3, uint8_t* prim_flags) {
    using namespace cray;
    if (!desc || !out3) { set_error("bad arguments"); return CRAY_E_INVALID; }
    RefBvh ref;
    build_reference_bvh(*desc, ref);
    if (!ref.error.empty()) { set_error(ref.error); return CRAY_E_BVH; }
    ContactInfo info;
    find_contacts(*desc, ref, info);
    out3[0] = info.n_nodes; out3[1] = info.n_prims; out3[2] = ref.nodes.size();
    if (prim_flags) std::memcpy(prim_flags, info.prim_flag.data(), info.prim_flag.size());
    return CRAY_OK;
}

// Structural check of the 8-wide BVH against the reference binary tree it was collapsed from (host only):
// every primitive exactly once, every quantised child box encloses the exact f64 box of the subtree (or the single
// primitive) it stands for.  out[0..3] = wide nodes, depth, interior children, leaf children.
int cray_debug_check_wide_bvh(const cray_scene_desc* desc, uint64_t* out4) {
    using namespace cray;
    RefBvh ref;
    build_reference_bvh(*desc, ref);
    if (!ref.error.empty()) { set_error(ref.error); return CRAY_E_BVH; }
    WideBvh wide;
    collapse_to_wide(*desc, ref, wide);
    std::vector<uint32_t> seen(desc->n_primitives, 0);
    for (uint32_t p : wide.prim_order) {
        if (p >= desc->n_primitives) { set_error("wide order references a missing primitive"); return CRAY_E_INVALID; }
        seen[p] += 1;
    }
    for (uint32_t c : seen)
        if (c != 1) { set_error("a primitive does not appear exactly once in the wide leaf order"); return CRAY_E_INVALID; }
    // reference leaves as (first primitive -> (count, box)) for matching leaf children
    uint64_t interior = 0, leaves = 0;
    // recompute exact boxes bottom-up from the wide structure and compare with the quantised ones
    struct Rec { Box3 box; };
    std::vector<Box3> node_box(wide.nodes.size());
    std::vector<char> done(wide.nodes.size(), 0);
    // children always have larger indices than their parent, so a reverse sweep sees children first
    for (size_t ni = wide.nodes.size(); ni-- > 0;) {
        const WideNode& w = wide.nodes[ni];
        const double p[3] = {w.px, w.py, w.pz};
        const double sc[3] = {std::ldexp(1.0, (int)w.ex - 127), std::ldexp(1.0, (int)w.ey - 127), std::ldexp(1.0, (int)w.ez - 127)};
        bool any = false;
        Box3 acc{};
        uint32_t child = w.child_base, prim = w.prim_base;
        if (w.imask & w.leafmask) { set_error("slot marked both interior and leaf"); return CRAY_E_INVALID; }
        for (int s = 0; s < 8; ++s) {
            const bool is_interior = (w.imask >> s) & 1, is_leaf = (w.leafmask >> s) & 1;
            if (!is_interior && !is_leaf) continue;
            Box3 exact{};
            if (is_interior) {
                if (child <= ni || child >= wide.nodes.size() || !done[child]) { set_error("bad child index"); return CRAY_E_INVALID; }
                exact = node_box[child];
                child += 1;
                interior += 1;
            } else {
                if (prim >= wide.prim_order.size()) { set_error("bad primitive index"); return CRAY_E_INVALID; }
                exact = primitive_bounds(*desc, wide.prim_order[prim]);
                prim += 1;
                leaves += 1;
            }
            const double qlo[3] = {p[0] + w.qlo[0][s] * sc[0], p[1] + w.qlo[1][s] * sc[1], p[2] + w.qlo[2][s] * sc[2]};
            const double qhi[3] = {p[0] + w.qhi[0][s] * sc[0], p[1] + w.qhi[1][s] * sc[1], p[2] + w.qhi[2][s] * sc[2]};
            for (int ax = 0; ax < 3; ++ax)
                if (qlo[ax] > exact.lo[ax] || qhi[ax] < exact.hi[ax]) { set_error("quantised child box does not enclose the exact box"); return CRAY_E_INVALID; }
            acc = any ? box_union(acc, exact) : exact;
            any = true;
        }
        if (!any) { set_error("empty wide node"); return CRAY_E_INVALID; }
        node_box[ni] = acc;
        done[ni] = 1;
    }
    for (int ax = 0; ax < 3; ++ax)
        if (node_box[0].lo[ax] != ref.bounds.lo[ax] || node_box[0].hi[ax] != ref.bounds.hi[ax]) { set_error("root box differs from the scene bounds"); return CRAY_E_INVALID; }
    out4[0] = wide.nodes.size(); out4[1] = wide.depth; out4[2] = interior; out4[3] = leaves;
    return CRAY_OK;
}

// Parser introspection: JSON of tokenize(input) / RawValue::from_tokens(tokenize(input)); caller frees with cray_free.
int cray_debug_tokenize(const char* input, char** json_out) {
    g_line = g_column = 0;
    try {
        *json_out = dup_string(cray::tokens_to_json(cray::tokenize(input)));
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    }
}
int cray_debug_parse_raw_value(const char* input, char** json_out) {
    g_line = g_column = 0;
    try {
        auto tokens = cray::tokenize(input);
        size_t pos = 0;
        auto v = cray::parse_raw_value(tokens, pos);
        *json_out = dup_string(cray::raw_value_to_json(*v));
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    }
}

}  // extern "C"
