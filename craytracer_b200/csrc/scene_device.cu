// cray_scene_create / destroy: flatten a cray_scene_desc into the HBM layout of device_types.cuh.
// Replaces Scene::new (src/scene.rs:25-53): BVH build + LightSampler::new, then one upload.
#include "scene_device.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>

#include "host_math.hpp"
#include "sampler.cuh"
#include "../../include/cray_sobol_directions.h"

namespace cray {

namespace {
thread_local std::string g_error;
}
void set_error(const std::string& msg) { g_error = msg; }
const std::string& last_error() { return g_error; }
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
    return CRAY_E_CUDA;
}

namespace {

V3 v3(const double* p) { return mk(p[0], p[1], p[2]); }

DiskXf make_disk(const cray_disk_desc& k) {  // Shape::new_disk shape.rs:133-153
    const Xform o2w = xf_translate(k.origin[0], k.origin[1], k.origin[2]) * xf_rotate(0, to_radians(k.rotate_x)) * xf_rotate(1, to_radians(k.rotate_y));
    DiskXf d;
    d.o2w = affine_of(o2w.fwd);
    d.w2o = affine_of(o2w.inv);
    d.radius = k.radius;
    d.inner_radius = k.inner_radius;
    return d;
}

LeafPrim make_leaf_prim(const cray_scene_desc& d, uint32_t prim) {
    const cray_primitive_desc& p = d.primitives[prim];
    LeafPrim lp{};
    lp.prim = prim;
    switch (p.shape_kind) {
        case CRAY_SHAPE_TRIANGLE: {
            const cray_triangle_desc& t = d.triangles[p.shape_index];
            std::memcpy(lp.d, t.v0, 24);
            std::memcpy(lp.d + 3, t.e1, 24);
            std::memcpy(lp.d + 6, t.e2, 24);
            lp.kind = PRIM_TRIANGLE;
            {   // flat: every component of n01 and n02 is +0.0 bit for bit, so n0 + n01 * u + n02 * v == n0 + (+0.0) + (+0.0)
                const double z[6] = {t.n01[0], t.n01[1], t.n01[2], t.n02[0], t.n02[1], t.n02[2]};
                const uint64_t zero_bits[6] = {0, 0, 0, 0, 0, 0};
                if (std::memcmp(z, zero_bits, sizeof(z)) == 0) lp.kind |= kKindFlatTriangle;
            }
            break;
        }
        case CRAY_SHAPE_SPHERE: {
            const cray_sphere_desc& s = d.spheres[p.shape_index];
            std::memcpy(lp.d, s.origin, 24);
            lp.d[3] = s.radius;
            lp.kind = PRIM_SPHERE;
            break;
        }
        default:
            lp.kind = PRIM_DISK | (p.shape_index << 16);
            break;
    }
    // shade class: which family of shading code a hit on this primitive runs (area lights carry a black matte, primitive.rs:43-46)
    const uint32_t shade_class = p.area_light >= 0 ? (uint32_t)CRAY_MAT_MATTE : d.materials[p.material].kind;
    lp.kind |= (shade_class & 0xFFu) << 8;
    return lp;
}

// F32 mode record of the same leaf slot: v0, v1 = v0 + e1, v2 = v0 + e2 rounded to f32 (the sums reproduce the mesh's own
// vertices to within an f64 ulp, so the f32 values of a shared vertex agree between its triangles)
Tri32 make_tri32(const LeafPrim& lp) {
    Tri32 t{};
    t.prim = lp.prim;
    t.kind = lp.kind;
    if ((lp.kind & 0xFFu) == PRIM_TRIANGLE)
        for (int a = 0; a < 3; ++a) {
            t.v0[a] = (float)lp.d[a];
            t.v1[a] = (float)(lp.d[a] + lp.d[3 + a]);
            t.v2[a] = (float)(lp.d[a] + lp.d[6 + a]);
        }
    return t;
}

double shape_area(const cray_scene_desc& d, const cray_primitive_desc& p) {  // Shape::area shape.rs:504-514
    switch (p.shape_kind) {
        case CRAY_SHAPE_SPHERE: { const double r = d.spheres[p.shape_index].radius; return kPi * (r * r); }
        case CRAY_SHAPE_TRIANGLE: { const cray_triangle_desc& t = d.triangles[p.shape_index]; return magnitude(cross(v3(t.e1), v3(t.e2))) / 2.0; }
        default: { const cray_disk_desc& k = d.disks[p.shape_index]; return kPi * (k.radius * k.radius - k.inner_radius * k.inner_radius); }
    }
}

bool tex_is_black(const cray_texture_desc& t) {  // texture.rs:82-90
    auto black = [](const double* c) { return c[0] == 0.0 && c[1] == 0.0 && c[2] == 0.0; };
    if (t.kind == CRAY_TEX_CONSTANT) return black(t.a);
    if (t.kind == CRAY_TEX_CHECKERBOARD) return black(t.a) && black(t.b);
    return false;
}
bool tex_is_zero(const cray_texture_desc& t) {  // texture.rs:92-100
    if (t.kind == CRAY_TEX_CONSTANT) return t.a[0] == 0.0;
    if (t.kind == CRAY_TEX_CHECKERBOARD) return t.a[0] == 0.0 && t.b[0] == 0.0;
    return false;
}

DevMaterial make_material(const cray_material_desc& m) {  // Material::new_* material.rs:20-70
    DevMaterial r{};
    auto lobe = [](uint32_t kind, const cray_texture_desc& t0, const cray_texture_desc& t1, const cray_texture_desc& sigma, double eta_i, double eta_t) {
        DevLobe l{};
        l.kind = kind; l.t0 = t0; l.t1 = t1; l.sigma = sigma; l.eta_i = eta_i; l.eta_t = eta_t;
        return l;
    };
    switch (m.kind) {
        case CRAY_MAT_MATTE:
            r.is_bsdf = 0; r.n_lobes = 1;
            r.lobes[0] = lobe(tex_is_zero(m.t2) ? LOBE_LAMBERTIAN : LOBE_OREN_NAYAR, m.t0, m.t0, m.t2, 1.0, 1.0);
            break;
        case CRAY_MAT_GLASS:
            r.is_bsdf = 0; r.n_lobes = 1;
            r.lobes[0] = lobe(LOBE_FRESNEL_SPECULAR, m.t0, m.t1, m.t2, 1.0, m.eta);
            break;
        case CRAY_MAT_PLASTIC:
            r.is_bsdf = 1; r.n_lobes = 0;
            if (!tex_is_black(m.t0)) r.lobes[r.n_lobes++] = lobe(!tex_is_zero(m.t2) ? LOBE_OREN_NAYAR : LOBE_LAMBERTIAN, m.t0, m.t0, m.t2, 1.0, 1.0);
            if (!tex_is_black(m.t1)) r.lobes[r.n_lobes++] = lobe(LOBE_SPECULAR_BRDF, m.t1, m.t1, m.t2, 1.0, 1.5);
            break;
        default:  // CRAY_MAT_METAL
            r.is_bsdf = 1; r.n_lobes = 1;
            r.lobes[0] = lobe(LOBE_CONDUCTOR, m.t0, m.t1, m.t2, 1.0, 1.0);
            break;
    }
    r.all_delta = 1;
    for (uint32_t i = 0; i < r.n_lobes; ++i)
        if (r.lobes[i].kind == LOBE_LAMBERTIAN || r.lobes[i].kind == LOBE_OREN_NAYAR) r.all_delta = 0;
    for (uint32_t i = 0; i < r.n_lobes; ++i)
        if (r.lobes[i].t0.kind != CRAY_TEX_CONSTANT || r.lobes[i].t1.kind != CRAY_TEX_CONSTANT || r.lobes[i].sigma.kind != CRAY_TEX_CONSTANT) r.needs_uv = 1;
    return r;
}

template <class T>
int upload(cray_scene* sc, const std::vector<T>& host, const T** dev) {
    *dev = nullptr;
    if (host.empty()) return CRAY_OK;
    void* p = nullptr;
    CRAY_CUDA(cudaMalloc(&p, host.size() * sizeof(T)));
    sc->allocations.push_back(p);
    CRAY_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T*>(p);
    return CRAY_OK;
}

double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

int validate(const cray_scene_desc* d) {
    if (!d) { set_error("null scene description"); return CRAY_E_INVALID; }
    if (d->n_primitives == 0) { set_error("scene has no primitives"); return CRAY_E_INVALID; }
    if (d->n_lights == 0) { set_error("No lights in the scene."); return CRAY_E_INVALID; }  // scene_parser.rs:1103
    if (d->n_primitives >= 0xFFFFFFF0ull) { set_error("too many primitives"); return CRAY_E_INVALID; }
    if (d->n_disks > 0x7FFFull) { set_error("more than 32767 disks"); return CRAY_E_UNSUPPORTED; }
    // The sampler has 256 dimensions (sobol_burley panics beyond them) and a path vertex at bounce b draws dimensions 4 + 8 b .. 11 + 8 b
    // (path_integrator.rs:25-36): depths beyond 31 have no sample values to draw.  The wavefront's path state also packs the
    // bounce count into 8 bits.
    if (d->max_depth > 31) { set_error("max_depth > 31: the Sobol sampler has 256 dimensions, 8 per bounce (the reference panics there)"); return CRAY_E_UNSUPPORTED; }
    if (d->camera.width == 0 || d->camera.height == 0 || d->camera.width > 65535 || d->camera.height > 65535) { set_error("film size out of range"); return CRAY_E_INVALID; }
    for (uint64_t i = 0; i < d->n_primitives; ++i) {
        const cray_primitive_desc& p = d->primitives[i];
        const uint64_t lim = p.shape_kind == CRAY_SHAPE_SPHERE ? d->n_spheres : (p.shape_kind == CRAY_SHAPE_TRIANGLE ? d->n_triangles : (p.shape_kind == CRAY_SHAPE_DISK ? d->n_disks : 0));
        if (p.shape_index >= lim) { set_error("primitive references a missing shape"); return CRAY_E_INVALID; }
        if (p.area_light >= 0) {
            if ((uint64_t)p.area_light >= d->n_lights || d->lights[p.area_light].kind != CRAY_LIGHT_AREA) { set_error("Non area light provided as area light for shape"); return CRAY_E_INVALID; }
        } else if (p.material < 0 || (uint64_t)p.material >= d->n_materials) { set_error("primitive references a missing material"); return CRAY_E_INVALID; }
    }
    for (uint64_t i = 0; i < d->n_lights; ++i)
        if (d->lights[i].kind == CRAY_LIGHT_AREA && (d->lights[i].primitive < 0 || (uint64_t)d->lights[i].primitive >= d->n_primitives)) { set_error("area light references a missing primitive"); return CRAY_E_INVALID; }
    for (uint64_t i = 0; i < d->n_materials; ++i) {
        const cray_texture_desc* ts[3] = {&d->materials[i].t0, &d->materials[i].t1, &d->materials[i].t2};
        for (auto* t : ts)
            if (t->kind == CRAY_TEX_IMAGE && (t->image < 0 || (uint64_t)t->image >= d->n_images)) { set_error("texture references a missing image"); return CRAY_E_INVALID; }
    }
    return CRAY_OK;
}

// Everything cray_scene_create derives from a description on the host: built once, uploaded to one or several devices.
struct HostBuild {
    uint32_t build_flags = 0;
    RefBvh ref;
    WideBvh wide;
    std::vector<LeafPrim> bin_prims, wide_prims;
    std::vector<Tri32> wide_tris32;
    std::vector<uint32_t> rank_of_prim, wide_slot_of_prim;
    ContactInfo contact;
    std::vector<DiskXf> disks;
    std::vector<TriShade> tri_shade;
    std::vector<DevMaterial> materials;
    std::vector<cray_primitive_desc> prims;
    std::vector<DevImage> images;
    std::vector<uint8_t> texels;
    std::vector<double> gamma_lut;
    std::vector<DevLight> lights;
    std::vector<double> cdf;
    std::vector<uint32_t> pixel_order, sobol;
    std::vector<cray_sphere_desc> spheres;
    DevCamera cam{};
    double build_ms = 0.0;
};

// Scene::new (scene.rs:25-53) on the host: the reference BVH, its wide collapse, and every flat array of device_types.cuh.
int build_host_side(const cray_scene_desc* d, uint32_t build_flags, HostBuild& hb, int device) {
    hb.build_flags = build_flags;
    RefBvh& ref = hb.ref;
    WideBvh& wide = hb.wide;
    auto t0 = std::chrono::steady_clock::now();
    // the per-primitive shading records do not depend on the trees: they are built beside them (the host cores are idle while the
    // device builds the reference tree)
    std::thread shading_records([&hb, d] {
        const size_t np = (size_t)d->n_primitives;
        // per-primitive shading records for triangles (flat ones are flagged in their LeafPrim, see make_leaf_prim)
        std::vector<TriShade>& tri_shade = hb.tri_shade;
        tri_shade.resize(d->n_triangles ? np : 0);
        {
            const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
            std::vector<std::thread> pool;
            auto work = [&](unsigned t0) {
                for (size_t i = t0; i < tri_shade.size(); i += nt) {
                    const cray_primitive_desc& p = d->primitives[i];
                    TriShade& ts = tri_shade[i];
                    std::memset(&ts, 0, sizeof(ts));
                    if (p.shape_kind != CRAY_SHAPE_TRIANGLE) continue;
                    const cray_triangle_desc& t = d->triangles[p.shape_index];
                    std::memcpy(ts.n0, t.n0, 24); std::memcpy(ts.n01, t.n01, 24); std::memcpy(ts.n02, t.n02, 24);
                    std::memcpy(ts.uv0, t.uv0, 16); std::memcpy(ts.uv01, t.uv01, 16); std::memcpy(ts.uv02, t.uv02, 16);
                    ts.material = p.area_light >= 0 ? (int32_t)d->n_materials : p.material;
                    ts.area_light = p.area_light;
                }
            };
            for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto& th : pool) th.join();
        }
    });
    struct JoinGuard { std::thread& th; ~JoinGuard() { if (th.joinable()) th.join(); } } shading_guard{shading_records};
    build_reference_bvh(*d, ref, 0, device);
    if (!ref.error.empty()) { set_error(ref.error); return CRAY_E_BVH; }
    if (build_flags & CRAY_BUILD_FAST) {
        // CRAY_WIDE_TREE=sah3 (tuning): collapse the wide BVH from a second tree built with three-axis SAH splits
        const char* tree = std::getenv("CRAY_WIDE_TREE");
        if (tree && std::string(tree) == "sah3") {
            RefBvh quality;
            build_quality_bvh(*d, quality);
            if (!quality.error.empty()) { set_error(quality.error); return CRAY_E_BVH; }
            collapse_to_wide(*d, quality, wide);
        } else {
            collapse_to_wide(*d, ref, wide);
        }
        if (wide.depth >= (uint32_t)kWideStackLimit) { set_error("wide BVH deeper than the traversal stack"); return CRAY_E_BVH; }
        if (d->n_primitives >= (1ull << 27)) { set_error("more than 2^27 primitives: the fast traversal's queue entries hold 27-bit leaf slots"); return CRAY_E_UNSUPPORTED; }
    }
    // planar contact (bvh_build.hpp): which boxes and primitives can produce the reference's false box misses
    if (build_flags & CRAY_BUILD_FAST) find_contacts(*d, ref, hb.contact);
    hb.build_ms = ms_since(t0);
    PhaseTimer timer;
    const size_t np = (size_t)d->n_primitives;
    // leaf-ordered intersection records
    std::vector<LeafPrim>& bin_prims = hb.bin_prims;
    std::vector<LeafPrim>& wide_prims = hb.wide_prims;
    std::vector<uint32_t>& rank_of_prim = hb.rank_of_prim;
    bin_prims.resize(np); wide_prims.resize(wide.prim_order.size()); rank_of_prim.resize(np);
    if (!wide_prims.empty()) hb.wide_slot_of_prim.resize(np);
    const std::vector<uint8_t>& contact_prim = hb.contact.prim_flag;
    std::vector<Tri32>& wide_tris32 = hb.wide_tris32;
    if (build_flags & CRAY_BUILD_F32) wide_tris32.resize(wide_prims.size());
    {
        const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        auto work = [&](unsigned t) {
            for (size_t i = t; i < np; i += nt) {
                bin_prims[i] = make_leaf_prim(*d, ref.prim_order[i]);
                rank_of_prim[ref.prim_order[i]] = (uint32_t)i;
                if (!wide_prims.empty()) {
                    const uint32_t prim = wide.prim_order[i];
                    wide_prims[i] = make_leaf_prim(*d, prim);
                    if (!contact_prim.empty() && contact_prim[prim]) wide_prims[i].kind |= kKindContact;
                    hb.wide_slot_of_prim[prim] = (uint32_t)i;
                }
                if (!wide_tris32.empty()) wide_tris32[i] = make_tri32(wide_prims[i]);
            }
        };
        for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    }
    timer.mark("leaf records");
    std::vector<DiskXf>& disks = hb.disks;
    disks.resize(d->n_disks);
    for (size_t i = 0; i < disks.size(); ++i) disks[i] = make_disk(d->disks[i]);
    shading_records.join();   // (built while the tree was)
    timer.mark("shading records");
    // materials (+ the black matte that area-light primitives carry, primitive.rs:43-46)
    std::vector<DevMaterial>& materials = hb.materials;
    materials.resize(d->n_materials + 1);
    for (size_t i = 0; i < d->n_materials; ++i) materials[i] = make_material(d->materials[i]);
    {
        cray_material_desc black{};
        black.kind = CRAY_MAT_MATTE;
        black.t0.kind = black.t1.kind = black.t2.kind = CRAY_TEX_CONSTANT;
        materials[d->n_materials] = make_material(black);
    }
    std::vector<cray_primitive_desc>& prims = hb.prims;
    prims.assign(d->primitives, d->primitives + np);
    for (auto& p : prims)
        if (p.area_light >= 0) p.material = (int32_t)d->n_materials;
    // images
    std::vector<DevImage>& images = hb.images;
    images.resize(d->n_images);
    std::vector<uint8_t>& texels = hb.texels;
    for (size_t i = 0; i < images.size(); ++i) {
        images[i] = {d->images[i].width, d->images[i].height, (uint64_t)texels.size()};
        const size_t bytes = (size_t)d->images[i].width * d->images[i].height * 3;
        texels.insert(texels.end(), d->images[i].rgb, d->images[i].rgb + bytes);
    }
    std::vector<double>& gamma_lut = hb.gamma_lut;
    gamma_lut.resize(256);
    for (int c = 0; c < 256; ++c) gamma_lut[c] = std::pow((double)c / 255.0, 2.2);  // Color::from_rgb color.rs:39-46
    // lights + LightSampler::new (light.rs:187-199); world radius = half the BVH diagonal (scene.rs:42)
    const double world_radius = magnitude(ref.bounds.hi - ref.bounds.lo) * 0.5;
    std::vector<DevLight>& lights = hb.lights;
    std::vector<double>& cdf = hb.cdf;
    lights.resize(d->n_lights); cdf.resize(d->n_lights);
    double total_power = 0.0;
    for (size_t i = 0; i < lights.size(); ++i) {
        const cray_light_desc& l = d->lights[i];
        DevLight dl{};
        dl.kind = l.kind;
        dl.prim = l.primitive;
        std::memcpy(dl.v, l.v, 24);
        std::memcpy(dl.color, l.color, 24);
        Color3 c = mkc(l.color[0], l.color[1], l.color[2]), power;
        switch (l.kind) {  // Light::power light.rs:170-177
            case CRAY_LIGHT_POINT: power = c * 4.0 * kPi; break;
            case CRAY_LIGHT_DISTANT:
            case CRAY_LIGHT_INFINITE: power = c * kPi * world_radius * world_radius; break;
            default:
                dl.shape = make_leaf_prim(*d, (uint32_t)l.primitive);
                dl.area = shape_area(*d, d->primitives[l.primitive]);
                power = c * kPi * dl.area;
                break;
        }
        const double power_avg = (power.r + power.g + power.b) / 3.0;
        total_power += power_avg;
        cdf[i] = total_power;
        lights[i] = dl;
    }
    for (double& c : cdf) c = c / total_power;
    // camera (camera.rs:55-129)
    DevCamera& cam = hb.cam;
    {
        const cray_camera_desc& c = d->camera;
        const Xform wfc = xf_look_at(v3(c.origin), v3(c.target), v3(c.up));
        const Xform sfc = c.kind == CRAY_CAMERA_PERSPECTIVE ? xf_perspective(c.fov, 1e-2, 1000.0) : xf_orthographic(0.0, 1.0);
        const Xform cfr = camera_from_raster(sfc, c.width);
        std::memcpy(cam.camera_from_raster, cfr.fwd.m, sizeof(cam.camera_from_raster));
        cam.world_from_camera = affine_of(wfc.fwd);
        cam.lens_radius = c.lens_radius;
        cam.focal_distance = c.focal_distance;
        cam.perspective = c.kind == CRAY_CAMERA_PERSPECTIVE;
        cam.width = c.width;
        cam.height = c.height;
    }
    // pixel order: 8 x 4 pixel tiles (one warp = one tile), tiles row-major
    std::vector<uint32_t>& pixel_order = hb.pixel_order;
    pixel_order.reserve((size_t)cam.width * cam.height);
    for (uint32_t ty = 0; ty < cam.height; ty += 4)
        for (uint32_t tx = 0; tx < cam.width; tx += 8)
            for (uint32_t y = ty; y < std::min(ty + 4, cam.height); ++y)
                for (uint32_t x = tx; x < std::min(tx + 8, cam.width); ++x) pixel_order.push_back(x | (y << 16));
    // Sobol direction vectors folded into byte tables (sampler.cuh:sobol_sample): [dimension][half][byte]
#if !CRAY_SOBOL_BYTES
    std::vector<uint32_t>& sobol = hb.sobol;
    sobol.assign(&SOBOL_DIRECTIONS_INIT[0][0], &SOBOL_DIRECTIONS_INIT[0][0] + 256 * 32);
#else
    std::vector<uint32_t>& sobol = hb.sobol;
    sobol.resize(256 * 512);
    for (uint32_t dim = 0; dim < 256; ++dim)
        for (uint32_t half = 0; half < 2; ++half)
            for (uint32_t b = 0; b < 256; ++b) {
                uint32_t acc = 0;
                for (uint32_t j = 0; j < 8; ++j)
                    if (b & (0x80u >> j)) acc ^= SOBOL_DIRECTIONS_INIT[dim][8 * half + j];
                sobol[dim * 512 + half * 256 + b] = acc;
            }
#endif

    timer.mark("materials, lights, camera");
    hb.spheres.assign(d->spheres, d->spheres + d->n_spheres);
    return CRAY_OK;
}

// One device copy of a HostBuild.
int upload_scene(const HostBuild& hb, const cray_scene_desc* d, int device, cray_scene** out) {
    CRAY_CUDA(cudaSetDevice(device));
    const auto t0 = std::chrono::steady_clock::now();
    PhaseTimer timer;
    auto sc = new cray_scene();
    sc->device = device;
    sc->build_flags = hb.build_flags;
    struct Guard { cray_scene* s; ~Guard() { if (s) cray_scene_destroy(s); } } guard{sc};
    CRAY_CUDA(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
    SceneView& v = sc->view;
    const uint32_t* d_order = nullptr;
    const uint32_t* d_sobol = nullptr;
    int rc = CRAY_OK;
#define UP(vec, field) do { rc = upload(sc, vec, &field); if (rc != CRAY_OK) return rc; } while (0)
    UP(hb.ref.nodes, v.bin_nodes);
    UP(hb.bin_prims, v.bin_prims);
    UP(hb.wide.nodes, v.wide_nodes);
    UP(hb.wide_prims, v.wide_prims);
    UP(hb.wide_tris32, v.wide_tris32);
    UP(hb.rank_of_prim, v.rank_of_prim);
    UP(hb.wide_slot_of_prim, v.wide_slot_of_prim);
    if (hb.contact.n_nodes) UP(hb.contact.node_flags, v.bin_contact);
    UP(hb.disks, v.disks);
    UP(hb.prims, v.prims);
    UP(hb.tri_shade, v.tri_shade);
    UP(hb.spheres, v.spheres);
    UP(hb.materials, v.materials);
    UP(hb.images, v.images);
    UP(hb.texels, v.texels);
    UP(hb.gamma_lut, v.gamma_lut);
    UP(hb.lights, v.lights);
    UP(hb.cdf, v.light_cdf);
    UP(hb.pixel_order, d_order);
    UP(hb.sobol, d_sobol);
#undef UP
    sc->d_pixel_order = const_cast<uint32_t*>(d_order);
    sc->d_sobol = const_cast<uint32_t*>(d_sobol);
    v.n_lights = (uint32_t)d->n_lights;
    v.max_depth = d->max_depth;
    v.camera = hb.cam;
    v.bounds = hb.ref.bounds;
    CRAY_CUDA(cudaDeviceSynchronize());
    timer.mark("upload");

    cray_scene_info& info = sc->info;
    info.n_primitives = d->n_primitives;
    info.n_lights = d->n_lights;
    info.exact_nodes = hb.ref.nodes.size();
    info.exact_bytes = hb.ref.nodes.size() * sizeof(BinNode);
    info.wide_nodes = hb.wide.nodes.size();
    info.wide_bytes = hb.wide.nodes.size() * sizeof(WideNode);
    info.leaf_prim_bytes = (uint64_t)d->n_primitives * sizeof(LeafPrim);
    info.wide_depth = hb.wide.depth;
    info.width = hb.cam.width; info.height = hb.cam.height;
    info.max_depth = d->max_depth; info.num_samples = d->num_samples;
    info.contact_nodes = hb.contact.n_nodes;
    info.contact_primitives = hb.contact.n_prims;
    info.bvh_build_ms = hb.build_ms;
    info.upload_ms = ms_since(t0);
    guard.s = nullptr;
    *out = sc;
    return CRAY_OK;
}

int check_devices(const int* devices, int n) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { set_error("no CUDA device available (this library has no CPU fallback)"); return CRAY_E_CUDA; }
    for (int k = 0; k < n; ++k)
        if (devices[k] < 0 || devices[k] >= count) { set_error("CUDA device index out of range"); return CRAY_E_INVALID; }
    return CRAY_OK;
}

}  // namespace
}  // namespace cray

using namespace cray;

extern "C" {

const char* cray_last_error(void) { return cray::last_error().c_str(); }
const char* cray_version(void) { return "craytracer_b200 0.1 (sm_100a)"; }
void cray_free(void* p) { std::free(p); }

// Tests: the same dump with the tree built on `device` (bvh_build_gpu.cu); CRAY_E_UNSUPPORTED if the device build was not used.
int cray_debug_build_reference_bvh_on(const cray_scene_desc* desc, int device, cray_bvh_node_dump** nodes, uint64_t* n_nodes, uint32_t** prim_order, uint64_t* n_prims);
int cray_build_reference_bvh(const cray_scene_desc* desc, cray_bvh_node_dump** nodes, uint64_t* n_nodes, uint32_t** prim_order, uint64_t* n_prims) {
    return cray_debug_build_reference_bvh_on(desc, -1, nodes, n_nodes, prim_order, n_prims);
}
int cray_debug_build_reference_bvh_on(const cray_scene_desc* desc, int device, cray_bvh_node_dump** nodes, uint64_t* n_nodes, uint32_t** prim_order, uint64_t* n_prims) {
    if (!desc || desc->n_primitives == 0 || !nodes || !n_nodes || !prim_order || !n_prims) { set_error("bad arguments"); return CRAY_E_INVALID; }
    RefBvh bvh;
    build_reference_bvh(*desc, bvh, 0, device);
    if (!bvh.error.empty()) { set_error(bvh.error); return CRAY_E_BVH; }
    finish_device_build_release();
    if (device >= 0 && !last_reference_build_was_on_device()) { set_error("the tree was built on the host"); return CRAY_E_UNSUPPORTED; }
    auto* out = (cray_bvh_node_dump*)std::malloc(sizeof(cray_bvh_node_dump) * bvh.nodes.size());
    auto* order = (uint32_t*)std::malloc(sizeof(uint32_t) * bvh.prim_order.size());
    for (size_t i = 0; i < bvh.nodes.size(); ++i) {
        const BinNode& n = bvh.nodes[i];
        out[i] = {{n.box.lo.x, n.box.lo.y, n.box.lo.z}, {n.box.hi.x, n.box.hi.y, n.box.hi.z}, n.axis, n.a, n.b, 0};
    }
    std::memcpy(order, bvh.prim_order.data(), sizeof(uint32_t) * bvh.prim_order.size());
    *nodes = out; *n_nodes = bvh.nodes.size();
    *prim_order = order; *n_prims = bvh.prim_order.size();
    return CRAY_OK;
}

void cray_scene_destroy(cray_scene* sc) {
    if (!sc) return;
    cudaSetDevice(sc->device);
    extern void cray_pool_release(cray_scene*);
    cray_pool_release(sc);
    for (void* p : sc->allocations) cudaFree(p);
    if (sc->stream) cudaStreamDestroy(sc->stream);
    delete sc;
}

int cray_scene_create(const cray_scene_desc* d, int device, uint32_t build_flags, cray_scene** out) {
    return cray_scene_create_multi(d, &device, 1, build_flags, out);
}

// The same scene on several devices of this process: the host side (BVH build, record arrays) runs once.
int cray_scene_create_multi(const cray_scene_desc* d, const int* devices, int n, uint32_t build_flags, cray_scene** out) {
    if (!out || !devices || n <= 0) { set_error("bad arguments"); return CRAY_E_INVALID; }
    for (int k = 0; k < n; ++k) out[k] = nullptr;
    int rc = validate(d);
    if (rc != CRAY_OK) return rc;
    if (build_flags == 0) build_flags = CRAY_BUILD_EXACT | CRAY_BUILD_FAST;
    if (build_flags & CRAY_BUILD_F32) build_flags |= CRAY_BUILD_FAST;  // the f32 records hang off the wide BVH's leaf slots
    build_flags |= CRAY_BUILD_EXACT;  // the binary tree also resolves exact-t ties for the fast mode
    rc = check_devices(devices, n);
    if (rc != CRAY_OK) return rc;
    // (the first device's context exists before the build is timed: the tree of a large scene is built there)
    if (cudaSetDevice(devices[0]) == cudaSuccess) cudaFree(nullptr);
    cudaGetLastError();
    HostBuild hb;
    rc = build_host_side(d, build_flags, hb, devices[0]);
    finish_device_build_release();   // (the build buffers of a device-built tree were freed while the host collapsed it)
    if (rc != CRAY_OK) return rc;
    // one uploading thread per device (cray_last_error is thread-local: carry a failure back to this thread)
    std::vector<int> rcs(n, CRAY_OK);
    std::vector<std::string> errors(n);
    std::vector<std::thread> pool;
    auto work = [&](int k) {
        rcs[k] = upload_scene(hb, d, devices[k], &out[k]);
        if (rcs[k] != CRAY_OK) errors[k] = last_error();
    };
    for (int k = 1; k < n; ++k) pool.emplace_back(work, k);
    work(0);
    for (auto& th : pool) th.join();
    for (int k = 0; k < n; ++k)
        if (rcs[k] != CRAY_OK) {
            for (int j = 0; j < n; ++j) { if (out[j]) cray_scene_destroy(out[j]); out[j] = nullptr; }
            set_error(errors[k]);
            return rcs[k];
        }
    return CRAY_OK;
}

int cray_scene_get_info(const cray_scene* sc, cray_scene_info* out) {
    if (!sc || !out) { set_error("bad arguments"); return CRAY_E_INVALID; }
    *out = sc->info;
    return CRAY_OK;
}

}  // extern "C"
