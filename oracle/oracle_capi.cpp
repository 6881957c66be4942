// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement of the
// reference in this directory.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product never does.
//
// Input is the same flat `cray_scene_desc` (include/cray_b200.h) the product consumes, so oracle and
// GPU see bit-identical scenes.
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "../include/cray_b200.h"
#include "scene.hpp"

using namespace orc;

namespace {
thread_local std::string g_error;

V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
Color col(const double* p) { return {p[0], p[1], p[2]}; }

Texture make_texture(const cray_texture_desc& t, const std::vector<Image>& images) {
    Texture r;
    r.kind = (int)t.kind;
    r.a = col(t.a);
    r.b = col(t.b);
    r.scale = t.scale;
    if (t.kind == CRAY_TEX_IMAGE) r.image = &images[t.image];
    return r;
}

struct OracleScene {
    Scene scene;
};

bool build_shape(const cray_scene_desc* d, const cray_primitive_desc& p, Shape& out) {
    switch (p.shape_kind) {
        case CRAY_SHAPE_SPHERE: {
            const cray_sphere_desc& s = d->spheres[p.shape_index];
            out = Shape::new_sphere(v3(s.origin), s.radius);
            return true;
        }
        case CRAY_SHAPE_TRIANGLE: {
            const cray_triangle_desc& t = d->triangles[p.shape_index];
            out = Shape{};
            out.kind = TRIANGLE;
            out.v0 = v3(t.v0); out.e1 = v3(t.e1); out.e2 = v3(t.e2);
            out.n0 = v3(t.n0); out.n01 = v3(t.n01); out.n02 = v3(t.n02);
            for (int i = 0; i < 2; ++i) { out.uv0[i] = t.uv0[i]; out.uv01[i] = t.uv01[i]; out.uv02[i] = t.uv02[i]; }
            return true;
        }
        case CRAY_SHAPE_DISK: {
            const cray_disk_desc& k = d->disks[p.shape_index];
            out = Shape::new_disk(v3(k.origin), k.rotate_x, k.rotate_y, k.radius, k.inner_radius);
            return true;
        }
    }
    return false;
}
Ray to_ray(const cray_ray& r) { return {v3(r.origin), v3(r.direction), r.max_distance}; }

template <class F>
void parallel_for(uint64_t n, int threads, F&& fn) {
    if (threads <= 1 || n < 1024) { fn(0, n); return; }
    std::vector<std::thread> pool;
    uint64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        pool.emplace_back([=, &fn] { fn(b, e); });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

const char* orc_last_error() { return g_error.c_str(); }

void* orc_scene_create(const cray_scene_desc* d, int use_sah) {
    if (!d || d->n_primitives == 0) { g_error = "empty scene"; return nullptr; }
    if (d->n_lights == 0) { g_error = "No lights in the scene."; return nullptr; }  // scene_parser.rs:1103
    auto* os = new OracleScene();
    Scene& s = os->scene;
    s.max_depth = d->max_depth;
    s.num_samples = d->num_samples;
    const cray_camera_desc& c = d->camera;
    s.camera = Camera::make(c.kind == CRAY_CAMERA_PERSPECTIVE, c.width, c.height, v3(c.origin), v3(c.target), v3(c.up), c.fov, c.lens_radius, c.focal_distance);
    s.images.resize(d->n_images);
    for (uint64_t i = 0; i < d->n_images; ++i) s.images[i] = {d->images[i].width, d->images[i].height, d->images[i].rgb};
    for (uint64_t i = 0; i < d->n_materials; ++i) {
        const cray_material_desc& m = d->materials[i];
        Texture t0 = make_texture(m.t0, s.images), t1 = make_texture(m.t1, s.images), t2 = make_texture(m.t2, s.images);
        switch (m.kind) {
            case CRAY_MAT_MATTE: s.materials.push_back(Material::new_matte(t0, t2)); break;
            case CRAY_MAT_GLASS: s.materials.push_back(Material::new_glass(t0, t1, m.eta)); break;
            case CRAY_MAT_PLASTIC: s.materials.push_back(Material::new_plastic(t0, t1, t2)); break;
            case CRAY_MAT_METAL: s.materials.push_back(Material::new_metal(t0, t1)); break;
            default: g_error = "unknown material kind"; delete os; return nullptr;
        }
    }
    // primitive.rs:43-46: area-light primitives carry a black matte material
    int black = (int)s.materials.size();
    s.materials.push_back(Material::new_matte(Texture{}, Texture{}));
    for (uint64_t i = 0; i < d->n_lights; ++i) {
        const cray_light_desc& l = d->lights[i];
        s.lights.push_back({(int)l.kind, v3(l.v), col(l.color), l.primitive});
    }
    s.primitives.resize(d->n_primitives);
    for (uint64_t i = 0; i < d->n_primitives; ++i) {
        const cray_primitive_desc& p = d->primitives[i];
        if (!build_shape(d, p, s.primitives[i].shape)) { g_error = "unknown shape kind"; delete os; return nullptr; }
        s.primitives[i].area_light = p.area_light;
        s.primitives[i].material = p.area_light >= 0 ? black : p.material;
    }
    s.finish(use_sah != 0);
    if (!s.bvh.error.empty()) { g_error = s.bvh.error; delete os; return nullptr; }
    return os;
}
void orc_scene_destroy(void* h) { delete (OracleScene*)h; }

// Scene::intersect on a batch.  hits[i].prim = CRAY_NO_HIT on a miss.
void orc_intersect(void* h, const cray_ray* rays, uint64_t n, cray_hit* hits, cray_surface* surf, int threads) {
    const Scene& s = ((OracleScene*)h)->scene;
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            Ray ray = to_ray(rays[i]);
            PrimitiveIntersection pi{};
            if (s.intersect(ray, pi)) {
                hits[i] = {(uint32_t)pi.primitive, 0, pi.distance, pi.bary_u, pi.bary_v};
                if (surf) surf[i] = {{pi.location.x, pi.location.y, pi.location.z}, {pi.normal.x, pi.normal.y, pi.normal.z}, {pi.uv[0], pi.uv[1]}};
            } else {
                hits[i] = {CRAY_NO_HIT, 0, 0.0, 0.0, 0.0};
                if (surf) std::memset(&surf[i], 0, sizeof(cray_surface));
            }
        }
    });
}
void orc_intersects(void* h, const cray_ray* rays, uint64_t n, uint8_t* occluded, int threads) {
    const Scene& s = ((OracleScene*)h)->scene;
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) occluded[i] = s.intersects(to_ray(rays[i])) ? 1 : 0;
    });
}

// SobolSampler::start_pixel + Camera::sample for each (x, y, sample_index): the ray render_pixel traces.
void orc_camera_rays(void* h, uint64_t seed, const uint32_t* x, const uint32_t* y, const uint32_t* si, uint64_t n, cray_ray* out) {
    const Scene& s = ((OracleScene*)h)->scene;
    SobolSampler sampler;
    sampler.seed = seed;
    for (uint64_t i = 0; i < n; ++i) {
        sampler.start_pixel(x[i], y[i], si[i]);
        double fu, fv, lu, lv;
        sampler.sample_2d(fu, fv);
        sampler.sample_2d(lu, lv);
        Ray r = s.camera.sample(fu, fv, lu, lv, x[i], y[i]);
        out[i] = {{r.origin.x, r.origin.y, r.origin.z}, {r.direction.x, r.direction.y, r.direction.z}, r.max_distance};
    }
}

// render_pixel (+ estimate_Li) for each (x, y, sample_index).  ok[i] = 0 where the reference would panic.
void orc_estimate_li(void* h, uint64_t seed, const uint32_t* x, const uint32_t* y, const uint32_t* si, uint64_t n, double* rgb, uint8_t* ok, int threads) {
    const Scene& s = ((OracleScene*)h)->scene;
    parallel_for(n, threads, [&](uint64_t b, uint64_t e) {
        SobolSampler sampler;
        sampler.seed = seed;
        RayCounters rc;
        for (uint64_t i = b; i < e; ++i) {
            Color L;
            bool good = s.render_pixel(sampler, x[i], y[i], si[i], L, rc);
            rgb[3 * i] = L.r; rgb[3 * i + 1] = L.g; rgb[3 * i + 2] = L.b;
            if (ok) ok[i] = good ? 1 : 0;
        }
    });
}

// The secondary rays of the first `depth` vertices of a path (fixed-batch B3 of SURVEY 8d): for each
// (x,y,sample) returns the shadow ray and the continuation ray generated at the first hit vertex.
// valid[i] bit0: shadow ray valid, bit1: continuation ray valid.
void orc_bounce_rays(void* h, uint64_t seed, const uint32_t* x, const uint32_t* y, const uint32_t* si, uint64_t n,
                     cray_ray* shadow, cray_ray* cont, uint8_t* valid) {
    const Scene& s = ((OracleScene*)h)->scene;
    SobolSampler sampler;
    sampler.seed = seed;
    for (uint64_t i = 0; i < n; ++i) {
        valid[i] = 0;
        sampler.start_pixel(x[i], y[i], si[i]);
        double fu, fv, lu, lv;
        sampler.sample_2d(fu, fv);
        sampler.sample_2d(lu, lv);
        Ray ray = s.camera.sample(fu, fv, lu, lv, x[i], y[i]);
        V3 w_o = neg(ray.direction);
        PrimitiveIntersection isect;
        if (!s.intersect(ray, isect)) continue;
        double mat_1d = sampler.sample_1d(), mat_u, mat_v;
        sampler.sample_2d(mat_u, mat_v);
        double li_1d = sampler.sample_1d(), l_1d = sampler.sample_1d(), l_u, l_v;
        sampler.sample_2d(l_u, l_v);
        size_t light_index;
        double lp;
        s.light_sampler.sample(li_1d, light_index, lp);
        Scene::LightSample ls = s.sample_Li(s.lights[light_index], l_1d, l_u, l_v, isect, nullptr);
        const Ray& sr = ls.shadow_ray;
        shadow[i] = {{sr.origin.x, sr.origin.y, sr.origin.z}, {sr.direction.x, sr.direction.y, sr.direction.z}, sr.max_distance};
        valid[i] |= 1;
        SurfaceSample ss;
        const Material& m = s.materials[s.primitives[isect.primitive].material];
        if (m.sample(mat_1d, mat_u, mat_v, w_o, isect.normal, isect.uv, ss, nullptr)) {
            cont[i] = {{isect.location.x, isect.location.y, isect.location.z}, {ss.w_i.x, ss.w_i.y, ss.w_i.z}, INF};
            valid[i] |= 2;
        }
    }
}

// render(): film SUM over [sample_begin, sample_end); counts[0..2] = closest rays, shadow rays, dropped samples.
void orc_render(void* h, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, int threads, float* film, uint64_t* counts) {
    const Scene& s = ((OracleScene*)h)->scene;
    RayCounters rc;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    s.render(seed, sample_begin, sample_end, (unsigned)threads, film, rc);
    if (counts) { counts[0] = rc.closest; counts[1] = rc.shadow; counts[2] = rc.nan_samples; }
}

uint64_t orc_bvh_num_nodes(void* h) { return ((OracleScene*)h)->scene.bvh.nodes.size(); }
void orc_bvh_dump(void* h, cray_bvh_node_dump* nodes, uint32_t* prim_order) {
    const Bvh& b = ((OracleScene*)h)->scene.bvh;
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const BvhNode& n = b.nodes[i];
        nodes[i] = {{n.bounds.min.x, n.bounds.min.y, n.bounds.min.z}, {n.bounds.max.x, n.bounds.max.y, n.bounds.max.z}, (uint32_t)n.axis, n.a, n.b, 0};
    }
    for (size_t i = 0; i < b.prim_order.size(); ++i) prim_order[i] = b.prim_order[i];
}
void orc_light_cdf(void* h, double* cdf) {
    const auto& c = ((OracleScene*)h)->scene.light_sampler.cdfs;
    for (size_t i = 0; i < c.size(); ++i) cdf[i] = c[i];
}
// Camera matrices: camera_from_raster then world_from_camera, row-major 4x4 each.
void orc_camera_matrices(void* h, double* out32) {
    const Camera& c = ((OracleScene*)h)->scene.camera;
    std::memcpy(out32, c.camera_from_raster.matrix.m, 16 * sizeof(double));
    std::memcpy(out32 + 16, c.world_from_camera.matrix.m, 16 * sizeof(double));
}

// ---- micro entry points for the reference's known-answer tests -----------------------------------
// shape: kind + params (sphere: origin[3], radius; triangle: v0,v1,v2; disk: origin[3], rx, ry, radius, inner)
static bool make_shape(int kind, const double* p, Shape& s) {
    if (kind == SPHERE) { s = Shape::new_sphere(v3(p), p[3]); return true; }
    if (kind == TRIANGLE) return Shape::new_triangle(v3(p), v3(p + 3), v3(p + 6), s);
    s = Shape::new_disk(v3(p), p[3], p[4], p[5], p[6]);
    return true;
}
// returns 1 on hit; out = location[3], normal[3], uv[2], ray.max_distance after the call
int orc_shape_intersect(int kind, const double* params, const cray_ray* ray, double* out9) {
    Shape s;
    if (!make_shape(kind, params, s)) return -1;
    Ray r = to_ray(*ray);
    ShapeIntersection si;
    bool hit = s.intersect(r, si);
    if (hit) {
        out9[0] = si.location.x; out9[1] = si.location.y; out9[2] = si.location.z;
        out9[3] = si.normal.x; out9[4] = si.normal.y; out9[5] = si.normal.z;
        out9[6] = si.uv[0]; out9[7] = si.uv[1];
    }
    out9[8] = r.max_distance;
    return hit ? 1 : 0;
}
int orc_shape_intersects(int kind, const double* params, const cray_ray* ray) {
    Shape s;
    if (!make_shape(kind, params, s)) return -1;
    return s.intersects(to_ray(*ray)) ? 1 : 0;
}
int orc_shape_bounds(int kind, const double* params, double* out6) {
    Shape s;
    if (!make_shape(kind, params, s)) return -1;
    Bounds b = s.bounds();
    out6[0] = b.min.x; out6[1] = b.min.y; out6[2] = b.min.z; out6[3] = b.max.x; out6[4] = b.max.y; out6[5] = b.max.z;
    return 0;
}
double orc_shape_area(int kind, const double* params) {
    Shape s;
    if (!make_shape(kind, params, s)) return -1.0;
    return s.area();
}
int orc_bounds_intersects(const double* mn, const double* mx, const cray_ray* ray) {
    Bounds b{v3(mn), v3(mx)};
    return b.intersects(to_ray(*ray)) ? 1 : 0;
}
void orc_bounds_union(const double* a6, const double* b6, double* out6) {
    Bounds r = bunion(Bounds{v3(a6), v3(a6 + 3)}, Bounds{v3(b6), v3(b6 + 3)});
    out6[0] = r.min.x; out6[1] = r.min.y; out6[2] = r.min.z; out6[3] = r.max.x; out6[4] = r.max.y; out6[5] = r.max.z;
}
// Vector / Point algebra of src/geometry.rs (the operators tests/test_geometry.rs exercises).  op: 0 a + b, 1 a - b, 2 a * s,
// 3 a / s, 4 cross(a, b), 5 normalized(a); returns dot(a, b) for op 6 and magnitude(a) for op 7 (out untouched).
double orc_vector_op(int op, const double* a3, const double* b3, double s, double* out3) {
    const V3 a = v3(a3), b = b3 ? v3(b3) : V3{0.0, 0.0, 0.0};
    V3 r{0.0, 0.0, 0.0};
    switch (op) {
        case 0: r = a + b; break;
        case 1: r = a - b; break;
        case 2: r = a * s; break;
        case 3: r = a / s; break;
        case 4: r = cross(a, b); break;
        case 5: r = normalized(a); break;
        case 6: return dot(a, b);
        case 7: return magnitude(a);
        default: break;
    }
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
    return 0.0;
}

void orc_reflect(const double* d, const double* n, double* out) {
    V3 r = reflect(v3(d), v3(n));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
int orc_refract(const double* d, const double* n, double cos_theta_i, double eta_i, double eta_t, double* out) {
    V3 r;
    if (!refract(v3(d), v3(n), cos_theta_i, eta_i, eta_t, r)) return 0;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    return 1;
}
double orc_fresnel_dielectric(double eta_i, double eta_t, double c) { return fresnel_dielectric(eta_i, eta_t, c); }
void orc_fresnel_conductor(const double* eta_i, const double* eta_t, const double* k, double c, double* out) {
    Color r = fresnel_conductor(col(eta_i), col(eta_t), col(k), c);
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
}
void orc_matrix_mul(const double* a, const double* b, double* out) {
    Matrix A, B;
    std::memcpy(A.m, a, sizeof(A.m));
    std::memcpy(B.m, b, sizeof(B.m));
    Matrix R = matmul(A, B);
    std::memcpy(out, R.m, sizeof(R.m));
}
int orc_matrix_inverse(const double* a, double* out) {
    Matrix A, R;
    std::memcpy(A.m, a, sizeof(A.m));
    if (!inverse(A, R)) return 0;
    std::memcpy(out, R.m, sizeof(R.m));
    return 1;
}
// kind: 0 translate(p0..2) 1 scale(p0..2) 2 rotate_x(p0) 3 rotate_y 4 rotate_z 5 look_at(origin,target,up) 6 perspective(fov,near,far) 7 orthographic(near,far)
void orc_transformation(int kind, const double* p, double* matrix16, double* inverse16) {
    Transformation t;
    switch (kind) {
        case 0: t = translate(p[0], p[1], p[2]); break;
        case 1: t = scale(p[0], p[1], p[2]); break;
        case 2: t = rotate_x(p[0]); break;
        case 3: t = rotate_y(p[0]); break;
        case 4: t = rotate_z(p[0]); break;
        case 5: t = look_at(v3(p), v3(p + 3), v3(p + 6)); break;
        case 6: t = perspective(p[0], p[1], p[2]); break;
        default: t = orthographic(p[0], p[1]); break;
    }
    std::memcpy(matrix16, t.matrix.m, 16 * sizeof(double));
    std::memcpy(inverse16, t.inv.m, 16 * sizeof(double));
}
// what: 0 point 1 vector 2 normal
void orc_transform_apply(const double* matrix16, const double* inverse16, int what, const double* in3, double* out3) {
    Transformation t;
    std::memcpy(t.matrix.m, matrix16, 16 * sizeof(double));
    std::memcpy(t.inv.m, inverse16, 16 * sizeof(double));
    V3 r = what == 0 ? xf_point(t, v3(in3)) : (what == 1 ? xf_vector(t, v3(in3)) : xf_normal(t, v3(in3)));
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_transform_bounds(const double* matrix16, const double* inverse16, const double* in6, double* out6) {
    Transformation t;
    std::memcpy(t.matrix.m, matrix16, 16 * sizeof(double));
    std::memcpy(t.inv.m, inverse16, 16 * sizeof(double));
    Bounds b = xf_bounds(t, Bounds::make(v3(in6), v3(in6 + 3)));
    out6[0] = b.min.x; out6[1] = b.min.y; out6[2] = b.min.z; out6[3] = b.max.x; out6[4] = b.max.y; out6[5] = b.max.z;
}
void orc_from_rgb(uint8_t r, uint8_t g, uint8_t b, double* out) {
    Color c = from_rgb(r, g, b);
    out[0] = c.r; out[1] = c.g; out[2] = c.b;
}
void orc_to_rgb(const double* c, uint8_t* out) { to_rgb(col(c), out); }
// partition_by on u32 with predicate kind: 0 (> k), 1 (== k), 2 (% 2 == k)
uint64_t orc_partition_by(uint32_t* data, uint64_t n, int pred_kind, uint32_t k) {
    return partition_by(data, n, [&](uint32_t v) { return pred_kind == 0 ? v > k : (pred_kind == 1 ? v == k : (v % 2 == k)); });
}
uint64_t orc_siphash(int c, int d, uint64_t k0, uint64_t k1, const uint8_t* data, uint64_t len) { return siphash(c, d, k0, k1, data, len); }
uint32_t orc_pixel_hash(uint64_t seed, uint64_t x, uint64_t y) { return pixel_hash(seed, x, y); }
float orc_sobol_sample(uint32_t index, uint32_t dim, uint32_t seed) { return sobol_burley_sample(index, dim, seed); }
void orc_sampling_fn(int which, double u, double v, const double* normal, double* out3) {
    switch (which) {
        case 0: { double x, y; sample_disk(u, v, x, y); out3[0] = x; out3[1] = y; out3[2] = 0; break; }
        case 1: { V3 r = sample_sphere(u, v); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; break; }
        case 2: { V3 r = sample_hemisphere(u, v, v3(normal)); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; break; }
        case 3: { double a, b; sample_triangle(u, v, a, b); out3[0] = a; out3[1] = b; out3[2] = 0; break; }
        default: { V3 r = cosine_sample_hemisphere(u, v, v3(normal)); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; break; }
    }
}
// Material::{sample,f,pdf} of scene material `mat` (index into desc.materials).
// out: w_i[3], f[3], pdf, flags (bit0 some, bit1 delta, bit2 specular)
void orc_material_sample(void* h, int mat, const double* s3, const double* w_o, const double* n, const double* uv, double* out8) {
    const Scene& s = ((OracleScene*)h)->scene;
    SurfaceSample ss;
    bool some = s.materials[mat].sample(s3[0], s3[1], s3[2], v3(w_o), v3(n), uv, ss, nullptr);
    std::memset(out8, 0, 8 * sizeof(double));
    if (!some) return;
    out8[0] = ss.w_i.x; out8[1] = ss.w_i.y; out8[2] = ss.w_i.z;
    out8[3] = ss.f.r; out8[4] = ss.f.g; out8[5] = ss.f.b;
    out8[6] = ss.pdf.delta ? 0.0 : ss.pdf.value;
    out8[7] = 1.0 + (ss.pdf.delta ? 2.0 : 0.0) + (ss.is_specular ? 4.0 : 0.0);
}
void orc_material_f_pdf(void* h, int mat, const double* w_o, const double* w_i, const double* n, const double* uv, double* out5) {
    const Scene& s = ((OracleScene*)h)->scene;
    Color f = s.materials[mat].f(v3(w_o), v3(w_i), v3(n), uv);
    Pdf p = s.materials[mat].pdf(v3(w_o), v3(w_i), v3(n));
    out5[0] = f.r; out5[1] = f.g; out5[2] = f.b;
    out5[3] = p.delta ? 0.0 : p.value;
    out5[4] = p.delta ? 1.0 : 0.0;
}

}  // extern "C"
