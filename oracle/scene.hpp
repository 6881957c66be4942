// ORACLE — TEST INFRASTRUCTURE ONLY (see geometry.hpp).
// Restatement of src/primitive.rs, src/bvh.rs (+ src/util.rs), src/light.rs, src/camera.rs,
// src/scene.rs, src/path_integrator.rs and the tile renderer of src/bin/craytracer.rs.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "geometry.hpp"
#include "material.hpp"
#include "sampling.hpp"
#include "shape.hpp"

namespace orc {

struct Primitive {  // primitive.rs:15-25
    Shape shape;
    int material;    // index into Scene::materials (area lights: the black matte of primitive.rs:43-46)
    int area_light;  // index into Scene::lights or -1
};

struct PrimitiveIntersection {  // intersection.rs:17-24
    double distance;
    V3 location, normal;
    double uv[2];
    int primitive;  // index in reference primitive order
    double bary_u, bary_v;
};

enum LightKind { L_POINT = 0, L_DISTANT = 1, L_INFINITE = 2, L_AREA = 3 };
struct Light {  // light.rs:25-43
    int kind;
    V3 v;          // origin | direction
    Color color;   // intensity | emittance
    int primitive; // AREA: primitive whose shape emits
};

// ---- util.rs:4-26 ---------------------------------------------------------------------
template <class T, class F>
size_t partition_by(T* slice, size_t len, F&& pred) {
    if (len == 0) return 0;
    size_t left = 0, right = len - 1;
    while (left != right) {
        while (left < right && pred(slice[left])) left += 1;
        while (right > left && !pred(slice[right])) right -= 1;
        std::swap(slice[left], slice[right]);
    }
    return pred(slice[left]) ? left + 1 : left;
}

// ---- bvh.rs ----------------------------------------------------------------------------
struct BvhNode {  // bvh.rs:13-24, flattened into a pre-order array
    Bounds bounds;
    int axis;        // 0..2 interior, 3 leaf
    uint32_t a, b;   // interior: left, right node index; leaf: first, count into prim_order
};
struct PrimitiveInfo {  // bvh.rs:149-154
    uint32_t primitive;
    Bounds bounds;
    V3 centroid;
};

struct Bvh {
    std::vector<BvhNode> nodes;
    std::vector<uint32_t> prim_order;
    Bounds bounds;
    std::string error;  // set when the reference would panic (bvh.rs:245, :327-328)

    uint32_t leaf_node(const PrimitiveInfo* infos, size_t n, const Bounds& b) {  // bvh.rs:181-189
        uint32_t idx = (uint32_t)nodes.size();
        nodes.push_back({b, 3, (uint32_t)prim_order.size(), (uint32_t)n});
        for (size_t i = 0; i < n; ++i) prim_order.push_back(infos[i].primitive);
        return idx;
    }

    uint32_t from_sah_splitting(PrimitiveInfo* infos, size_t n) {  // bvh.rs:234-336
        constexpr size_t NUM_BUCKETS = 12;
        constexpr double TRAVERSAL_TO_INTERSECTION_COST_RATIO = 1.0 / 8.0;
        constexpr size_t MAX_LEAF_PRIMITIVES = 4;

        Bounds bounds = infos[0].bounds;
        for (size_t i = 1; i < n; ++i) bounds = bunion(bounds, infos[i].bounds);
        if (n <= 1) return leaf_node(infos, n, bounds);

        double total_surface_area = bounds.surface_area();
        if (!(total_surface_area > 0.0)) {
            if (error.empty()) error = "Encountered primitives with no surface area";
            return leaf_node(infos, n, bounds);
        }
        Bounds centroid_bounds = Bounds::make(infos[0].centroid, infos[0].centroid);
        for (size_t i = 1; i < n; ++i) centroid_bounds = bunion(centroid_bounds, Bounds::make(infos[i].centroid, infos[i].centroid));
        int split_axis = centroid_bounds.maximum_extent();

        struct Bucket { bool some = false; Bounds bounds{}; size_t count = 0; };
        Bucket buckets[NUM_BUCKETS];
        auto get_bucket_idx = [&](const PrimitiveInfo& p) -> size_t {
            double centroid_offset = centroid_bounds.offset(p.centroid)[split_axis];
            size_t idx = (size_t)as_usize((double)NUM_BUCKETS * centroid_offset);
            return std::min(idx, NUM_BUCKETS - 1);
        };
        for (size_t i = 0; i < n; ++i) {
            size_t bi = get_bucket_idx(infos[i]);
            if (buckets[bi].some) { buckets[bi].bounds = bunion(buckets[bi].bounds, infos[i].bounds); buckets[bi].count += 1; }
            else { buckets[bi].some = true; buckets[bi].bounds = infos[i].bounds; buckets[bi].count = 1; }
        }
        double costs[NUM_BUCKETS - 1];
        for (size_t i = 0; i < NUM_BUCKETS - 1; ++i) {
            double cost = TRAVERSAL_TO_INTERSECTION_COST_RATIO;
            for (int part = 0; part < 2; ++part) {
                size_t lo = part == 0 ? 0 : i + 1, hi = part == 0 ? i + 1 : NUM_BUCKETS;
                Bucket merged;
                for (size_t k = lo; k < hi; ++k) {
                    if (!buckets[k].some) continue;
                    if (merged.some) { merged.bounds = bunion(merged.bounds, buckets[k].bounds); merged.count += buckets[k].count; }
                    else merged = buckets[k];
                }
                if (merged.some) cost += (double)merged.count * merged.bounds.surface_area() / total_surface_area;
            }
            if (!std::isfinite(cost) && error.empty()) error = "SAH cost is not finite";
            costs[i] = cost;
        }
        size_t min_cost_bucket_idx = 0;
        for (size_t i = 0; i < NUM_BUCKETS - 1; ++i)
            if (costs[i] < costs[min_cost_bucket_idx]) min_cost_bucket_idx = i;

        double leaf_cost = (double)n;
        if (leaf_cost <= costs[min_cost_bucket_idx] && n <= MAX_LEAF_PRIMITIVES) return leaf_node(infos, n, bounds);

        size_t mid = partition_by(infos, n, [&](const PrimitiveInfo& p) { return get_bucket_idx(p) <= min_cost_bucket_idx; });
        if (mid == 0 || mid == n) {  // assert!(left.len() > 0); assert!(right.len() > 0);
            if (error.empty()) error = "SAH split produced an empty side";
            return leaf_node(infos, n, bounds);
        }
        uint32_t idx = (uint32_t)nodes.size();
        nodes.push_back({bounds, split_axis, 0, 0});
        uint32_t l = from_sah_splitting(infos, mid);
        uint32_t r = from_sah_splitting(infos + mid, n - mid);
        nodes[idx].a = l;
        nodes[idx].b = r;
        return idx;
    }

    // Median split, bvh.rs:191-230.  select_nth_unstable_by's permutation is implementation-defined in
    // Rust; only the <= 4 primitive case (a single leaf, as in tests/test_bvh.rs) is order-exact here.
    uint32_t from_median_splitting(PrimitiveInfo* infos, size_t n) {
        Bounds bounds = infos[0].bounds;
        for (size_t i = 1; i < n; ++i) bounds = bunion(bounds, infos[i].bounds);
        if (n <= 4) return leaf_node(infos, n, bounds);
        Bounds cb = Bounds::make(infos[0].centroid, infos[0].centroid);
        for (size_t i = 1; i < n; ++i) cb = bunion(cb, Bounds::make(infos[i].centroid, infos[i].centroid));
        int axis = cb.maximum_extent();
        if (cb.min[axis] == cb.max[axis]) return leaf_node(infos, n, bounds);
        size_t mid = (n - 1) / 2;
        std::nth_element(infos, infos + mid, infos + n, [&](const PrimitiveInfo& a, const PrimitiveInfo& b) { return a.centroid[axis] < b.centroid[axis]; });
        uint32_t idx = (uint32_t)nodes.size();
        nodes.push_back({bounds, axis, 0, 0});
        uint32_t l = from_median_splitting(infos, mid);
        uint32_t r = from_median_splitting(infos + mid, n - mid);
        nodes[idx].a = l;
        nodes[idx].b = r;
        return idx;
    }

    void build(const std::vector<Primitive>& prims, bool sah) {  // Bvh::new bvh.rs:38-56
        std::vector<PrimitiveInfo> infos(prims.size());
        for (size_t i = 0; i < prims.size(); ++i) {
            Bounds b = prims[i].shape.bounds();
            infos[i] = {(uint32_t)i, b, b.centroid()};
        }
        nodes.clear();
        prim_order.clear();
        nodes.reserve(prims.size());
        prim_order.reserve(prims.size());
        if (sah) from_sah_splitting(infos.data(), infos.size());
        else from_median_splitting(infos.data(), infos.size());
        bounds = infos[0].bounds;
        for (size_t i = 1; i < infos.size(); ++i) bounds = bunion(bounds, infos[i].bounds);
    }
};

// ---- light.rs:182-219 ----------------------------------------------------------------------
struct LightSampler {
    std::vector<double> cdfs;
    void sample(double u, size_t& idx, double& p) const {  // :203-211 (binary_search_by total_cmp; Err(i) = insertion point)
        size_t lo = 0, hi = cdfs.size();
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (cdfs[mid] < u) lo = mid + 1;
            else hi = mid;
        }
        idx = lo;
        p = pdf(idx);
    }
    double pdf(size_t i) const {  // :213-219
        if (i > 0) return cdfs[i] - cdfs[i - 1];
        return cdfs[i];
    }
};

struct Camera {  // camera.rs
    uint32_t width, height;
    Transformation camera_from_raster, world_from_camera;
    double lens_radius, focal_distance;
    bool perspective;

    static Transformation get_camera_from_raster(const Transformation& screen_from_camera, uint32_t film_w) {  // :25-53
        double film_width = (double)film_w;
        double film_height = (double)film_w;  // sic: camera.rs:30 uses film.width for both
        double screen_width, screen_height;
        if (film_width > film_height) { screen_width = film_width / film_height; screen_height = 1.0; }
        else { screen_width = 1.0; screen_height = film_height / film_width; }
        Transformation screen_from_raster = tmul(scale(2.0 * screen_width / film_width, -2.0 * screen_height / film_height, 1.0),
                                                 translate(-film_width / 2.0, -film_height / 2.0, 0.0));
        return tmul(screen_from_camera.inverse(), screen_from_raster);
    }
    static Camera make(bool persp, uint32_t w, uint32_t h, V3 origin, V3 target, V3 up, double fov, double lens_radius, double focal_distance) {
        Camera c;
        c.width = w; c.height = h;
        c.perspective = persp;
        c.lens_radius = lens_radius;
        c.focal_distance = focal_distance;
        c.world_from_camera = look_at(origin, target, up);  // :67
        Transformation sfc = persp ? orc::perspective(fov, 1e-2, 1000.0) : orthographic(0.0, 1.0);  // :85-91, :114-117
        c.camera_from_raster = get_camera_from_raster(sfc, w);
        return c;
    }
    Ray generate_ray(double lens_u, double lens_v, V3 p_camera) const {  // :147-162
        Ray ray = perspective ? Ray::make(p_camera, normalized(p_camera - V3{0, 0, 0})) : Ray::make(p_camera, V3{0, 0, 1});
        if (lens_radius == 0.0) return ray;
        double lens_x = 2.0 * lens_u - 1.0, lens_y = 2.0 * lens_v - 1.0;
        V3 p_lens{lens_x * lens_radius, lens_y * lens_radius, 0.0};
        V3 p_focal_plane = ray.at(focal_distance / ray.direction.z);
        return Ray::make(p_lens, normalized(p_focal_plane - p_lens));
    }
    Ray sample(double film_u, double film_v, double lens_u, double lens_v, uint64_t raster_x, uint64_t raster_y) const {  // :131-145
        double dx = 2.0 * film_u - 1.0, dy = 2.0 * film_v - 1.0;
        V3 p_raster{(double)raster_x + dx, (double)raster_y + dy, 0.0};
        V3 p_camera = xf_point(camera_from_raster, p_raster);
        Ray ray = generate_ray(lens_u, lens_v, p_camera);
        return xf_ray(world_from_camera, ray);
    }
};

struct RayCounters {
    uint64_t closest = 0, shadow = 0, nan_samples = 0;
};

struct Scene {  // scene.rs:15-22
    uint32_t max_depth, num_samples;
    Camera camera;
    std::vector<Light> lights;
    LightSampler light_sampler;
    std::vector<Primitive> primitives;
    std::vector<Material> materials;
    std::vector<Image> images;
    Bvh bvh;

    void finish(bool sah = true) {  // Scene::new scene.rs:25-53
        bvh.build(primitives, sah);
        double world_radius = magnitude(bvh.bounds.diagonal()) * 0.5;
        // LightSampler::new light.rs:187-199
        double total_power = 0.0;
        light_sampler.cdfs.clear();
        for (const Light& l : lights) {
            Color p = power(l, world_radius);
            double power_avg = (p.r + p.g + p.b) / 3.0;
            total_power += power_avg;
            light_sampler.cdfs.push_back(total_power);
        }
        for (double& c : light_sampler.cdfs) c = c / total_power;
    }

    Color power(const Light& l, double world_radius) const {  // light.rs:170-177
        switch (l.kind) {
            case L_POINT: return l.color * 4.0 * PI;
            case L_DISTANT:
            case L_INFINITE: return l.color * PI * world_radius * world_radius;
            default: return l.color * PI * primitives[l.primitive].shape.area();
        }
    }

    bool primitive_intersect(int pi, Ray& ray, PrimitiveIntersection& out) const {  // primitive.rs:50-73
        ShapeIntersection si;
        if (!primitives[pi].shape.intersect(ray, si)) return false;
        out.distance = ray.max_distance;
        out.location = si.location;
        out.normal = si.normal;
        out.uv[0] = si.uv[0]; out.uv[1] = si.uv[1];
        out.primitive = pi;
        out.bary_u = si.bary_u; out.bary_v = si.bary_v;
        return true;
    }

    // Bvh::intersect bvh.rs:58-104
    bool intersect(Ray& ray, PrimitiveIntersection& current) const {
        uint32_t stack[256];
        int sp = 0;
        stack[sp++] = 0;
        bool have = false;
        while (sp > 0) {
            const BvhNode& node = bvh.nodes[stack[--sp]];
            if (!node.bounds.intersects(ray) && !node.bounds.contains(ray.origin)) continue;
            if (node.axis == 3) {
                for (uint32_t i = 0; i < node.b; ++i) {
                    PrimitiveIntersection isect;
                    if (primitive_intersect((int)bvh.prim_order[node.a + i], ray, isect)) {
                        if (!have || isect.distance < current.distance) { current = isect; have = true; }
                    }
                }
            } else {
                if (ray.direction[node.axis] < 0.0) { stack[sp++] = node.a; stack[sp++] = node.b; }
                else { stack[sp++] = node.b; stack[sp++] = node.a; }
            }
        }
        return have;
    }
    // Bvh::intersects bvh.rs:106-147
    bool intersects(const Ray& ray) const {
        uint32_t stack[256];
        int sp = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const BvhNode& node = bvh.nodes[stack[--sp]];
            if (!node.bounds.intersects(ray) && !node.bounds.contains(ray.origin)) continue;
            if (node.axis == 3) {
                for (uint32_t i = 0; i < node.b; ++i)
                    if (primitives[bvh.prim_order[node.a + i]].shape.intersects(ray)) return true;
            } else {
                if (ray.direction[node.axis] < 0.0) { stack[sp++] = node.a; stack[sp++] = node.b; }
                else { stack[sp++] = node.b; stack[sp++] = node.a; }
            }
        }
        return false;
    }

    // ---- light.rs ---------------------------------------------------------------------
    Color Le(const Light& l) const { return l.kind == L_INFINITE ? l.color : BLACK; }  // light.rs:161-168
    Pdf pdf_Li(const Light& l, V3 location, V3 normal, V3 w_i) const {  // light.rs:136-143
        switch (l.kind) {
            case L_POINT:
            case L_DISTANT: return Pdf::Delta();
            case L_INFINITE: return Pdf::NonDelta(FRAC_1_PI / 4.0);
            default: return Pdf::NonDelta(primitives[l.primitive].shape.pdf_from(location, normal, w_i));
        }
    }
    struct LightSample {
        Color Li;
        V3 w_i;
        Pdf pdf;
        Ray shadow_ray;
    };
    LightSample sample_Li(const Light& l, double s1, double s2u, double s2v, const PrimitiveIntersection& isect, bool* assert_failed) const {  // light.rs:59-133
        switch (l.kind) {
            case L_POINT: {
                V3 op = l.v - isect.location;
                double dist_squared = magnitude_squared(op);
                double dist = std::sqrt(dist_squared);
                V3 w_i = op / dist;
                Ray shadow_ray = Ray::make(isect.location, w_i);
                shadow_ray.update_max_distance(dist);
                return {l.color / dist_squared, w_i, pdf_Li(l, isect.location, isect.normal, w_i), shadow_ray};
            }
            case L_DISTANT: {
                if (assert_failed && !(std::fabs(magnitude(l.v) - 1.0) <= EPSILON)) *assert_failed = true;
                Ray shadow_ray = Ray::make(isect.location, l.v);
                return {l.color, l.v, pdf_Li(l, isect.location, isect.normal, l.v), shadow_ray};
            }
            case L_INFINITE: {
                V3 normal = s1 < 0.5 ? V3{1, 0, 0} : V3{-1, 0, 0};
                V3 w_i = sample_hemisphere(s2u, s2v, normal);
                Ray shadow_ray = Ray::make(isect.location, w_i);
                return {l.color, w_i, pdf_Li(l, isect.location, isect.normal, w_i), shadow_ray};
            }
            default: {
                const Shape& shape = primitives[l.primitive].shape;
                // Shape::sample_from shape.rs:472-484
                V3 shape_point = shape.sample(s2u, s2v);
                V3 w_i = normalized(shape_point - isect.location);
                double pdf = shape.pdf_from(isect.location, isect.normal, w_i);
                double distance = magnitude(shape_point - isect.location);
                Ray shadow_ray = Ray::make(isect.location, w_i);
                shadow_ray.update_max_distance(distance - EPSILON);
                return {l.color, w_i, Pdf::NonDelta(pdf), shadow_ray};
            }
        }
    }

    // `lights.iter().position(|l| l == light)` path_integrator.rs:116 compares lights BY VALUE; two area lights on
    // bit-identical shapes with equal emittance resolve to the first one.
    bool shapes_equal(const Shape& a, const Shape& b) const {
        if (a.kind != b.kind) return false;
        auto meq = [](const Matrix& x, const Matrix& y) {
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (!(x.m[i][j] == y.m[i][j])) return false;
            return true;
        };
        if (a.kind == TRIANGLE)
            return a.v0 == b.v0 && a.e1 == b.e1 && a.e2 == b.e2 && a.n0 == b.n0 && a.n01 == b.n01 && a.n02 == b.n02 &&
                   a.uv0[0] == b.uv0[0] && a.uv0[1] == b.uv0[1] && a.uv01[0] == b.uv01[0] && a.uv01[1] == b.uv01[1] &&
                   a.uv02[0] == b.uv02[0] && a.uv02[1] == b.uv02[1];
        bool eq = meq(a.object_to_world.matrix, b.object_to_world.matrix) && meq(a.object_to_world.inv, b.object_to_world.inv) &&
                  meq(a.world_to_object.matrix, b.world_to_object.matrix) && meq(a.world_to_object.inv, b.world_to_object.inv) && a.radius == b.radius;
        if (a.kind == DISK) eq = eq && a.inner_radius == b.inner_radius;
        return eq;
    }
    size_t light_position(int light_idx) const {
        const Light& me = lights[light_idx];
        for (size_t i = 0; i < lights.size(); ++i) {
            const Light& o = lights[i];
            if (o.kind != L_AREA) continue;
            if (o.color.r == me.color.r && o.color.g == me.color.g && o.color.b == me.color.b &&
                shapes_equal(primitives[o.primitive].shape, primitives[me.primitive].shape))
                return i;
        }
        return (size_t)light_idx;
    }

    // path_integrator.rs:41-215.  Returns false where the reference would panic (asserts at :208-209 and in callees).
    bool estimate_Li(SobolSampler& sampler, Ray ray, Color& L_out, RayCounters& rc) const {
        Color L = BLACK;
        Color beta = WHITE;
        uint32_t bounces = 0;
        bool is_specular_bounce = true;
        double prev_bsdf_pdf = 0.0;
        bool have_prev = false;
        PrimitiveIntersection prev_intersection{};
        bool assert_failed = false;

        while (bounces < max_depth && !is_black(beta)) {
            V3 w_o = neg(ray.direction);
            PrimitiveIntersection intersection;
            rc.closest += 1;
            if (!intersect(ray, intersection)) {
                if (is_specular_bounce) {
                    for (const Light& light : lights) L += beta * Le(light);
                } else {
                    (void)have_prev;
                    for (size_t light_idx = 0; light_idx < lights.size(); ++light_idx) {
                        Color le = Le(lights[light_idx]);
                        if (!is_black(le)) {
                            double light_pdf = pdf_Li(lights[light_idx], prev_intersection.location, prev_intersection.normal, w_o).value * light_sampler.pdf(light_idx);
                            double weight = power_heuristic(1, light_pdf, 1, prev_bsdf_pdf);
                            L += beta * le * weight;
                        }
                    }
                }
                break;
            }
            V3 normal = intersection.normal;
            V3 location = intersection.location;
            const Primitive& prim = primitives[intersection.primitive];
            const Material& material = materials[prim.material];
            const double* uv = intersection.uv;

            // PathSegmentSamples::from path_integrator.rs:25-36
            double mat_1d = sampler.sample_1d();
            double mat_u, mat_v;
            sampler.sample_2d(mat_u, mat_v);
            double light_index_1d = sampler.sample_1d();
            double light_1d = sampler.sample_1d();
            double light_u, light_v;
            sampler.sample_2d(light_u, light_v);
            double rr_1d = sampler.sample_1d();

            // Emission (:106-126); intersection.Le intersection.rs:28-33, Light::L light.rs:147-157
            Color le = prim.area_light >= 0 ? lights[prim.area_light].color : BLACK;
            if (!is_black(le)) {
                if (is_specular_bounce) {
                    L += beta * le;
                } else {
                    size_t light_idx = light_position(prim.area_light);
                    double light_pdf = pdf_Li(lights[prim.area_light], intersection.location, intersection.normal, w_o).value * light_sampler.pdf(light_idx);
                    double weight = power_heuristic(1, light_pdf, 1, prev_bsdf_pdf);
                    L += beta * le * weight;
                }
            }

            // NEE (:129-164)
            {
                size_t light_index;
                double light_sampler_pdf;
                light_sampler.sample(light_index_1d, light_index, light_sampler_pdf);
                const Light& light = lights[light_index];
                LightSample ls = sample_Li(light, light_1d, light_u, light_v, intersection, &assert_failed);
                rc.shadow += 1;
                if (!intersects(ls.shadow_ray)) {
                    Color f = material.f(w_o, ls.w_i, normal, uv);
                    double cos_theta = std::fabs(dot(ls.w_i, normal));
                    if (!ls.pdf.delta) {
                        if (ls.pdf.value > 0.0) {
                            double light_pdf = ls.pdf.value * light_sampler_pdf;
                            Pdf bp = material.pdf(w_o, ls.w_i, normal);
                            double bsdf_pdf = bp.delta ? 0.0 : bp.value;
                            double weight = power_heuristic(1, light_pdf, 1, bsdf_pdf);
                            L += beta * ls.Li * f * cos_theta * weight / light_pdf;
                        }
                    } else {
                        double light_pdf = light_sampler_pdf;
                        L += beta * ls.Li * f * cos_theta / light_pdf;
                    }
                }
            }

            // BSDF sample (:167-195)
            {
                SurfaceSample ss;
                if (!material.sample(mat_1d, mat_u, mat_v, w_o, normal, uv, ss, &assert_failed)) break;
                if (is_black(ss.f)) break;
                double cos_theta = std::fabs(dot(ss.w_i, normal));
                double bsdf_pdf = ss.pdf.delta ? 1.0 : ss.pdf.value;
                if (bsdf_pdf == 0.0) break;
                beta = beta * ss.f * cos_theta / bsdf_pdf;
                ray = Ray::make(location, ss.w_i);
                is_specular_bounce = ss.is_specular;
                prev_bsdf_pdf = bsdf_pdf;
                prev_intersection = intersection;
                have_prev = true;
            }

            // Russian roulette (:197-206)
            if (bounces > 0) {
                double max_beta_component = rmax(beta.r, rmax(beta.g, beta.b));
                if (max_beta_component < 1.0) {
                    double q = 1.0 - max_beta_component;
                    if (rr_1d < q) break;
                    beta = beta / (1.0 - q);
                }
            }
            if (!is_finite(L) || !is_finite(beta)) { assert_failed = true; break; }  // :208-209
            bounces += 1;
        }
        L_out = L;
        if (assert_failed || !is_finite(L)) { rc.nan_samples += 1; return false; }
        return true;
    }

    // render_pixel  src/bin/craytracer.rs:148-162
    bool render_pixel(SobolSampler& sampler, uint64_t x, uint64_t y, uint64_t sample_index, Color& L, RayCounters& rc, Ray* camera_ray = nullptr) const {
        sampler.start_pixel(x, y, sample_index);
        double fu, fv, lu, lv;
        sampler.sample_2d(fu, fv);
        sampler.sample_2d(lu, lv);
        Ray ray = camera.sample(fu, fv, lu, lv, x, y);
        if (camera_ray) *camera_ray = ray;
        return estimate_Li(sampler, ray, L, rc);
    }

    // generate_tiles + render + render_tile, src/bin/craytracer.rs:22-43, :164-206, :224-291: 64x64 pixel x 8 sample
    // tiles in sample-major order, one atomic tile counter, `threads` workers; per (pixel, batch) the <= 8 colours are
    // summed in f64, cast to f32 and added to the f32 film.  A sample where the reference would panic is dropped
    // (counted in rc.nan_samples).  Film is the SUM over [sample_begin, sample_end).
    void render(uint64_t seed, uint32_t sample_begin, uint32_t sample_end, unsigned threads, float* film, RayCounters& total) const {
        const uint32_t W = camera.width, H = camera.height;
        struct Tile { uint32_t x0, x1, y0, y1, s0, s1; };
        std::vector<Tile> tiles;
        for (uint32_t si = sample_begin; si < sample_end; si += 8)
            for (uint32_t ty = 0; ty < H; ty += 64)
                for (uint32_t tx = 0; tx < W; tx += 64)
                    tiles.push_back({tx, std::min(tx + 64, W), ty, std::min(ty + 64, H), si, std::min(si + 8, sample_end)});
        for (size_t i = 0; i < (size_t)W * H * 3; ++i) film[i] = 0.0f;
        std::atomic<size_t> tile_index{0};
        std::vector<RayCounters> counters(threads);
        auto worker = [&](unsigned tid) {
            SobolSampler sampler;
            sampler.seed = seed;
            RayCounters& rc = counters[tid];
            for (;;) {
                size_t index = tile_index.fetch_add(1);
                if (index >= tiles.size()) break;
                const Tile& t = tiles[index];
                for (uint32_t y = t.y0; y < t.y1; ++y)
                    for (uint32_t x = t.x0; x < t.x1; ++x) {
                        Color color = BLACK;
                        for (uint32_t s = t.s0; s < t.s1; ++s) {
                            Color L;
                            if (render_pixel(sampler, x, y, s, L, rc)) color += L;
                        }
                        size_t offset = (size_t)x + (size_t)y * W;
                        // Distinct tiles of one sample batch never share a pixel, but batches do: the reference
                        // serialises with a mutex (craytracer.rs:185); float atomics give the same sums up to order.
                        std::atomic_ref<float> r(film[3 * offset]), g(film[3 * offset + 1]), b(film[3 * offset + 2]);
                        r.fetch_add((float)color.r);
                        g.fetch_add((float)color.g);
                        b.fetch_add((float)color.b);
                    }
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker, t);
        worker(0);
        for (auto& th : pool) th.join();
        for (auto& c : counters) { total.closest += c.closest; total.shadow += c.shadow; total.nan_samples += c.nan_samples; }
    }
};

}  // namespace orc
