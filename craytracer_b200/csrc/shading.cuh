// Materials, textures and lights on the device, restating src/texture.rs, src/bxdf.rs, src/bsdf.rs,
// src/material.rs and src/light.rs in f64 with the reference's operation order.
#pragma once
#include "sampler.cuh"
#include "shapes.cuh"

namespace cray {

struct PdfValue {  // Pdf, pdf.rs:1-6
    bool delta;
    double value;
};
__device__ __forceinline__ PdfValue non_delta(double v) { PdfValue p; p.delta = false; p.value = v; return p; }
__device__ __forceinline__ PdfValue delta_pdf() { PdfValue p; p.delta = true; p.value = 0.0; return p; }

// ---- Texture::eval texture.rs:19-47 ----
__device__ __forceinline__ const uint8_t* texel_ptr(const SceneView& s, const cray_texture_desc& t, double u, double v) {
    const DevImage im = s.images[t.image];
    double uu = u - trunc(u);
    if (uu < 0.0) uu += 1.0;
    double vv = v - trunc(v);
    if (vv < 0.0) vv += 1.0;
    const uint32_t x = as_u32((double)(im.width - 1) * uu);
    const uint32_t y = as_u32((double)(im.height - 1) * vv);
    return s.texels + im.offset + 3ull * ((uint64_t)y * im.width + x);
}
__device__ __forceinline__ bool checker_is_a(const cray_texture_desc& t, double u, double v) {
    const uint64_t iu = as_usize(u * t.scale * 2.0);
    const uint64_t iv = as_usize(v * t.scale * 2.0);
    return ((iu & 1ull) ^ (iv & 1ull)) == 0ull;
}
__device__ __forceinline__ Color3 eval_color(const SceneView& s, const cray_texture_desc& t, double u, double v) {
    if (t.kind == CRAY_TEX_CONSTANT) return mkc(t.a[0], t.a[1], t.a[2]);
    if (t.kind == CRAY_TEX_CHECKERBOARD) return checker_is_a(t, u, v) ? mkc(t.a[0], t.a[1], t.a[2]) : mkc(t.b[0], t.b[1], t.b[2]);
    const uint8_t* px = texel_ptr(s, t, u, v);
    return mkc(s.gamma_lut[px[0]], s.gamma_lut[px[1]], s.gamma_lut[px[2]]);  // Color::from_rgb, decoded through a host-built table
}
__device__ __forceinline__ double eval_f64(const SceneView& s, const cray_texture_desc& t, double u, double v) {
    if (t.kind == CRAY_TEX_CONSTANT) return t.a[0];
    if (t.kind == CRAY_TEX_CHECKERBOARD) return checker_is_a(t, u, v) ? t.a[0] : t.b[0];
    const uint8_t* px = texel_ptr(s, t, u, v);
    const uint32_t l = (2126u * px[0] + 7152u * px[1] + 722u * px[2]) / 10000u;  // image::Rgb::to_luma
    return (double)l / 255.0;
}

// ---- bxdf.rs helpers ----
__device__ __forceinline__ V3 reflect(V3 direction, V3 normal) { return normal * (dot(normal, direction) * 2.0) - direction; }  // :287-290
__device__ __forceinline__ bool refract(V3 direction, V3 normal, double cos_theta_i, double eta_i, double eta_t, V3& out) {  // :292-314
    double eta_relative, cos_theta;
    if (sign_negative(cos_theta_i)) { normal = neg(normal); eta_relative = eta_i / eta_t; cos_theta = -cos_theta_i; }
    else { eta_relative = eta_t / eta_i; cos_theta = cos_theta_i; }
    const double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
    if (sin_theta > eta_relative) return false;
    const V3 r_perpendicular = (normal * cos_theta - direction) / eta_relative;
    const V3 r_parallel = normal * -sqrt(1.0 - dot(r_perpendicular, r_perpendicular));
    out = r_perpendicular + r_parallel;
    return true;
}
__device__ __forceinline__ double fresnel_dielectric(double eta_i, double eta_t, double cos_theta_i) {  // :338-357
    if (sign_negative(cos_theta_i)) {
        cos_theta_i = -cos_theta_i;
        const double tmp = eta_i; eta_i = eta_t; eta_t = tmp;
    }
    const double sin_theta_i = sqrt(1.0 - cos_theta_i * cos_theta_i);
    const double sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0) return 1.0;
    const double cos_theta_t = sqrt(1.0 - sin_theta_t * sin_theta_t);
    const double r_parallel = (eta_t * cos_theta_i - eta_i * cos_theta_t) / (eta_t * cos_theta_i + eta_i * cos_theta_t);
    const double r_perpendicular = (eta_i * cos_theta_i - eta_t * cos_theta_t) / (eta_i * cos_theta_i + eta_t * cos_theta_t);
    return (r_parallel * r_parallel + r_perpendicular * r_perpendicular) * 0.5;
}
__device__ __forceinline__ Color3 csqrt(Color3 c) { return mkc(sqrt_rn(c.r), sqrt_rn(c.g), sqrt_rn(c.b)); }  // powf(0.5) lowers to sqrt
__device__ __forceinline__ Color3 fresnel_conductor(Color3 eta_i, Color3 eta_t, Color3 k, double cos_theta_i) {  // :359-382
    const Color3 white = mkc(1.0, 1.0, 1.0);
    // (the one caller passes eta_i = 1: x / 1.0 is x bit for bit, six divisions saved per conductor vertex)
    const bool unit_eta = eta_i.r == 1.0 && eta_i.g == 1.0 && eta_i.b == 1.0;
    const Color3 eta_rel = unit_eta ? eta_t : eta_t / eta_i;
    const Color3 eta_rel_2 = eta_rel * eta_rel;
    const Color3 k_rel = unit_eta ? k : k / eta_i;
    const Color3 k_rel_2 = k_rel * k_rel;
    const double cos_theta_2 = cos_theta_i * cos_theta_i;
    const double sin_theta_2 = 1.0 - cos_theta_2;
    const Color3 t0 = eta_rel_2 - k_rel_2 - white * sin_theta_2;
    const Color3 a2_plus_b2 = csqrt(t0 * t0 + eta_rel_2 * k_rel_2 * 4.0);
    const Color3 a = csqrt((a2_plus_b2 + t0) * 0.5);
    const Color3 t1 = a2_plus_b2 + white * cos_theta_2;
    const Color3 t2 = a * cos_theta_i * 2.0;
    const Color3 r_perpendicular = (t1 - t2) / (t1 + t2);
    const Color3 t3 = a2_plus_b2 * cos_theta_2 + white * sin_theta_2 * sin_theta_2;
    const Color3 t4 = a * cos_theta_i * sin_theta_2 * 2.0;
    const Color3 r_parallel = r_perpendicular * (t3 - t4) / (t3 + t4);
    return (r_parallel * r_parallel + r_perpendicular * r_perpendicular) * 0.5;
}

struct SurfaceSample {  // bxdf.rs:12-19
    V3 w_i;
    Color3 f;
    PdfValue pdf;
    bool is_specular;
};

__device__ __forceinline__ bool lobe_has_reflection(const DevLobe& l) { return l.kind != LOBE_SPECULAR_BTDF; }
__device__ __forceinline__ bool lobe_has_transmission(const DevLobe& l) { return l.kind == LOBE_SPECULAR_BTDF || l.kind == LOBE_FRESNEL_SPECULAR; }

// The shading kernels are instantiated per shade class (k_shade_class, wavefront.cu): LOBES is the set of lobe kinds a material of
// that class can hold (bit k = lobe kind k), so each kernel carries only its own family's code and registers.
__host__ __device__ constexpr uint32_t lobe_bit(uint32_t kind) { return 1u << kind; }
constexpr uint32_t kLobesAll = 0x3Fu;
constexpr uint32_t kLobesDiffuse = lobe_bit(LOBE_LAMBERTIAN) | lobe_bit(LOBE_OREN_NAYAR);
// Material::new_* material.rs:20-70 (make_material, scene_device.cu): matte, glass, plastic, metal
__host__ __device__ constexpr uint32_t lobes_of_class(uint32_t cls) {
    return cls == CRAY_MAT_MATTE ? kLobesDiffuse
         : cls == CRAY_MAT_GLASS ? lobe_bit(LOBE_FRESNEL_SPECULAR)
         : cls == CRAY_MAT_PLASTIC ? (kLobesDiffuse | lobe_bit(LOBE_SPECULAR_BRDF))
         : cls == CRAY_MAT_METAL ? lobe_bit(LOBE_CONDUCTOR) : kLobesAll;
}

// BxDF::f bxdf.rs:214-265
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ Color3 lobe_f_inline(const SceneView& s, const DevLobe& l, V3 w_o, V3 w_i, V3 normal, double tu, double tv) {
    const Color3 black = mkc(0.0, 0.0, 0.0);
    if constexpr (!(LOBES & kLobesDiffuse)) return black;
    if (l.kind == LOBE_LAMBERTIAN) {
        if (same_hemisphere(normal, w_o, w_i)) return eval_color(s, l.t0, tu, tv) * kFrac1Pi;
        return black;
    }
    if (l.kind == LOBE_OREN_NAYAR) {
        if (!same_hemisphere(normal, w_o, w_i)) return black;
        const double cos_theta_i = fabs(dot(w_i, normal));
        const double cos_theta_o = fabs(dot(w_o, normal));
        const double sin_theta_i = sqrt(rmax(1.0 - cos_theta_i * cos_theta_i, 0.0));
        const double sin_theta_o = sqrt(rmax(1.0 - cos_theta_o * cos_theta_o, 0.0));
        double max_cos = 0.0;
        if (sin_theta_i > 1e-4 && sin_theta_o > 1e-4) {
            V3 tangent, bitangent;
            generate_tangents(normal, tangent, bitangent);
            const double cos_phi_i = fabs(dot(w_i, tangent));
            const double cos_phi_o = fabs(dot(w_o, tangent));
            const double sin_phi_i = sqrt(1.0 - cos_phi_i * cos_phi_i);
            const double sin_phi_o = sqrt(1.0 - cos_phi_o * cos_phi_o);
            max_cos = rmax(cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o, 0.0);
        }
        double sin_alpha, tan_beta;
        if (cos_theta_i > cos_theta_o) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / cos_theta_i; }
        else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / cos_theta_o; }
        const double sg = to_radians(eval_f64(s, l.sigma, tu, tv));
        const double sigma_2 = sg * sg;
        const double A = 1.0 - sigma_2 / (2.0 * (sigma_2 + 0.33));
        const double B = 0.45 * sigma_2 / (sigma_2 + 0.09);
        return eval_color(s, l.t0, tu, tv) * (A + B * max_cos * sin_alpha * tan_beta) * kFrac1Pi;
    }
    return black;
}
// Out of line where a class calls it from several places (plastic: two lobes; the generic instantiation), inline where the class has
// one lobe and one call site.
template <uint32_t LOBES>
__device__ __noinline__ Color3 lobe_f_outline(const SceneView& s, const DevLobe& l, V3 w_o, V3 w_i, V3 normal, double tu, double tv) {
    return lobe_f_inline<LOBES>(s, l, w_o, w_i, normal, tu, tv);
}
__host__ __device__ constexpr bool lobes_single_family(uint32_t lobes) { return lobes == kLobesDiffuse || (lobes & (lobes - 1u)) == 0u; }
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ Color3 lobe_f(const SceneView& s, const DevLobe& l, V3 w_o, V3 w_i, V3 normal, double tu, double tv) {
    if constexpr (lobes_single_family(LOBES)) return lobe_f_inline<LOBES>(s, l, w_o, w_i, normal, tu, tv);
    else return lobe_f_outline<LOBES>(s, l, w_o, w_i, normal, tu, tv);
}
// BxDF::pdf bxdf.rs:269-284
__device__ __forceinline__ PdfValue lobe_pdf(const DevLobe& l, V3 w_i, V3 normal) {
    if (l.kind == LOBE_LAMBERTIAN || l.kind == LOBE_OREN_NAYAR) return non_delta(kFrac1Pi * fabs(dot(w_i, normal)));
    return delta_pdf();
}
// BxDF::sample bxdf.rs:83-209
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ bool lobe_sample_inline(const SceneView& s, const DevLobe& l, const VertexSamples& vs, V3 w_o, V3 normal, double tu, double tv,
                                                   SurfaceSample& out, bool& assert_failed) {
    if ((LOBES & kLobesDiffuse) && (l.kind == LOBE_LAMBERTIAN || l.kind == LOBE_OREN_NAYAR)) {
        V3 w_i = cosine_sample_hemisphere(vs.get(VertexSamples::MATERIAL_U), vs.get(VertexSamples::MATERIAL_V), normal, assert_failed);
        if (dot(normal, w_o) < 0.0) w_i = neg(w_i);
        out.w_i = w_i;
        out.f = lobe_f<LOBES>(s, l, w_o, w_i, normal, tu, tv);
        out.pdf = lobe_pdf(l, w_i, normal);
        out.is_specular = false;
        return true;
    }
    if ((LOBES & lobe_bit(LOBE_CONDUCTOR)) && l.kind == LOBE_CONDUCTOR) {
        const V3 w_i = reflect(w_o, normal);
        if (!(fabs(magnitude(w_i) - 1.0) <= kEpsilon)) assert_failed = true;  // assert_abs_diff_eq! bxdf.rs:119
        const double cos_theta_i = fabs(dot(w_o, normal));
        const Color3 fr = fresnel_conductor(mkc(1.0, 1.0, 1.0), eval_color(s, l.t0, tu, tv), eval_color(s, l.t1, tu, tv), cos_theta_i);
        out.w_i = w_i;
        out.f = fr / cos_theta_i;
        out.pdf = delta_pdf();
        out.is_specular = true;
        return true;
    }
    if ((LOBES & lobe_bit(LOBE_SPECULAR_BRDF)) && l.kind == LOBE_SPECULAR_BRDF) {
        const V3 w_i = reflect(w_o, normal);
        if (!(fabs(magnitude(w_i) - 1.0) <= kEpsilon)) assert_failed = true;
        const double cos_theta_i = fabs(dot(w_o, normal));
        const Color3 fr = mkc(1.0, 1.0, 1.0) * fresnel_dielectric(l.eta_i, l.eta_t, cos_theta_i);
        out.w_i = w_i;
        out.f = eval_color(s, l.t0, tu, tv) * fr / fabs(cos_theta_i);
        out.pdf = delta_pdf();
        out.is_specular = true;
        return true;
    }
    if ((LOBES & lobe_bit(LOBE_SPECULAR_BTDF)) && l.kind == LOBE_SPECULAR_BTDF) {
        const double cos_theta_i = fabs(dot(w_o, normal));
        V3 w_i;
        if (!refract(w_o, normal, cos_theta_i, l.eta_i, l.eta_t, w_i)) return false;
        if (!(fabs(magnitude(w_i) - 1.0) <= kEpsilon)) assert_failed = true;
        const double fr = fresnel_dielectric(l.eta_i, l.eta_t, cos_theta_i);
        out.w_i = w_i;
        out.f = eval_color(s, l.t1, tu, tv) * (1.0 - fr) / cos_theta_i;
        out.pdf = delta_pdf();
        out.is_specular = true;
        return true;
    }
    if constexpr ((LOBES & lobe_bit(LOBE_FRESNEL_SPECULAR)) != 0u) {  // LOBE_FRESNEL_SPECULAR bxdf.rs:176-207 (the last kind)
        const double cos_theta_i = dot(w_o, normal);
        const double fresnel_reflectance = fresnel_dielectric(l.eta_i, l.eta_t, cos_theta_i);
        if (vs.get(VertexSamples::MATERIAL_U) < fresnel_reflectance) {
            out.w_i = reflect(w_o, normal);
            out.f = eval_color(s, l.t0, tu, tv) * fresnel_reflectance / fabs(cos_theta_i);
            out.pdf = non_delta(fresnel_reflectance);
            out.is_specular = true;
            return true;
        }
        V3 w_i;
        if (!refract(w_o, normal, cos_theta_i, l.eta_i, l.eta_t, w_i)) return false;
        out.w_i = w_i;
        out.f = eval_color(s, l.t1, tu, tv) * (1.0 - fresnel_reflectance) / fabs(cos_theta_i);
        out.pdf = non_delta(1.0 - fresnel_reflectance);
        out.is_specular = true;
        return true;
    }
    return false;  // (a lobe kind outside LOBES: scene_device.cu never builds one for this class)
}

template <uint32_t LOBES>
__device__ __noinline__ bool lobe_sample_outline(const SceneView& s, const DevLobe& l, const VertexSamples& vs, V3 w_o, V3 normal, double tu, double tv,
                                                 SurfaceSample& out, bool& assert_failed) {
    return lobe_sample_inline<LOBES>(s, l, vs, w_o, normal, tu, tv, out, assert_failed);
}
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ bool lobe_sample(const SceneView& s, const DevLobe& l, const VertexSamples& vs, V3 w_o, V3 normal, double tu, double tv,
                                            SurfaceSample& out, bool& assert_failed) {
    if constexpr (lobes_single_family(LOBES)) return lobe_sample_inline<LOBES>(s, l, vs, w_o, normal, tu, tv, out, assert_failed);
    else return lobe_sample_outline<LOBES>(s, l, vs, w_o, normal, tu, tv, out, assert_failed);
}

__device__ __forceinline__ bool lobe_relevant(const DevLobe& l, bool is_reflecting) {  // bsdf.rs:66-76
    return is_reflecting ? lobe_has_reflection(l) : lobe_has_transmission(l);
}

// Material::f material.rs:84-89, BSDF::f bsdf.rs:79-85
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ Color3 material_f(const SceneView& s, const DevMaterial& m, V3 w_o, V3 w_i, V3 normal, double tu, double tv) {
    if (!m.is_bsdf) return lobe_f<LOBES>(s, m.lobes[0], w_o, w_i, normal, tu, tv);
    Color3 acc = mkc(0.0, 0.0, 0.0);
    const bool is_reflecting = dot(w_o, normal) * dot(w_i, normal) > 0.0;
    for (uint32_t i = 0; i < m.n_lobes; ++i)
        if (lobe_relevant(m.lobes[i], is_reflecting)) acc = acc + lobe_f<LOBES>(s, m.lobes[i], w_o, w_i, normal, tu, tv);
    return acc;
}
// Material::pdf material.rs:90-95, BSDF::pdf bsdf.rs:87-98
__device__ __forceinline__ PdfValue material_pdf(const DevMaterial& m, V3 w_o, V3 w_i, V3 normal) {
    if (!m.is_bsdf) return lobe_pdf(m.lobes[0], w_i, normal);
    double pdf = 0.0;
    int matching = 0;
    const bool is_reflecting = dot(w_o, normal) * dot(w_i, normal) > 0.0;
    for (uint32_t i = 0; i < m.n_lobes; ++i)
        if (lobe_relevant(m.lobes[i], is_reflecting)) {
            const PdfValue p = lobe_pdf(m.lobes[i], w_i, normal);
            if (!p.delta) { pdf += p.value; matching += 1; }
        }
    if (matching > 0) return non_delta(pdf / (double)matching);
    return delta_pdf();
}
// Material::sample material.rs:72-83, BSDF::sample bsdf.rs:15-60
template <uint32_t LOBES = kLobesAll>
__device__ __forceinline__ bool material_sample(const SceneView& s, const DevMaterial& m, const VertexSamples& vs, V3 w_o, V3 normal,
                                                double tu, double tv, SurfaceSample& out, bool& assert_failed) {
    if (!m.is_bsdf) return lobe_sample<LOBES>(s, m.lobes[0], vs, w_o, normal, tu, tv, out, assert_failed);
    if (m.n_lobes == 0) return false;
    // one lobe: floor(u * 1) is 0 for every u in [0, 1)
    const uint32_t sample_index = m.n_lobes == 1 ? 0u : (uint32_t)as_usize(vs.get(VertexSamples::MATERIAL_1D) * (double)m.n_lobes);
    SurfaceSample smp;
    if (!lobe_sample<LOBES>(s, m.lobes[sample_index], vs, w_o, normal, tu, tv, smp, assert_failed)) return false;
    if (!smp.pdf.delta) {
        double pdf = smp.pdf.value;
        Color3 f = smp.f;
        const bool is_reflecting = dot(w_o, normal) * dot(smp.w_i, normal) > 0.0;
        for (uint32_t i = 0; i < m.n_lobes; ++i) {
            if (i == sample_index || !lobe_relevant(m.lobes[i], is_reflecting)) continue;
            f = f + lobe_f<LOBES>(s, m.lobes[i], w_o, smp.w_i, normal, tu, tv);
            const PdfValue op = lobe_pdf(m.lobes[i], smp.w_i, normal);
            if (!op.delta) pdf += op.value;
        }
        out.w_i = smp.w_i;
        out.f = f;
        out.pdf = non_delta(pdf / (double)m.n_lobes);
        out.is_specular = smp.is_specular;
    } else {
        out = smp;
    }
    return true;
}

// ---- lights (light.rs) ----

// Shape::sample shape.rs:445-470
__device__ __forceinline__ V3 shape_sample(const SceneView& s, const LeafPrim& sh, double su, double sv) {
    const uint32_t kind = sh.kind & 0xFFu;
    if (kind == PRIM_SPHERE) {
        const V3 p = mk(0.0, 0.0, 0.0) + sample_sphere(su, sv) * sh.d[3];
        return mk(p.x + sh.d[0], p.y + sh.d[1], p.z + sh.d[2]);  // translate(origin): 1*x + 0*y + 0*z + origin.x
    }
    if (kind == PRIM_TRIANGLE) {
        const double su_sqrt = sqrt(su);  // sample_triangle sampling.rs:51-55
        const double b1 = 1.0 - su_sqrt, b2 = sv * su_sqrt;
        return mk(sh.d[0], sh.d[1], sh.d[2]) + mk(sh.d[3], sh.d[4], sh.d[5]) * b1 + mk(sh.d[6], sh.d[7], sh.d[8]) * b2;
    }
    const DiskXf& k = s.disks[disk_index(sh.kind)];
    double x, y;
    sample_disk(su, sv, x, y);
    return xf_point(k.o2w, mk(x * k.radius, y * k.radius, 0.0));
}

// Shape::pdf_from shape.rs:487-502: intersect the light's own shape from the receiver along w_i
__device__ __forceinline__ double shape_pdf_from(const SceneView& s, const LeafPrim& sh, double area, V3 location, V3 normal, V3 w_i) {
    double ray_max = inf_f64(), u, v;
    if (!leaf_prim_closest(s, sh, location, w_i, ray_max, u, v)) return 0.0;
    V3 hit_location, hit_normal;
    double tu, tv;
    // only the location is needed; recompute it the way Shape::intersect does
    const uint32_t kind = sh.kind & 0xFFu;
    if (kind == PRIM_TRIANGLE) hit_location = location + w_i * ray_max;
    else surface_at(s, sh, location, w_i, ray_max, u, v, hit_location, hit_normal, tu, tv, false);
    const double distance_squared = magnitude_squared(hit_location - location);
    const double cos_theta = fabs(dot(w_i, normal));
    return distance_squared / (cos_theta * area);
}

__device__ __forceinline__ PdfValue light_pdf_li(const SceneView& s, const DevLight& l, V3 location, V3 normal, V3 w_i) {  // light.rs:136-143
    if (l.kind == CRAY_LIGHT_POINT || l.kind == CRAY_LIGHT_DISTANT) return delta_pdf();
    if (l.kind == CRAY_LIGHT_INFINITE) return non_delta(kFrac1Pi / 4.0);
    return non_delta(shape_pdf_from(s, l.shape, l.area, location, normal, w_i));
}

struct LightSample {  // light.rs:45-51
    Color3 Li;
    V3 w_i;
    PdfValue pdf;
    double shadow_max;  // shadow_ray.max_distance (origin = intersection.location, direction = w_i)
};

// Light::sample_Li light.rs:59-133
__device__ __forceinline__ LightSample light_sample_li(const SceneView& s, const DevLight& l, const VertexSamples& vs, V3 location, V3 normal,
                                                       bool& assert_failed) {
    LightSample out;
    const Color3 color = mkc(l.color[0], l.color[1], l.color[2]);
    if (l.kind == CRAY_LIGHT_POINT) {
        const V3 op = mk(l.v[0], l.v[1], l.v[2]) - location;
        const double dist_squared = magnitude_squared(op);
        const double dist = sqrt(dist_squared);
        out.w_i = op / dist;
        out.shadow_max = contains_distance(dist, inf_f64()) ? dist : inf_f64();
        out.Li = color / dist_squared;
        out.pdf = delta_pdf();
    } else if (l.kind == CRAY_LIGHT_DISTANT) {
        const V3 direction = mk(l.v[0], l.v[1], l.v[2]);
        if (!(fabs(magnitude(direction) - 1.0) <= kEpsilon)) assert_failed = true;
        out.w_i = direction;
        out.shadow_max = inf_f64();
        out.Li = color;
        out.pdf = delta_pdf();
    } else if (l.kind == CRAY_LIGHT_INFINITE) {
        const V3 n = vs.get(VertexSamples::LIGHT_1D) < 0.5 ? mk(1.0, 0.0, 0.0) : mk(-1.0, 0.0, 0.0);
        out.w_i = sample_hemisphere(vs.get(VertexSamples::LIGHT_U), vs.get(VertexSamples::LIGHT_V), n);
        out.shadow_max = inf_f64();
        out.Li = color;
        out.pdf = non_delta(kFrac1Pi / 4.0);
    } else {
        // Shape::sample_from shape.rs:472-484
        const V3 shape_point = shape_sample(s, l.shape, vs.get(VertexSamples::LIGHT_U), vs.get(VertexSamples::LIGHT_V));
        out.w_i = normalized(shape_point - location);
        out.pdf = non_delta(shape_pdf_from(s, l.shape, l.area, location, normal, out.w_i));
        const double distance = magnitude(shape_point - location);
        const double m = distance - kEpsilon;
        out.shadow_max = contains_distance(m, inf_f64()) ? m : inf_f64();
        out.Li = color;
    }
    return out;
}

// LightSampler::sample light.rs:203-211 (binary search; Err(i) is the insertion point)
__device__ __forceinline__ uint32_t light_pick(const SceneView& s, double u, double& pdf) {
    uint32_t lo = 0, hi = s.n_lights;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (s.light_cdf[mid] < u) lo = mid + 1;
        else hi = mid;
    }
    pdf = lo > 0 ? s.light_cdf[lo] - s.light_cdf[lo - 1] : s.light_cdf[lo];
    return lo;
}
__device__ __forceinline__ double light_pick_pdf(const SceneView& s, uint32_t i) {  // light.rs:213-219
    return i > 0 ? s.light_cdf[i] - s.light_cdf[i - 1] : s.light_cdf[i];
}

}  // namespace cray
