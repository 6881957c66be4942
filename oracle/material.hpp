// ORACLE — TEST INFRASTRUCTURE ONLY (see geometry.hpp).
// Restatement of src/texture.rs, src/bxdf.rs, src/bsdf.rs, src/material.rs, src/pdf.rs.
#pragma once
#include <vector>
#include "geometry.hpp"
#include "sampling.hpp"

namespace orc {

struct Pdf {  // pdf.rs:1-6
    bool delta;
    double value;
    static Pdf NonDelta(double v) { return {false, v}; }
    static Pdf Delta() { return {true, 0.0}; }
};

struct Image {
    uint32_t width = 0, height = 0;
    const uint8_t* rgb = nullptr;
};

enum TexKind { TEX_CONSTANT = 0, TEX_CHECKER = 1, TEX_IMAGE = 2 };

// Texture<Color> and Texture<f64> (texture.rs:8-12).  A Texture<f64> uses a.r / b.r.
struct Texture {
    int kind = TEX_CONSTANT;
    Color a{0, 0, 0}, b{0, 0, 0};
    double scale = 1.0;
    const Image* image = nullptr;

    void texel(double u, double v, const uint8_t*& px) const {  // texture.rs:31-45
        double uu = u - std::trunc(u);  // f64::fract
        if (uu < 0.0) uu += 1.0;
        double vv = v - std::trunc(v);
        if (vv < 0.0) vv += 1.0;
        uint32_t x = as_u32((double)(image->width - 1) * uu);
        uint32_t y = as_u32((double)(image->height - 1) * vv);
        px = image->rgb + 3 * ((size_t)y * image->width + x);
    }
    bool checker_is_a(double u, double v) const {  // texture.rs:22-30
        uint64_t iu = as_usize(u * scale * 2.0);
        uint64_t iv = as_usize(v * scale * 2.0);
        return ((iu & 1) ^ (iv & 1)) == 0;
    }
    Color eval_color(const double uv[2]) const {  // texture.rs:19-47, FromPixel for Color :108-112
        switch (kind) {
            case TEX_CONSTANT: return a;
            case TEX_CHECKER: return checker_is_a(uv[0], uv[1]) ? a : b;
            default: {
                const uint8_t* px;
                texel(uv[0], uv[1], px);
                return from_rgb(px[0], px[1], px[2]);
            }
        }
    }
    double eval_f64(const double uv[2]) const {  // FromPixel for f64  texture.rs:102-106
        switch (kind) {
            case TEX_CONSTANT: return a.r;
            case TEX_CHECKER: return checker_is_a(uv[0], uv[1]) ? a.r : b.r;
            default: {
                const uint8_t* px;
                texel(uv[0], uv[1], px);
                // image::Rgb<u8>::to_luma: (2126*r + 7152*g + 722*b) / 10000, integer arithmetic (image 0.24 color.rs)
                uint32_t l = (2126u * px[0] + 7152u * px[1] + 722u * px[2]) / 10000u;
                return (double)l / 255.0;
            }
        }
    }
    bool is_black() const {  // texture.rs:82-90
        switch (kind) {
            case TEX_CONSTANT: return orc::is_black(a);
            case TEX_CHECKER: return orc::is_black(a) && orc::is_black(b);
            default: return false;
        }
    }
    bool is_zero() const {  // texture.rs:92-100
        switch (kind) {
            case TEX_CONSTANT: return a.r == 0.0;
            case TEX_CHECKER: return a.r == 0.0 && b.r == 0.0;
            default: return false;
        }
    }
};

// ---- bxdf.rs helpers ---------------------------------------------------------------
inline V3 reflect(V3 direction, V3 normal) {  // bxdf.rs:287-290
    return normal * (dot(normal, direction) * 2.0) - direction;
}
inline bool refract(V3 direction, V3 normal, double cos_theta_i, double eta_i, double eta_t, V3& out) {  // bxdf.rs:292-314
    double eta_relative, cos_theta;
    if (std::signbit(cos_theta_i)) {
        normal = neg(normal);
        eta_relative = eta_i / eta_t;
        cos_theta = -cos_theta_i;
    } else {
        eta_relative = eta_t / eta_i;
        cos_theta = cos_theta_i;
    }
    double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    if (sin_theta > eta_relative) return false;
    V3 r_perpendicular = (normal * cos_theta - direction) / eta_relative;
    V3 r_parallel = normal * -std::sqrt(1.0 - dot(r_perpendicular, r_perpendicular));
    out = r_perpendicular + r_parallel;
    return true;
}
inline double fresnel_dielectric(double eta_i, double eta_t, double cos_theta_i) {  // bxdf.rs:338-357
    if (std::signbit(cos_theta_i)) {
        cos_theta_i = -cos_theta_i;
        std::swap(eta_i, eta_t);
    }
    double sin_theta_i = std::sqrt(1.0 - cos_theta_i * cos_theta_i);
    double sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0) return 1.0;
    double cos_theta_t = std::sqrt(1.0 - sin_theta_t * sin_theta_t);
    double r_parallel = (eta_t * cos_theta_i - eta_i * cos_theta_t) / (eta_t * cos_theta_i + eta_i * cos_theta_t);
    double r_perpendicular = (eta_i * cos_theta_i - eta_t * cos_theta_t) / (eta_i * cos_theta_i + eta_t * cos_theta_t);
    return (r_parallel * r_parallel + r_perpendicular * r_perpendicular) * 0.5;
}
inline Color csqrt(Color c) { return {std::sqrt(c.r), std::sqrt(c.g), std::sqrt(c.b)}; }  // powf(0.5): LLVM lowers pow(x,0.5) to sqrt
inline Color fresnel_conductor(Color eta_i, Color eta_t, Color k, double cos_theta_i) {  // bxdf.rs:359-382
    Color eta_rel = eta_t / eta_i;
    Color eta_rel_2 = eta_rel * eta_rel;
    Color k_rel = k / eta_i;
    Color k_rel_2 = k_rel * k_rel;
    double cos_theta_2 = cos_theta_i * cos_theta_i;
    double sin_theta_2 = 1.0 - cos_theta_2;
    Color t0 = eta_rel_2 - k_rel_2 - WHITE * sin_theta_2;
    Color a2_plus_b2 = csqrt(t0 * t0 + eta_rel_2 * k_rel_2 * 4.0);
    Color a = csqrt((a2_plus_b2 + t0) * 0.5);
    Color t1 = a2_plus_b2 + WHITE * cos_theta_2;
    Color t2 = a * cos_theta_i * 2.0;
    Color r_perpendicular = (t1 - t2) / (t1 + t2);
    Color t3 = a2_plus_b2 * cos_theta_2 + WHITE * sin_theta_2 * sin_theta_2;
    Color t4 = a * cos_theta_i * sin_theta_2 * 2.0;
    Color r_parallel = r_perpendicular * (t3 - t4) / (t3 + t4);
    return (r_parallel * r_parallel + r_perpendicular * r_perpendicular) * 0.5;
}

struct SurfaceSample {  // bxdf.rs:12-19
    V3 w_i;
    Color f;
    Pdf pdf;
    bool is_specular;
};

enum BxdfKind { LAMBERTIAN = 0, OREN_NAYAR = 1, FRESNEL_CONDUCTOR = 2, SPECULAR_BRDF = 3, SPECULAR_BTDF = 4, FRESNEL_SPECULAR = 5 };

struct BxDF {  // bxdf.rs:23-50
    int kind;
    Texture t0, t1;   // reflectance | eta ; transmittance | k
    Texture sigma;    // OrenNayar
    double eta_i = 1.0, eta_t = 1.0;  // SpecularBRDF dielectric (1, 1.5), FresnelSpecular (1, eta)

    bool has_reflection() const { return kind != SPECULAR_BTDF; }  // bxdf.rs:57-66
    bool has_transmission() const { return kind == SPECULAR_BTDF || kind == FRESNEL_SPECULAR; }  // :68-77

    Color f(V3 w_o, V3 w_i, V3 normal, const double uv[2]) const {  // bxdf.rs:214-265
        switch (kind) {
            case LAMBERTIAN:
                if (same_hemisphere(normal, w_o, w_i)) return t0.eval_color(uv) * FRAC_1_PI;
                return BLACK;
            case OREN_NAYAR: {
                if (!same_hemisphere(normal, w_o, w_i)) return BLACK;
                double cos_theta_i = std::fabs(dot(w_i, normal));
                double cos_theta_o = std::fabs(dot(w_o, normal));
                double sin_theta_i = std::sqrt(rmax(1.0 - cos_theta_i * cos_theta_i, 0.0));
                double sin_theta_o = std::sqrt(rmax(1.0 - cos_theta_o * cos_theta_o, 0.0));
                double max_cos;
                if (sin_theta_i > 1e-4 && sin_theta_o > 1e-4) {
                    V3 tangent, bitangent;
                    generate_tangents(normal, tangent, bitangent);
                    double cos_phi_i = std::fabs(dot(w_i, tangent));
                    double cos_phi_o = std::fabs(dot(w_o, tangent));
                    double sin_phi_i = std::sqrt(1.0 - cos_phi_i * cos_phi_i);
                    double sin_phi_o = std::sqrt(1.0 - cos_phi_o * cos_phi_o);
                    max_cos = rmax(cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o, 0.0);
                } else {
                    max_cos = 0.0;
                }
                double sin_alpha, tan_beta;
                if (cos_theta_i > cos_theta_o) { sin_alpha = sin_theta_o; tan_beta = sin_theta_i / cos_theta_i; }
                else { sin_alpha = sin_theta_i; tan_beta = sin_theta_o / cos_theta_o; }
                double sg = to_radians(sigma.eval_f64(uv));
                double sigma_2 = sg * sg;
                double A = 1.0 - sigma_2 / (2.0 * (sigma_2 + 0.33));
                double B = 0.45 * sigma_2 / (sigma_2 + 0.09);
                return t0.eval_color(uv) * (A + B * max_cos * sin_alpha * tan_beta) * FRAC_1_PI;
            }
            default: return BLACK;
        }
    }
    Pdf pdf(V3 /*w_o*/, V3 w_i, V3 normal) const {  // bxdf.rs:269-284
        if (kind == LAMBERTIAN || kind == OREN_NAYAR) {
            double cos_theta = std::fabs(dot(w_i, normal));
            return Pdf::NonDelta(FRAC_1_PI * cos_theta);
        }
        return Pdf::Delta();
    }
    // bxdf.rs:83-209.  `assert_failed` records the reference's assert!/assert_abs_diff_eq! panics.
    bool sample(double su, double sv, V3 w_o, V3 normal, const double uv[2], SurfaceSample& out, bool* assert_failed) const {
        switch (kind) {
            case LAMBERTIAN:
            case OREN_NAYAR: {
                V3 w_i = cosine_sample_hemisphere(su, sv, normal, assert_failed);
                if (dot(normal, w_o) < 0.0) w_i = neg(w_i);
                out = {w_i, f(w_o, w_i, normal, uv), pdf(w_o, w_i, normal), false};
                return true;
            }
            case FRESNEL_CONDUCTOR: {
                V3 w_i = reflect(w_o, normal);
                if (assert_failed && !(std::fabs(magnitude(w_i) - 1.0) <= EPSILON)) *assert_failed = true;
                double cos_theta_i = std::fabs(dot(w_o, normal));
                Color fr = fresnel_conductor(WHITE, t0.eval_color(uv), t1.eval_color(uv), cos_theta_i);
                out = {w_i, fr / cos_theta_i, pdf(w_o, w_i, normal), true};
                return true;
            }
            case SPECULAR_BRDF: {
                V3 w_i = reflect(w_o, normal);
                if (assert_failed && !(std::fabs(magnitude(w_i) - 1.0) <= EPSILON)) *assert_failed = true;
                double cos_theta_i = std::fabs(dot(w_o, normal));
                Color fr = WHITE * fresnel_dielectric(eta_i, eta_t, cos_theta_i);
                out = {w_i, t0.eval_color(uv) * fr / std::fabs(cos_theta_i), pdf(w_o, w_i, normal), true};
                return true;
            }
            case SPECULAR_BTDF: {
                double cos_theta_i = std::fabs(dot(w_o, normal));
                V3 w_i;
                if (!refract(w_o, normal, cos_theta_i, eta_i, eta_t, w_i)) return false;
                double fr = fresnel_dielectric(eta_i, eta_t, cos_theta_i);
                out = {w_i, t1.eval_color(uv) * (1.0 - fr) / cos_theta_i, pdf(w_o, w_i, normal), true};
                return true;
            }
            default: {  // FRESNEL_SPECULAR  bxdf.rs:176-207
                double cos_theta_i = dot(w_o, normal);
                double fresnel_reflectance = fresnel_dielectric(eta_i, eta_t, cos_theta_i);
                if (su < fresnel_reflectance) {
                    out = {reflect(w_o, normal), t0.eval_color(uv) * fresnel_reflectance / std::fabs(cos_theta_i),
                           Pdf::NonDelta(fresnel_reflectance), true};
                    return true;
                }
                V3 w_i;
                if (!refract(w_o, normal, cos_theta_i, eta_i, eta_t, w_i)) return false;
                out = {w_i, t1.eval_color(uv) * (1.0 - fresnel_reflectance) / std::fabs(cos_theta_i),
                       Pdf::NonDelta(1.0 - fresnel_reflectance), true};
                return true;
            }
        }
    }
};

struct Material {  // material.rs:13-17; `is_bsdf == false` <=> Material::BxDF(bxdfs[0])
    bool is_bsdf = false;
    std::vector<BxDF> bxdfs;

    static Material new_matte(const Texture& reflectance, const Texture& sigma) {  // material.rs:20-26
        Material m;
        BxDF b{};
        b.t0 = reflectance;
        if (sigma.is_zero()) b.kind = LAMBERTIAN;
        else { b.kind = OREN_NAYAR; b.sigma = sigma; }
        m.bxdfs.push_back(b);
        return m;
    }
    static Material new_glass(const Texture& reflectance, const Texture& transmittance, double eta) {  // :27-38
        Material m;
        BxDF b{};
        b.kind = FRESNEL_SPECULAR;
        b.t0 = reflectance; b.t1 = transmittance; b.eta_i = 1.0; b.eta_t = eta;
        m.bxdfs.push_back(b);
        return m;
    }
    static Material new_plastic(const Texture& diffuse, const Texture& specular, const Texture& roughness) {  // :39-64
        Material m;
        m.is_bsdf = true;
        if (!diffuse.is_black()) {
            BxDF b{};
            b.t0 = diffuse;
            if (!roughness.is_zero()) { b.kind = OREN_NAYAR; b.sigma = roughness; }
            else b.kind = LAMBERTIAN;
            m.bxdfs.push_back(b);
        }
        if (!specular.is_black()) {
            BxDF b{};
            b.kind = SPECULAR_BRDF;
            b.t0 = specular; b.eta_i = 1.0; b.eta_t = 1.5;
            m.bxdfs.push_back(b);
        }
        return m;
    }
    static Material new_metal(const Texture& eta, const Texture& k) {  // :65-70
        Material m;
        m.is_bsdf = true;
        BxDF b{};
        b.kind = FRESNEL_CONDUCTOR;
        b.t0 = eta; b.t1 = k;
        m.bxdfs.push_back(b);
        return m;
    }

    template <class F>
    void for_each_relevant(V3 w_o, V3 w_i, V3 normal, F&& fn) const {  // bsdf.rs:62-77
        bool is_reflecting = dot(w_o, normal) * dot(w_i, normal) > 0.0;
        for (size_t idx = 0; idx < bxdfs.size(); ++idx) {
            bool relevant = is_reflecting ? bxdfs[idx].has_reflection() : bxdfs[idx].has_transmission();
            if (relevant) fn(idx, bxdfs[idx]);
        }
    }

    // material.rs:72-83, bsdf.rs:15-60
    bool sample(double s1, double s2u, double s2v, V3 w_o, V3 normal, const double uv[2], SurfaceSample& out, bool* assert_failed) const {
        if (!is_bsdf) return bxdfs[0].sample(s2u, s2v, w_o, normal, uv, out, assert_failed);
        if (bxdfs.empty()) return false;
        size_t sample_index = (size_t)as_usize(s1 * (double)bxdfs.size());
        const BxDF& bxdf = bxdfs[sample_index];
        SurfaceSample s;
        if (!bxdf.sample(s2u, s2v, w_o, normal, uv, s, assert_failed)) return false;
        if (!s.pdf.delta) {
            double pdf = s.pdf.value;
            Color f = s.f;
            for_each_relevant(w_o, s.w_i, normal, [&](size_t other_idx, const BxDF& other) {
                if (other_idx != sample_index) {
                    f += other.f(w_o, s.w_i, normal, uv);
                    Pdf op = other.pdf(w_o, s.w_i, normal);
                    if (!op.delta) pdf += op.value;
                }
            });
            out = {s.w_i, f, Pdf::NonDelta(pdf / (double)bxdfs.size()), s.is_specular};
        } else {
            out = s;
        }
        return true;
    }
    Color f(V3 w_o, V3 w_i, V3 normal, const double uv[2]) const {  // material.rs:84-89, bsdf.rs:79-85
        if (!is_bsdf) return bxdfs[0].f(w_o, w_i, normal, uv);
        Color acc = BLACK;
        for_each_relevant(w_o, w_i, normal, [&](size_t, const BxDF& b) { acc += b.f(w_o, w_i, normal, uv); });
        return acc;
    }
    Pdf pdf(V3 w_o, V3 w_i, V3 normal) const {  // material.rs:90-95, bsdf.rs:87-98
        if (!is_bsdf) return bxdfs[0].pdf(w_o, w_i, normal);
        double pdf = 0.0;
        int num_matching = 0;
        for_each_relevant(w_o, w_i, normal, [&](size_t, const BxDF& b) {
            Pdf p = b.pdf(w_o, w_i, normal);
            if (!p.delta) { pdf += p.value; num_matching += 1; }
        });
        if (num_matching > 0) return Pdf::NonDelta(pdf / (double)num_matching);
        return Pdf::Delta();
    }
};

}  // namespace orc
