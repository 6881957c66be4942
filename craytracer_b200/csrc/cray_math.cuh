// f64 vector / matrix / ray / bounds arithmetic shared by the host scene builder and the sm_100a kernels.
//
// Parity rule: every expression below keeps the operand order of the reference (src/geometry.rs,
// src/transformation.rs, src/bounds.rs, src/ray.rs) and the whole library is compiled with
// `--fmad=false` (device) and `-ffp-contract=off` (host), so + - * / sqrt results are bit-identical
// to the Rust f64 code.  f32 traversal code that wants FMAs asks for them explicitly (fmaf).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CRAY_HD __host__ __device__ __forceinline__
#else
#define CRAY_HD inline
#endif

namespace cray {

constexpr double kEpsilon = 1e-9;  // src/constants.rs:1
constexpr double kPi = 3.14159265358979323846;
constexpr double kFrac1Pi = 0.318309886183790671537767526745028724;
constexpr double kFracPi2 = 1.57079632679489661923132169163975144;
constexpr double kFracPi4 = 0.785398163397448309615660845819875721;

CRAY_HD double inf_f64() {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff0000000000000LL);
#else
    return HUGE_VAL;
#endif
}
// Rust f64::min/max return the non-NaN operand; fmin/fmax have the same rule.
CRAY_HD double rmin(double a, double b) { return fmin(a, b); }
CRAY_HD double rmax(double a, double b) { return fmax(a, b); }
CRAY_HD bool sign_negative(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x) < 0;
#else
    return std::signbit(x);
#endif
}
CRAY_HD double to_radians(double deg) { return deg * (kPi / 180.0); }
// Rust `x as usize` / `as u32`: saturating, NaN -> 0
CRAY_HD uint64_t as_usize(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)x;
}
CRAY_HD uint32_t as_u32(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 4294967296.0) return 0xFFFFFFFFu;
    return (uint32_t)x;
}

// f64 division and square root expand to ~25 instructions each.  The shading code has hundreds of call sites, and its
// instruction-cache misses cost more than a call does, so the vector / colour operators below go through ONE out-of-line copy
// on the device (same IEEE result).  The traversal kernels divide scalars directly and stay inline.
#if defined(__CUDA_ARCH__) && !defined(CRAY_INLINE_DIV)
static __device__ __noinline__ double div_rn(double a, double b) { return a / b; }
static __device__ __noinline__ double sqrt_rn(double a) { return sqrt(a); }
#else
CRAY_HD double div_rn(double a, double b) { return a / b; }
CRAY_HD double sqrt_rn(double a) { return sqrt(a); }
#endif

struct V3 {
    double x, y, z;
    CRAY_HD double operator[](int a) const { return a == 0 ? x : (a == 1 ? y : z); }
};
CRAY_HD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CRAY_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
CRAY_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
CRAY_HD V3 operator*(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
CRAY_HD V3 operator/(V3 a, double s) { return mk(div_rn(a.x, s), div_rn(a.y, s), div_rn(a.z, s)); }
CRAY_HD V3 neg(V3 a) { return a * -1.0; }  // Neg is `self * -1.0` (geometry.rs:155)
CRAY_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CRAY_HD double magnitude_squared(V3 a) { return dot(a, a); }
CRAY_HD double magnitude(V3 a) { return sqrt_rn(magnitude_squared(a)); }
CRAY_HD V3 normalized(V3 a) {
    double mag = magnitude(a);
    return mk(div_rn(a.x, mag), div_rn(a.y, mag), div_rn(a.z, mag));
}
CRAY_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CRAY_HD bool same_hemisphere(V3 n, V3 v1, V3 v2) { return dot(n, v1) * dot(n, v2) > 0.0; }  // geometry.rs:403
// Normal::generate_tangents geometry.rs:406-417
CRAY_HD void generate_tangents(V3 n, V3& t, V3& b) {
    V3 v = normalized(n);
    double sign = copysign(1.0, v.z);
    double a = -1.0 / (sign + v.z);
    double bb = v.x * v.y * a;
    t = mk(1.0 + sign * v.x * v.x * a, sign * bb, -sign * v.x);
    b = mk(bb, sign + v.y * v.y * a, -v.y);
}

struct Color3 {
    double r, g, b;
};
CRAY_HD Color3 mkc(double r, double g, double b) { Color3 c; c.r = r; c.g = g; c.b = b; return c; }
CRAY_HD Color3 operator+(Color3 a, Color3 c) { return mkc(a.r + c.r, a.g + c.g, a.b + c.b); }
CRAY_HD Color3 operator-(Color3 a, Color3 c) { return mkc(a.r - c.r, a.g - c.g, a.b - c.b); }
CRAY_HD Color3 operator*(Color3 a, Color3 c) { return mkc(a.r * c.r, a.g * c.g, a.b * c.b); }
CRAY_HD Color3 operator/(Color3 a, Color3 c) { return mkc(div_rn(a.r, c.r), div_rn(a.g, c.g), div_rn(a.b, c.b)); }
CRAY_HD Color3 operator*(Color3 a, double s) { return mkc(a.r * s, a.g * s, a.b * s); }
CRAY_HD Color3 operator/(Color3 a, double s) { return mkc(div_rn(a.r, s), div_rn(a.g, s), div_rn(a.b, s)); }
CRAY_HD bool is_black(Color3 c) { return c.r == 0.0 && c.g == 0.0 && c.b == 0.0; }
CRAY_HD bool finite_f64(double x) { return fabs(x) <= 1.7976931348623157e308; }  // false for +-inf and NaN
CRAY_HD bool is_finite3(Color3 c) { return finite_f64(c.r) && finite_f64(c.g) && finite_f64(c.b); }

// Row-major 3x4 affine part of the reference's 4x4 matrices.  Every transformation the hot path
// applies at run time (sphere translate, disk translate*rotate, look_at) has bottom row exactly
// (0,0,0,1), so the homogeneous divide of transformation.rs:431 is a division by exactly 1.0.
struct Affine {
    double m[3][4];
};
CRAY_HD V3 xf_point(const Affine& t, V3 p) {  // transformation.rs:421-432
    return mk(t.m[0][0] * p.x + t.m[0][1] * p.y + t.m[0][2] * p.z + t.m[0][3],
              t.m[1][0] * p.x + t.m[1][1] * p.y + t.m[1][2] * p.z + t.m[1][3],
              t.m[2][0] * p.x + t.m[2][1] * p.y + t.m[2][2] * p.z + t.m[2][3]);
}
CRAY_HD V3 xf_vector(const Affine& t, V3 v) {  // transformation.rs:434-444
    return mk(t.m[0][0] * v.x + t.m[0][1] * v.y + t.m[0][2] * v.z, t.m[1][0] * v.x + t.m[1][1] * v.y + t.m[1][2] * v.z,
              t.m[2][0] * v.x + t.m[2][1] * v.y + t.m[2][2] * v.z);
}
// Normal transform by the inverse transpose: pass the INVERSE matrix (transformation.rs:446-457)
CRAY_HD V3 xf_normal_with_inverse(const Affine& inv, V3 n) {
    return mk(inv.m[0][0] * n.x + inv.m[1][0] * n.y + inv.m[2][0] * n.z, inv.m[0][1] * n.x + inv.m[1][1] * n.y + inv.m[2][1] * n.z,
              inv.m[0][2] * n.x + inv.m[1][2] * n.y + inv.m[2][2] * n.z);
}

// Ray::contains_distance ray.rs:26
CRAY_HD bool contains_distance(double t, double max_distance) { return t > kEpsilon && t < max_distance; }
// max_distance of `transformation.transform(ray)` (transformation.rs:459-466): Ray::new then update_max_distance
CRAY_HD double transformed_max_distance(double max_distance) {
    return contains_distance(max_distance, inf_f64()) ? max_distance : inf_f64();
}

struct Box3 {
    V3 lo, hi;
};
// Bounds::intersects bounds.rs:62-88 -- true divisions, the reference's accept rule.
CRAY_HD bool bounds_intersects(const Box3& b, V3 o, V3 d, double ray_max) {
    double min_distance = -inf_f64();
    double max_distance = inf_f64();
    for (int axis = 0; axis < 3; ++axis) {
        double d_i = d[axis], o_i = o[axis];
        double min_i = b.lo[axis], max_i = b.hi[axis];
        if (sign_negative(d_i)) { double t = min_i; min_i = max_i; max_i = t; }
        max_distance = rmin(max_distance, (max_i - o_i) / d_i);
        if (max_distance < kEpsilon) return false;
        min_distance = rmax(min_distance, (min_i - o_i) / d_i);
        if (min_distance > max_distance) return false;
    }
    return contains_distance(min_distance, ray_max) || contains_distance(max_distance, ray_max);
}
CRAY_HD bool bounds_contains(const Box3& b, V3 p) {  // bounds.rs:46-53
    return b.lo.x <= p.x && b.lo.y <= p.y && b.lo.z <= p.z && b.hi.x >= p.x && b.hi.y >= p.y && b.hi.z >= p.z;
}

}  // namespace cray
