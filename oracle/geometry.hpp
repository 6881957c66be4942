// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product: nothing under
// craytracer_b200/ may include, link or call anything in oracle/.
//
// CPU restatement (f64, no FMA contraction: build with -ffp-contract=off) of the
// reference's math layer: src/geometry.rs, src/transformation.rs, src/color.rs,
// src/ray.rs, src/bounds.rs, src/constants.rs.  Operation order is kept exactly as
// in the Rust source so that results are bit-identical to the reference for
// everything that only uses + - * / sqrt.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <algorithm>

namespace orc {

constexpr double EPSILON = 1e-9;               // src/constants.rs:1
constexpr double PI = 3.14159265358979323846;  // std::f64::consts::PI
constexpr double FRAC_1_PI = 0.318309886183790671537767526745028724;
constexpr double FRAC_PI_2 = 1.57079632679489661923132169163975144;
constexpr double FRAC_PI_4 = 0.785398163397448309615660845819875721;
constexpr double INF = std::numeric_limits<double>::infinity();

// Rust f64::min / f64::max ignore NaN (return the other operand): fmin/fmax.
inline double rmin(double a, double b) { return std::fmin(a, b); }
inline double rmax(double a, double b) { return std::fmax(a, b); }
// Rust `x as usize`: saturating, NaN -> 0.
inline uint64_t as_usize(double x) {
    if (!(x > 0.0)) return 0;  // negative, -0, NaN
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
inline uint32_t as_u32(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 4294967296.0) return UINT32_MAX;
    return (uint32_t)x;
}
inline double to_radians(double deg) { return deg * (PI / 180.0); }  // f64::to_radians

// Vector / Point / Normal share one representation (src/geometry.rs:19, :207, :367);
// the reference's three types have identical component-wise arithmetic.
struct V3 {
    double x, y, z;
    double operator[](int a) const { return a == 0 ? x : (a == 1 ? y : z); }
    double& at(int a) { return a == 0 ? x : (a == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return a * -1.0; }  // Neg = self * -1.0, geometry.rs:155
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // geometry.rs:560
inline double magnitude_squared(V3 a) { return dot(a, a); }
inline double magnitude(V3 a) { return std::sqrt(magnitude_squared(a)); }
inline V3 normalized(V3 a) {  // geometry.rs:50-53
    double mag = magnitude(a);
    return {a.x / mag, a.y / mag, a.z / mag};
}
inline V3 cross(V3 a, V3 b) {  // geometry.rs:54-60
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 vmin(V3 a, V3 b) { return {rmin(a.x, b.x), rmin(a.y, b.y), rmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return {rmax(a.x, b.x), rmax(a.y, b.y), rmax(a.z, b.z)}; }
// Normal::same_hemisphere  geometry.rs:403-405
inline bool same_hemisphere(V3 n, V3 v1, V3 v2) { return dot(n, v1) * dot(n, v2) > 0.0; }
// Normal::generate_tangents  geometry.rs:406-417 (pbrt-v4 branchless ONB; signum(+-0) = +-1)
inline void generate_tangents(V3 n, V3& t, V3& b) {
    V3 v = normalized(n);
    double sign = std::copysign(1.0, v.z);
    double a = -1.0 / (sign + v.z);
    double bb = v.x * v.y * a;
    t = {1.0 + sign * v.x * v.x * a, sign * bb, -sign * v.x};
    b = {bb, sign + v.y * v.y * a, -v.y};
}

// ---- Color  src/color.rs ---------------------------------------------------------
struct Color {
    double r, g, b;
};
inline Color operator+(Color a, Color c) { return {a.r + c.r, a.g + c.g, a.b + c.b}; }
inline Color operator-(Color a, Color c) { return {a.r - c.r, a.g - c.g, a.b - c.b}; }
inline Color operator*(Color a, Color c) { return {a.r * c.r, a.g * c.g, a.b * c.b}; }
inline Color operator/(Color a, Color c) { return {a.r / c.r, a.g / c.g, a.b / c.b}; }
inline Color operator*(Color a, double s) { return {a.r * s, a.g * s, a.b * s}; }
inline Color operator/(Color a, double s) { return {a.r / s, a.g / s, a.b / s}; }
inline Color& operator+=(Color& a, Color c) {
    a.r += c.r; a.g += c.g; a.b += c.b;
    return a;
}
constexpr Color BLACK{0.0, 0.0, 0.0};
constexpr Color WHITE{1.0, 1.0, 1.0};
inline bool is_black(Color c) { return c.r == 0.0 && c.g == 0.0 && c.b == 0.0; }  // color.rs:55
inline bool is_finite(Color c) { return std::isfinite(c.r) && std::isfinite(c.g) && std::isfinite(c.b); }
inline Color cpowf(Color c, double p) { return {std::pow(c.r, p), std::pow(c.g, p), std::pow(c.b, p)}; }
constexpr double GAMMA = 2.2;
inline Color from_rgb(uint8_t r, uint8_t g, uint8_t b) {  // color.rs:39-46
    return cpowf(Color{(double)r / 255.0, (double)g / 255.0, (double)b / 255.0}, GAMMA);
}
inline void to_rgb(Color c, uint8_t out[3]) {  // color.rs:47-54 (`as u8` saturates, NaN -> 0)
    Color p = cpowf(c, 1.0 / GAMMA);
    auto conv = [](double v) -> uint8_t {
        double cl = v;  // f64::clamp(0,1): NaN stays NaN
        if (cl < 0.0) cl = 0.0;
        if (cl > 1.0) cl = 1.0;
        double s = cl * 255.0;
        if (!(s > 0.0)) return 0;
        if (s >= 255.0) return 255;
        return (uint8_t)s;
    };
    out[0] = conv(p.r); out[1] = conv(p.g); out[2] = conv(p.b);
}

// ---- Ray  src/ray.rs -------------------------------------------------------------
struct Ray {
    V3 origin, direction;
    double max_distance;
    static Ray make(V3 o, V3 d) { return {o, d, INF}; }                       // ray.rs:14-20
    V3 at(double t) const { return origin + direction * t; }                    // ray.rs:22
    bool contains_distance(double t) const { return t > EPSILON && t < max_distance; }  // ray.rs:26
    bool update_max_distance(double t) {                                        // ray.rs:30-37
        if (contains_distance(t)) { max_distance = t; return true; }
        return false;
    }
};

// ---- Bounds  src/bounds.rs ---------------------------------------------------------
struct Bounds {
    V3 min, max;
    static Bounds make(V3 a, V3 b) { return {vmin(a, b), vmax(a, b)}; }  // bounds.rs:16-21
    V3 centroid() const {                                                 // bounds.rs:22-28
        return {(min.x + max.x) * 0.5, (min.y + max.y) * 0.5, (min.z + max.z) * 0.5};
    }
    V3 diagonal() const { return max - min; }
    double surface_area() const {  // bounds.rs:29-32
        V3 d = diagonal();
        return 2.0 * (d.x * d.y + d.y * d.z + d.z * d.x);
    }
    int maximum_extent() const {  // bounds.rs:36-45
        V3 d = diagonal();
        if (d.x > d.y && d.x > d.z) return 0;
        else if (d.y > d.z) return 1;
        else return 2;
    }
    bool contains(V3 p) const {  // bounds.rs:46-53
        return min.x <= p.x && min.y <= p.y && min.z <= p.z && max.x >= p.x && max.y >= p.y && max.z >= p.z;
    }
    V3 offset(V3 p) const {  // bounds.rs:55-61
        return {(p.x - min.x) / (max.x - min.x), (p.y - min.y) / (max.y - min.y), (p.z - min.z) / (max.z - min.z)};
    }
    bool intersects(const Ray& ray) const {  // bounds.rs:62-88
        double min_distance = -INF;
        double max_distance = INF;
        for (int axis = 0; axis < 3; ++axis) {
            double d_i = ray.direction[axis];
            double o_i = ray.origin[axis];
            double min_i = min[axis];
            double max_i = max[axis];
            if (std::signbit(d_i)) std::swap(min_i, max_i);
            max_distance = rmin(max_distance, (max_i - o_i) / d_i);
            if (max_distance < EPSILON) return false;
            min_distance = rmax(min_distance, (min_i - o_i) / d_i);
            if (min_distance > max_distance) return false;
        }
        return ray.contains_distance(min_distance) || ray.contains_distance(max_distance);
    }
};
inline Bounds bunion(const Bounds& a, const Bounds& b) {  // impl Add for Bounds, bounds.rs:91-108
    return {{rmin(a.min.x, b.min.x), rmin(a.min.y, b.min.y), rmin(a.min.z, b.min.z)},
            {rmax(a.max.x, b.max.x), rmax(a.max.y, b.max.y), rmax(a.max.z, b.max.z)}};
}

// ---- Matrix / Transformation  src/transformation.rs ---------------------------------
struct Matrix {
    double m[4][4];
};
inline Matrix identity() { return {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}}; }
inline Matrix transpose(const Matrix& a) {  // transformation.rs:58-68
    Matrix r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) r.m[i][j] = a.m[j][i];
    return r;
}
inline Matrix matmul(const Matrix& a, const Matrix& b) {  // transformation.rs:202-218 (accumulates from 0.0)
    Matrix r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += a.m[i][k] * b.m[k][j];
            r.m[i][j] = acc;
        }
    return r;
}
// Matrix::inverse (Cramer's rule), transformation.rs:71-195.  Term order kept.
inline bool inverse(const Matrix& a, Matrix& out) {
    const double(*m)[4] = a.m;
    double inv[4][4];
    inv[0][0] = m[1][1] * m[2][2] * m[3][3] - m[1][1] * m[2][3] * m[3][2] - m[2][1] * m[1][2] * m[3][3] + m[2][1] * m[1][3] * m[3][2] + m[3][1] * m[1][2] * m[2][3] - m[3][1] * m[1][3] * m[2][2];
    inv[0][1] = -m[0][1] * m[2][2] * m[3][3] + m[0][1] * m[2][3] * m[3][2] + m[2][1] * m[0][2] * m[3][3] - m[2][1] * m[0][3] * m[3][2] - m[3][1] * m[0][2] * m[2][3] + m[3][1] * m[0][3] * m[2][2];
    inv[0][2] = m[0][1] * m[1][2] * m[3][3] - m[0][1] * m[1][3] * m[3][2] - m[1][1] * m[0][2] * m[3][3] + m[1][1] * m[0][3] * m[3][2] + m[3][1] * m[0][2] * m[1][3] - m[3][1] * m[0][3] * m[1][2];
    inv[0][3] = -m[0][1] * m[1][2] * m[2][3] + m[0][1] * m[1][3] * m[2][2] + m[1][1] * m[0][2] * m[2][3] - m[1][1] * m[0][3] * m[2][2] - m[2][1] * m[0][2] * m[1][3] + m[2][1] * m[0][3] * m[1][2];
    inv[1][0] = -m[1][0] * m[2][2] * m[3][3] + m[1][0] * m[2][3] * m[3][2] + m[2][0] * m[1][2] * m[3][3] - m[2][0] * m[1][3] * m[3][2] - m[3][0] * m[1][2] * m[2][3] + m[3][0] * m[1][3] * m[2][2];
    inv[1][1] = m[0][0] * m[2][2] * m[3][3] - m[0][0] * m[2][3] * m[3][2] - m[2][0] * m[0][2] * m[3][3] + m[2][0] * m[0][3] * m[3][2] + m[3][0] * m[0][2] * m[2][3] - m[3][0] * m[0][3] * m[2][2];
    inv[1][2] = -m[0][0] * m[1][2] * m[3][3] + m[0][0] * m[1][3] * m[3][2] + m[1][0] * m[0][2] * m[3][3] - m[1][0] * m[0][3] * m[3][2] - m[3][0] * m[0][2] * m[1][3] + m[3][0] * m[0][3] * m[1][2];
    inv[1][3] = m[0][0] * m[1][2] * m[2][3] - m[0][0] * m[1][3] * m[2][2] - m[1][0] * m[0][2] * m[2][3] + m[1][0] * m[0][3] * m[2][2] + m[2][0] * m[0][2] * m[1][3] - m[2][0] * m[0][3] * m[1][2];
    inv[2][0] = m[1][0] * m[2][1] * m[3][3] - m[1][0] * m[2][3] * m[3][1] - m[2][0] * m[1][1] * m[3][3] + m[2][0] * m[1][3] * m[3][1] + m[3][0] * m[1][1] * m[2][3] - m[3][0] * m[1][3] * m[2][1];
    inv[2][1] = -m[0][0] * m[2][1] * m[3][3] + m[0][0] * m[2][3] * m[3][1] + m[2][0] * m[0][1] * m[3][3] - m[2][0] * m[0][3] * m[3][1] - m[3][0] * m[0][1] * m[2][3] + m[3][0] * m[0][3] * m[2][1];
    inv[2][2] = m[0][0] * m[1][1] * m[3][3] - m[0][0] * m[1][3] * m[3][1] - m[1][0] * m[0][1] * m[3][3] + m[1][0] * m[0][3] * m[3][1] + m[3][0] * m[0][1] * m[1][3] - m[3][0] * m[0][3] * m[1][1];
    inv[2][3] = -m[0][0] * m[1][1] * m[2][3] + m[0][0] * m[1][3] * m[2][1] + m[1][0] * m[0][1] * m[2][3] - m[1][0] * m[0][3] * m[2][1] - m[2][0] * m[0][1] * m[1][3] + m[2][0] * m[0][3] * m[1][1];
    inv[3][0] = -m[1][0] * m[2][1] * m[3][2] + m[1][0] * m[2][2] * m[3][1] + m[2][0] * m[1][1] * m[3][2] - m[2][0] * m[1][2] * m[3][1] - m[3][0] * m[1][1] * m[2][2] + m[3][0] * m[1][2] * m[2][1];
    inv[3][1] = m[0][0] * m[2][1] * m[3][2] - m[0][0] * m[2][2] * m[3][1] - m[2][0] * m[0][1] * m[3][2] + m[2][0] * m[0][2] * m[3][1] + m[3][0] * m[0][1] * m[2][2] - m[3][0] * m[0][2] * m[2][1];
    inv[3][2] = -m[0][0] * m[1][1] * m[3][2] + m[0][0] * m[1][2] * m[3][1] + m[1][0] * m[0][1] * m[3][2] - m[1][0] * m[0][2] * m[3][1] - m[3][0] * m[0][1] * m[1][2] + m[3][0] * m[0][2] * m[1][1];
    inv[3][3] = m[0][0] * m[1][1] * m[2][2] - m[0][0] * m[1][2] * m[2][1] - m[1][0] * m[0][1] * m[2][2] + m[1][0] * m[0][2] * m[2][1] + m[2][0] * m[0][1] * m[1][2] - m[2][0] * m[0][2] * m[1][1];
    double det = m[0][0] * inv[0][0] + m[0][1] * inv[1][0] + m[0][2] * inv[2][0] + m[0][3] * inv[3][0];
    if (det != 0.0) {
        double inv_det = 1.0 / det;
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) out.m[i][j] = inv[i][j] * inv_det;
        return true;
    }
    return false;
}

struct Transformation {
    Matrix matrix, inv;
    Transformation inverse() const { return {inv, matrix}; }  // transformation.rs:258-263
};
inline Transformation tmul(const Transformation& a, const Transformation& b) {  // transformation.rs:392-414
    return {matmul(a.matrix, b.matrix), matmul(b.inv, a.inv)};
}
inline Transformation translate(double dx, double dy, double dz) {  // transformation.rs:265-284
    return {{{{1, 0, 0, dx}, {0, 1, 0, dy}, {0, 0, 1, dz}, {0, 0, 0, 1}}},
            {{{1, 0, 0, -dx}, {0, 1, 0, -dy}, {0, 0, 1, -dz}, {0, 0, 0, 1}}}};
}
inline Transformation scale(double x, double y, double z) {  // transformation.rs:286-305
    return {{{{x, 0, 0, 0}, {0, y, 0, 0}, {0, 0, z, 0}, {0, 0, 0, 1}}},
            {{{1.0 / x, 0, 0, 0}, {0, 1.0 / y, 0, 0}, {0, 0, 1.0 / z, 0}, {0, 0, 0, 1}}}};
}
inline Transformation rotate_x(double radians) {  // transformation.rs:307-320
    double s = std::sin(radians), c = std::cos(radians);
    Matrix m{{{1, 0, 0, 0}, {0, c, -s, 0}, {0, s, c, 0}, {0, 0, 0, 1}}};
    return {m, transpose(m)};
}
inline Transformation rotate_y(double radians) {  // transformation.rs:322-335
    double s = std::sin(radians), c = std::cos(radians);
    Matrix m{{{c, 0, s, 0}, {0, 1, 0, 0}, {-s, 0, c, 0}, {0, 0, 0, 1}}};
    return {m, transpose(m)};
}
inline Transformation rotate_z(double radians) {  // transformation.rs:337-350
    double s = std::sin(radians), c = std::cos(radians);
    Matrix m{{{c, -s, 0, 0}, {s, c, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}};
    return {m, transpose(m)};
}
inline Transformation look_at(V3 origin, V3 target, V3 up) {  // transformation.rs:352-366
    V3 z = normalized(target - origin);
    V3 x = normalized(cross(normalized(up), z));
    V3 y = normalized(cross(z, x));
    Matrix m{{{x.x, y.x, z.x, origin.x}, {x.y, y.y, z.y, origin.y}, {x.z, y.z, z.z, origin.z}, {0, 0, 0, 1}}};
    Matrix i;
    inverse(m, i);
    return {m, i};
}
inline Transformation perspective(double fov, double near, double far) {  // transformation.rs:368-381
    Matrix m{{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, far / (far - near), -far * near / (far - near)}, {0, 0, 1, 0}}};
    Matrix i;
    inverse(m, i);
    Transformation persp{m, i};
    double inv_tan_ang = 1.0 / std::tan(to_radians(fov) * 0.5);
    return tmul(persp, scale(inv_tan_ang, inv_tan_ang, 1.0));
}
inline Transformation orthographic(double near, double far) {  // transformation.rs:383-385
    return tmul(scale(1.0, 1.0, 1.0 / (far - near)), translate(0.0, 0.0, -near));
}
inline V3 xf_point(const Transformation& t, V3 p) {  // transformation.rs:421-432
    const double(*m)[4] = t.matrix.m;
    V3 r{m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3],
         m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
         m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3]};
    double w = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
    return r / w;
}
inline V3 xf_vector(const Transformation& t, V3 v) {  // transformation.rs:434-444
    const double(*m)[4] = t.matrix.m;
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
}
inline V3 xf_normal(const Transformation& t, V3 n) {  // transformation.rs:446-457 (inverse transpose)
    const double(*inv)[4] = t.inv.m;
    return {inv[0][0] * n.x + inv[1][0] * n.y + inv[2][0] * n.z, inv[0][1] * n.x + inv[1][1] * n.y + inv[2][1] * n.z,
            inv[0][2] * n.x + inv[1][2] * n.y + inv[2][2] * n.z};
}
inline Ray xf_ray(const Transformation& t, const Ray& ray) {  // transformation.rs:459-466
    Ray r = Ray::make(xf_point(t, ray.origin), xf_vector(t, ray.direction));
    r.update_max_distance(ray.max_distance);
    return r;
}
inline Bounds xf_bounds(const Transformation& t, const Bounds& b) {  // transformation.rs:468-486
    V3 c[8] = {{b.min.x, b.min.y, b.min.z}, {b.min.x, b.min.y, b.max.z}, {b.min.x, b.max.y, b.min.z}, {b.min.x, b.max.y, b.max.z},
               {b.max.x, b.min.y, b.min.z}, {b.max.x, b.min.y, b.max.z}, {b.max.x, b.max.y, b.min.z}, {b.max.x, b.max.y, b.max.z}};
    V3 p0 = xf_point(t, c[0]);
    Bounds acc = Bounds::make(p0, p0);
    for (int i = 1; i < 8; ++i) {
        V3 p = xf_point(t, c[i]);
        acc = bunion(acc, Bounds::make(p, p));
    }
    return acc;
}

}  // namespace orc
