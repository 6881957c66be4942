// Host-side 4x4 transformation algebra for scene set-up (camera matrices, disk frames, shape bounds).
// Mirrors the *arithmetic* of src/transformation.rs (same operand order, so the matrices uploaded to
// the GPU are bit-identical to the reference's) with a rule-driven adjugate instead of the reference's
// unrolled 96-term listing.
#pragma once
#include <cstring>
#include "cray_math.cuh"

namespace cray {

struct Mat4 {
    double m[4][4];
    static Mat4 identity() {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0 : 0.0;
        return r;
    }
    Mat4 transposed() const {
        Mat4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = m[j][i];
        return r;
    }
};

inline Mat4 operator*(const Mat4& a, const Mat4& b) {  // transformation.rs:202-218: accumulate k = 0..3 onto 0.0
    Mat4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += a.m[i][k] * b.m[k][j];
            r.m[i][j] = acc;
        }
    return r;
}

// Matrix::inverse (transformation.rs:71-195).  Entry (i,j) of the adjugate is the signed 3x3 minor
// obtained by deleting row j and column i, expanded along its first remaining column; each triple
// product is evaluated left to right and the six terms are summed in the listing's order
// (+ - - + + -), which reproduces the reference's rounding exactly.
inline bool invert(const Mat4& a, Mat4& out) {
    double adj[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            int r[3], c[3];
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != j) r[n++] = k;
            for (int k = 0, n = 0; k < 4; ++k)
                if (k != i) c[n++] = k;
            const double s = ((i + j) & 1) ? -1.0 : 1.0;
            const double p0 = s * a.m[r[0]][c[0]], p1 = s * a.m[r[1]][c[0]], p2 = s * a.m[r[2]][c[0]];
            double acc = p0 * a.m[r[1]][c[1]] * a.m[r[2]][c[2]];
            acc = acc - p0 * a.m[r[1]][c[2]] * a.m[r[2]][c[1]];
            acc = acc - p1 * a.m[r[0]][c[1]] * a.m[r[2]][c[2]];
            acc = acc + p1 * a.m[r[0]][c[2]] * a.m[r[2]][c[1]];
            acc = acc + p2 * a.m[r[0]][c[1]] * a.m[r[1]][c[2]];
            acc = acc - p2 * a.m[r[0]][c[2]] * a.m[r[1]][c[1]];
            adj[i][j] = acc;
        }
    const double det = a.m[0][0] * adj[0][0] + a.m[0][1] * adj[1][0] + a.m[0][2] * adj[2][0] + a.m[0][3] * adj[3][0];
    if (det == 0.0) return false;
    const double inv_det = 1.0 / det;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out.m[i][j] = adj[i][j] * inv_det;
    return true;
}

struct Xform {  // Transformation { matrix, inverse }  transformation.rs:247-251
    Mat4 fwd, inv;
    Xform inverted() const { return {inv, fwd}; }
};
inline Xform operator*(const Xform& a, const Xform& b) { return {a.fwd * b.fwd, b.inv * a.inv}; }  // :392-414

inline Xform xf_translate(double dx, double dy, double dz) {
    Xform t{Mat4::identity(), Mat4::identity()};
    t.fwd.m[0][3] = dx; t.fwd.m[1][3] = dy; t.fwd.m[2][3] = dz;
    t.inv.m[0][3] = -dx; t.inv.m[1][3] = -dy; t.inv.m[2][3] = -dz;
    return t;
}
inline Xform xf_scale(double x, double y, double z) {
    Xform t{Mat4::identity(), Mat4::identity()};
    t.fwd.m[0][0] = x; t.fwd.m[1][1] = y; t.fwd.m[2][2] = z;
    t.inv.m[0][0] = 1.0 / x; t.inv.m[1][1] = 1.0 / y; t.inv.m[2][2] = 1.0 / z;
    return t;
}
// axis 0: rotate_x, 1: rotate_y, 2: rotate_z (transformation.rs:307-350); inverse is the transpose.
inline Xform xf_rotate(int axis, double radians) {
    const double s = std::sin(radians), c = std::cos(radians);
    Mat4 m = Mat4::identity();
    const int u = (axis + 1) % 3, v = (axis + 2) % 3;  // the rotated plane (u,v): x->(y,z) y->(z,x) z->(x,y)
    m.m[u][u] = c; m.m[u][v] = -s; m.m[v][u] = s; m.m[v][v] = c;
    return {m, m.transposed()};
}
inline Xform xf_look_at(V3 origin, V3 target, V3 up) {  // :352-366
    const V3 z = normalized(target - origin);
    const V3 x = normalized(cross(normalized(up), z));
    const V3 y = normalized(cross(z, x));
    Mat4 m = Mat4::identity();
    m.m[0][0] = x.x; m.m[0][1] = y.x; m.m[0][2] = z.x; m.m[0][3] = origin.x;
    m.m[1][0] = x.y; m.m[1][1] = y.y; m.m[1][2] = z.y; m.m[1][3] = origin.y;
    m.m[2][0] = x.z; m.m[2][1] = y.z; m.m[2][2] = z.z; m.m[2][3] = origin.z;
    Mat4 inv = Mat4::identity();
    invert(m, inv);
    return {m, inv};
}
inline Xform xf_perspective(double fov_deg, double near, double far) {  // :368-381
    Mat4 m = Mat4::identity();
    m.m[2][2] = far / (far - near);
    m.m[2][3] = -far * near / (far - near);
    m.m[3][2] = 1.0;
    m.m[3][3] = 0.0;
    Mat4 inv = Mat4::identity();
    invert(m, inv);
    const double inv_tan_ang = 1.0 / std::tan(to_radians(fov_deg) * 0.5);
    return Xform{m, inv} * xf_scale(inv_tan_ang, inv_tan_ang, 1.0);
}
inline Xform xf_orthographic(double near, double far) {  // :383-385
    return xf_scale(1.0, 1.0, 1.0 / (far - near)) * xf_translate(0.0, 0.0, -near);
}
// get_camera_from_raster_transformation camera.rs:25-53 (film_height = film.width, sic)
inline Xform camera_from_raster(const Xform& screen_from_camera, uint32_t film_width_px) {
    const double film_width = (double)film_width_px, film_height = (double)film_width_px;
    double screen_width, screen_height;
    if (film_width > film_height) { screen_width = film_width / film_height; screen_height = 1.0; }
    else { screen_width = 1.0; screen_height = film_height / film_width; }
    const Xform screen_from_raster = xf_scale(2.0 * screen_width / film_width, -2.0 * screen_height / film_height, 1.0) *
                                     xf_translate(-film_width / 2.0, -film_height / 2.0, 0.0);
    return screen_from_camera.inverted() * screen_from_raster;
}

inline V3 apply_point(const Mat4& m, V3 p) {  // full homogeneous divide, transformation.rs:421-432
    V3 r = mk(m.m[0][0] * p.x + m.m[0][1] * p.y + m.m[0][2] * p.z + m.m[0][3],
              m.m[1][0] * p.x + m.m[1][1] * p.y + m.m[1][2] * p.z + m.m[1][3],
              m.m[2][0] * p.x + m.m[2][1] * p.y + m.m[2][2] * p.z + m.m[2][3]);
    return r / (m.m[3][0] * p.x + m.m[3][1] * p.y + m.m[3][2] * p.z + m.m[3][3]);
}
inline Affine affine_of(const Mat4& m) {
    Affine a;
    std::memcpy(a.m, m.m, sizeof(a.m));
    return a;
}
inline Box3 box_of_point(V3 p) { return {p, p}; }
inline Box3 box_union(const Box3& a, const Box3& b) {  // bounds.rs:91-108
    return {mk(rmin(a.lo.x, b.lo.x), rmin(a.lo.y, b.lo.y), rmin(a.lo.z, b.lo.z)), mk(rmax(a.hi.x, b.hi.x), rmax(a.hi.y, b.hi.y), rmax(a.hi.z, b.hi.z))};
}
// Transformable<Bounds> transformation.rs:468-486: the 8 corners in (min/max) lexicographic order
inline Box3 transform_box(const Mat4& m, const Box3& b) {
    Box3 acc{};
    for (int k = 0; k < 8; ++k) {
        V3 corner = mk((k & 4) ? b.hi.x : b.lo.x, (k & 2) ? b.hi.y : b.lo.y, (k & 1) ? b.hi.z : b.lo.z);
        V3 p = apply_point(m, corner);
        acc = k == 0 ? box_of_point(p) : box_union(acc, box_of_point(p));
    }
    return acc;
}
inline V3 box_centroid(const Box3& b) { return mk((b.lo.x + b.hi.x) * 0.5, (b.lo.y + b.hi.y) * 0.5, (b.lo.z + b.hi.z) * 0.5); }  // bounds.rs:22
inline double box_surface_area(const Box3& b) {  // bounds.rs:29-32
    V3 d = b.hi - b.lo;
    return 2.0 * (d.x * d.y + d.y * d.z + d.z * d.x);
}
inline int box_maximum_extent(const Box3& b) {  // bounds.rs:36-45
    V3 d = b.hi - b.lo;
    if (d.x > d.y && d.x > d.z) return 0;
    if (d.y > d.z) return 1;
    return 2;
}

}  // namespace cray
