// Ray / primitive intersection in f64, restating src/shape.rs:157-400 operation by operation
// (compiled with --fmad=false, so every product and sum rounds exactly as in the reference).
#pragma once
#include "device_types.cuh"

namespace cray {

struct Hit {
    double t;
    double u, v;      // triangle barycentrics (Moeller-Trumbore); spheres/disks fill surface uv later
    uint32_t slot;    // index into the leaf-ordered LeafPrim array that was traversed, CRAY_NO_HIT = miss
};

#ifndef CRAY_PRIM_NOALLOC
#define CRAY_PRIM_NOALLOC 1   // 1: intersection records bypass L1 allocation (each is read about once; L1 is kept for the nodes)
#endif
__device__ __forceinline__ double2 ldg_record(const double2* p) {
#if CRAY_PRIM_NOALLOC
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

// Five 16-byte read-only loads straight into registers (no round trip through a local-memory copy of the record).
__device__ __forceinline__ LeafPrim load_leaf_prim(const LeafPrim* p) {
    const double2* src = reinterpret_cast<const double2*>(p);
    const double2 a = ldg_record(src), b = ldg_record(src + 1), c = ldg_record(src + 2), e = ldg_record(src + 3), f = ldg_record(src + 4);
    LeafPrim r;
    r.d[0] = a.x; r.d[1] = a.y; r.d[2] = b.x; r.d[3] = b.y; r.d[4] = c.x; r.d[5] = c.y; r.d[6] = e.x; r.d[7] = e.y; r.d[8] = f.x;
    r.prim = (uint32_t)__double2loint(f.y);
    r.kind = (uint32_t)__double2hiint(f.y);
    return r;
}

// Shape::Triangle arm of intersect / intersects (shape.rs:214-262, :345-367) up to, but not including, the
// final range check: true iff the determinant / u / v tests pass; t is the hit distance.
//
// The reference divides three times (u, v, t).  Most candidates fail the u or v test, so the quotients are only
// formed when the cheap tests on the numerators cannot decide; the guards below are chosen so that they can never
// disagree with the reference's comparisons on the rounded quotients:
//   * opposite signs of numerator and denominator  =>  quotient < 0 (the 1e-290 floor keeps clear of underflow to -0)
//   * |numerator| > |denominator| * (1 + 2^-40)     =>  the rounded quotient is > 1
// When the tests pass, u, v and t are exactly the reference's values.
__device__ __forceinline__ bool triangle_eval(const double* d9, V3 o, V3 dir, double& t, double& u, double& v) {
    const V3 v0 = mk(d9[0], d9[1], d9[2]), e1 = mk(d9[3], d9[4], d9[5]), e2 = mk(d9[6], d9[7], d9[8]);
    const V3 P = cross(dir, e2);
    const double denominator = dot(P, e1);
    if (denominator > -kEpsilon && denominator < kEpsilon) return false;
    const V3 T = o - v0;
    const double un = dot(P, T);
    const double ad = fabs(denominator), slack = ad * 1.0000000000009095;  // 1 + 2^-40
    if (fabs(un) > 1e-290 && (un < 0.0) != (denominator < 0.0)) return false;  // u < 0
    if (fabs(un) > slack) return false;                                       // u > 1
    const V3 Q = cross(T, e1);
    const double vn = dot(Q, dir);
    if (fabs(vn) > 1e-290 && (vn < 0.0) != (denominator < 0.0)) return false;  // v < 0
    if (fabs(vn) > slack) return false;                                       // v > 1 and u >= 0  =>  u + v > 1
    u = un / denominator;
    if (u < 0.0 || u > 1.0) return false;
    v = vn / denominator;
    if (v < 0.0 || u + v > 1.0) return false;
    t = dot(Q, e2) / denominator;  // T.cross(e1).dot(e2): the same value as Q, recomputed in the reference
    return true;
}
// ... and with it: true iff the reference accepts the hit, i.e. ray.contains_distance(t).
__device__ __forceinline__ bool triangle_hit(const double* d9, V3 o, V3 dir, double ray_max, double& t, double& u, double& v) {
    return triangle_eval(d9, o, dir, t, u, v) && contains_distance(t, ray_max);
}

// Shape::Sphere arm (shape.rs:159-212, :316-343).  world_to_object = translate(-origin): each object-space
// coordinate is 1*p.x + 0*p.y + 0*p.z + (-origin.x), which for finite p is exactly p.x - origin.x.
// `t_report` is what Primitive::intersect stores as `distance` (= ray.max_distance after the call).
__device__ __forceinline__ bool sphere_hit(const double* d9, V3 o, V3 dir, double ray_max, double& t, double& t_report) {
    const V3 oc = mk(o.x + (-d9[0]), o.y + (-d9[1]), o.z + (-d9[2]));
    const double radius = d9[3];
    const double obj_max = transformed_max_distance(ray_max);
    const double a = magnitude_squared(dir);
    const double b = 2.0 * dot(oc, dir);
    const double c = magnitude_squared(oc) - radius * radius;
    const double discriminant = b * b - 4.0 * a * c;
    if (discriminant < 0.0) return false;
    const double discriminant_sqrt = sqrt(discriminant);
    const double inv_2_a = 1.0 / (2.0 * a);
    double distance = (-b - discriminant_sqrt) * inv_2_a;
    if (!contains_distance(distance, obj_max)) {
        distance = (-b + discriminant_sqrt) * inv_2_a;
        if (!contains_distance(distance, obj_max)) return false;
    }
    t = distance;
    t_report = contains_distance(distance, ray_max) ? distance : ray_max;  // ray.update_max_distance(distance)
    return true;
}

// Shape::Disk arm (shape.rs:263-309, :369-398).
__device__ __forceinline__ bool disk_hit(const DiskXf& k, V3 o, V3 dir, double ray_max, double& t, double& lx, double& ly, double& d2) {
    const V3 oo = xf_point(k.w2o, o);
    const V3 od = xf_vector(k.w2o, dir);
    const double obj_max = transformed_max_distance(ray_max);
    if (od.z == 0.0) return false;
    t = -oo.z / od.z;
    if (!contains_distance(t, obj_max)) return false;
    lx = oo.x + od.x * t;
    ly = oo.y + od.y * t;
    d2 = lx * lx + ly * ly;
    if (d2 < k.inner_radius * k.inner_radius || d2 > k.radius * k.radius) return false;
    return contains_distance(t, ray_max);
}

// One primitive of a leaf, closest-hit flavour (Primitive::intersect primitive.rs:50-73): on acceptance the
// ray's max_distance shrinks to the hit distance.  Returns true if accepted.
__device__ __forceinline__ bool leaf_prim_closest(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double& ray_max, double& u, double& v) {
    const uint32_t kind = lp.kind & 0xFFu;
    double t;
    if (kind == PRIM_TRIANGLE) {
        double tu, tv;
        if (!triangle_hit(lp.d, o, dir, ray_max, t, tu, tv)) return false;
        ray_max = t; u = tu; v = tv;
        return true;
    }
    if (kind == PRIM_SPHERE) {
        double rep;
        if (!sphere_hit(lp.d, o, dir, ray_max, t, rep)) return false;
        ray_max = rep; u = 0.0; v = 0.0;
        return true;
    }
    double lx, ly, d2;
    if (!disk_hit(s.disks[disk_index(lp.kind)], o, dir, ray_max, t, lx, ly, d2)) return false;
    ray_max = t; u = 0.0; v = 0.0;
    return true;
}

// Primitive::intersects primitive.rs:75-82
__device__ __forceinline__ bool leaf_prim_any(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double ray_max) {
    const uint32_t kind = lp.kind & 0xFFu;
    double t, a, b, c;
    if (kind == PRIM_TRIANGLE) return triangle_hit(lp.d, o, dir, ray_max, t, a, b);
    if (kind == PRIM_SPHERE) {
        // Shape::intersects tests the object-space ray's range only (shape.rs:336-341)
        return sphere_hit(lp.d, o, dir, ray_max, t, a);
    }
    return disk_hit(s.disks[disk_index(lp.kind)], o, dir, ray_max, t, a, b, c);
}

// Spheres and disks of the wide-BVH closest-hit test below.  Out of line: scenes have a handful of them, and inlined their f64
// transforms would set the register allocation of the whole traversal kernel.
__device__ __noinline__ int analytic_candidate(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double& ray_max, bool have_hit) {
    double cand = ray_max, u, v;
    if (leaf_prim_closest(s, lp, o, dir, cand, u, v)) { ray_max = cand; return 1; }
    if (have_hit) {  // analytic shapes are few: re-evaluate against the next f64 above ray_max to detect a tie
        double tie = __longlong_as_double(__double_as_longlong(ray_max) + 1);
        if (leaf_prim_closest(s, lp, o, dir, tie, u, v) && tie == ray_max) return 2;
    }
    return 0;
}

#ifndef CRAY_PRIM_NOCOPY
#define CRAY_PRIM_NOCOPY 0   // 1: the out-of-line analytic tests read the record themselves (no local-memory copy of every tested record)
#endif
__device__ __noinline__ int analytic_candidate_at(const SceneView& s, const LeafPrim* record, V3 o, V3 dir, double& ray_max, bool have_hit) {
    const LeafPrim lp = load_leaf_prim(record);
    return analytic_candidate(s, lp, o, dir, ray_max, have_hit);
}
__device__ __noinline__ bool analytic_any_at(const SceneView& s, const LeafPrim* record, V3 o, V3 dir, double ray_max) {
    const LeafPrim lp = load_leaf_prim(record);
    return leaf_prim_any(s, lp, o, dir, ray_max);
}

// ... and the any-hit test of a sphere or disk, out of line for the same reason (F32 mode)
__device__ __noinline__ bool analytic_any(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double ray_max) {
    return leaf_prim_any(s, lp, o, dir, ray_max);
}

// F32 mode: the f64 triangle test of the parity mode for the rare f32 hit that needs confirming (prim_round_any32)
__device__ __noinline__ bool triangle_any_f64(const SceneView& s, uint32_t slot, V3 o, V3 dir, double ray_max) {
    const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
    double t, u, v;
    return triangle_hit(lp.d, o, dir, ray_max, t, u, v);
}

// Moeller-Trumbore without its range tests: t, u, v of the ray's crossing of the triangle's PLANE.  The F32 mode evaluates
// the hit it found with this when the f64 test proper rejects it (a ray grazing an edge that f32 placed just inside).
__device__ __forceinline__ bool triangle_eval_unchecked(const double* d9, V3 o, V3 dir, double& t, double& u, double& v) {
    const V3 v0 = mk(d9[0], d9[1], d9[2]), e1 = mk(d9[3], d9[4], d9[5]), e2 = mk(d9[6], d9[7], d9[8]);
    const V3 P = cross(dir, e2);
    const double denominator = dot(P, e1);
    if (denominator == 0.0) return false;
    const V3 T = o - v0;
    const V3 Q = cross(T, e1);
    u = dot(P, T) / denominator;
    v = dot(Q, dir) / denominator;
    t = dot(Q, e2) / denominator;
    return true;
}

// Wide-BVH closest-hit flavour.  0: rejected; 1: accepted (ray_max shrinks); 2: exact tie -- the strict `<` of
// ray.rs:26 rejects it against the current ray_max, but its distance is bit-equal to it, so the reference keeps
// whichever of the two primitives its own traversal reaches first.
__device__ __forceinline__ int leaf_prim_candidate(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double& ray_max, bool have_hit, double& u, double& v,
                                                   const LeafPrim* record = nullptr) {
    const uint32_t kind = lp.kind & 0xFFu;
    if (kind == PRIM_TRIANGLE) {
        double t, tu, tv;
        if (!triangle_eval(lp.d, o, dir, t, tu, tv) || !(t > kEpsilon)) return 0;
        if (t < ray_max) { ray_max = t; u = tu; v = tv; return 1; }
        if (have_hit && t == ray_max) { u = tu; v = tv; return 2; }
        return 0;
    }
#if CRAY_PRIM_NOCOPY
    return analytic_candidate_at(s, record, o, dir, ray_max, have_hit);
#else
    (void)record;
    return analytic_candidate(s, lp, o, dir, ray_max, have_hit);
#endif
}

// The rest of PrimitiveIntersection (location, normal, uv) for an accepted hit at distance t.  `want_uv` = false skips the
// texture coordinates (atan2 / acos on spheres and disks): nothing reads them when every texture of the material hit is a
// constant, and the light's self-intersection test (shape.rs:487-502) only needs the location.
__device__ __forceinline__ void surface_at(const SceneView& s, const LeafPrim& lp, V3 o, V3 dir, double t, double bu, double bv,
                                           V3& location, V3& normal, double& tex_u, double& tex_v, bool want_uv = true) {
    const uint32_t kind = lp.kind & 0xFFu;
    tex_u = 0.0; tex_v = 0.0;
    if (kind == PRIM_TRIANGLE) {  // shape.rs:249-258
        const TriShade& ts = s.tri_shade[lp.prim];
        location = o + dir * t;
        const V3 n0 = mk(ts.n0[0], ts.n0[1], ts.n0[2]);
        V3 n01 = mk(0.0, 0.0, 0.0), n02 = mk(0.0, 0.0, 0.0);
        if (!(lp.kind & kKindFlatTriangle)) {  // flat triangles: both are +0.0, and the record's first sector is all that is read
            n01 = mk(ts.n01[0], ts.n01[1], ts.n01[2]);
            n02 = mk(ts.n02[0], ts.n02[1], ts.n02[2]);
        }
        normal = normalized(n0 + n01 * bu + n02 * bv);
        if (want_uv) {
            tex_u = ts.uv0[0] + ts.uv01[0] * bu + ts.uv02[0] * bv;
            tex_v = ts.uv0[1] + ts.uv01[1] * bu + ts.uv02[1] * bv;
        }
        return;
    }
    if (kind == PRIM_SPHERE) {  // shape.rs:178-195
        const V3 oc = mk(o.x + (-lp.d[0]), o.y + (-lp.d[1]), o.z + (-lp.d[2]));
        const double radius = lp.d[3];
        const V3 loc = oc + dir * t;
        if (want_uv) {
            double phi = atan2(loc.y, loc.x);
            if (phi < 0.0) phi += kPi * 2.0;
            tex_u = phi / (kPi * 2.0);
            tex_v = acos(loc.z / radius) * kFrac1Pi;
        }
        location = mk(loc.x + lp.d[0], loc.y + lp.d[1], loc.z + lp.d[2]);  // object_to_world = translate(origin)
        normal = loc / radius;                                               // inverse-transpose of a translation is the identity
        return;
    }
    const DiskXf& k = s.disks[disk_index(lp.kind)];  // shape.rs:283-308
    const V3 oo = xf_point(k.w2o, o);
    const V3 od = xf_vector(k.w2o, dir);
    const double lx = oo.x + od.x * t, ly = oo.y + od.y * t;
    if (want_uv) {
        const double d2 = lx * lx + ly * ly;
        double theta = atan2(ly, lx);
        if (theta < 0.0) theta += kPi * 2.0;
        tex_u = theta / (kPi * 2.0);
        tex_v = sqrt(d2) / k.radius;
    }
    location = xf_point(k.o2w, mk(lx, ly, 0.0));
    normal = xf_normal_with_inverse(k.w2o, mk(0.0, 0.0, 1.0));
}

}  // namespace cray
