// BVH traversal device functions.
//
//  traverse_exact<ANY>: the reference's Bvh::intersect / Bvh::intersects (src/bvh.rs:58-147) on the flattened
//    binary SAH tree: f64 slab test with the reference's accept rule (bounds.rs:62-88) or-ed with
//    Bounds::contains, near/far push order by sign of dir[split_axis], strict `t < max_distance` acceptance.
//    Results are bit-identical to the reference by construction (including its false box misses).
//
//  traverse_wide<ANY>: production traversal of the 8-wide quantised BVH.  Boxes are tested in f32 with every
//    rounding error pushed outward (see "conservative slab test" below) so no primitive the exact f64 leaf test
//    would accept is ever culled; leaves run the same f64 primitive tests as the exact mode.  Among primitives
//    with bit-equal t the one the reference would have visited first wins (reference_visits_first).
#pragma once
#include "shapes.cuh"

namespace cray {

constexpr int kExactStack = 96;
constexpr int kWideStack = 32;

template <bool ANY>
__device__ __forceinline__ bool traverse_exact(const SceneView& s, V3 o, V3 dir, double ray_max, Hit& hit) {
    uint32_t stack[kExactStack];
    int sp = 0;
    stack[sp++] = 0;
    hit.slot = CRAY_NO_HIT;
    hit.t = ray_max;
    hit.u = hit.v = 0.0;
    while (sp > 0) {
        const BinNode* np = s.bin_nodes + stack[--sp];
        const int4* raw = reinterpret_cast<const int4*>(np);
        BinNode n;
        int4* dst = reinterpret_cast<int4*>(&n);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = __ldg(raw + i);
        if (!bounds_intersects(n.box, o, dir, ray_max) && !bounds_contains(n.box, o)) continue;
        if (n.axis == 3) {
            for (uint32_t i = 0; i < n.b; ++i) {
                const LeafPrim lp = load_leaf_prim(s.bin_prims + n.a + i);
                if (ANY) {
                    if (leaf_prim_any(s, lp, o, dir, ray_max)) return true;
                } else {
                    double u, v;
                    if (leaf_prim_closest(s, lp, o, dir, ray_max, u, v)) {
                        // bvh.rs:78-82: replace `current` only if the reported distance is strictly smaller
                        if (hit.slot == CRAY_NO_HIT || ray_max < hit.t) { hit.t = ray_max; hit.u = u; hit.v = v; hit.slot = n.a + i; }
                    }
                }
            }
        } else {
            const bool neg_dir = dir[(int)n.axis] < 0.0;
            if (sp + 2 > kExactStack) continue;  // deeper than any tree the builder can produce; keep memory safe
            if (neg_dir) { stack[sp++] = n.a; stack[sp++] = n.b; }
            else { stack[sp++] = n.b; stack[sp++] = n.a; }
        }
    }
    return hit.slot != CRAY_NO_HIT;
}

// Would the reference's depth-first traversal reach primitive `pa` before `pb` for this ray direction?
// (Both given as ranks in the reference leaf order.)  Walks down from the root to their lowest common ancestor.
__device__ __noinline__ bool reference_visits_first(const SceneView& s, uint32_t rank_a, uint32_t rank_b, V3 dir) {
    const bool swapped = rank_a > rank_b;
    const uint32_t lo = swapped ? rank_b : rank_a, hi = swapped ? rank_a : rank_b;
    uint32_t node = 0;
    bool lo_first = true;
    for (int guard = 0; guard < 256; ++guard) {
        const BinNode& n = s.bin_nodes[node];
        if (n.axis == 3) { lo_first = true; break; }                          // same leaf: list order
        if (hi < n.right_first) node = n.a;                                   // both in the left subtree
        else if (lo >= n.right_first) node = n.b;                             // both in the right subtree
        else { lo_first = !(dir[(int)n.axis] < 0.0); break; }                 // bvh.rs:92-98: right first iff dir[axis] < 0
    }
    return swapped ? !lo_first : lo_first;
}

// ---- conservative f32 slab test against 8 quantised child boxes -------------------------------------------------
//
// True entry/exit distance of a plane c = p + q*2^e along one axis: t = (c - o) / d.  The kernel evaluates
//   t~ = fma(q, A, B),  A = 2^e * idir32,  B = (p - o32) * idir32
// whose error is bounded by  |B|*2^-22 (rounding of p - o32 and of the product)  +  |o|*2^-23*|idir| (o rounded
// to f32)  + a relative 2^-22 of t~ (rounding of idir32 = RN(1/d) and of the fma).  The first two are subtracted
// from B for entry planes and added for exit planes once per node; the relative term is applied by scaling A and
// B by (1 -+ 2^-21).  Child boxes themselves are rounded outward when quantised (bvh_build.cpp).
struct WideRay {
    float ox, oy, oz;
    float idx, idy, idz;
    float eox, eoy, eoz;   // |o| * 2^-23 * |idir|
    float tmax;            // ray max distance rounded up
    uint32_t octinv;       // 7 - octant, octant bit (4,2,1) set where dir (x,y,z) is negative
    uint32_t negx, negy, negz;
};

__device__ __forceinline__ float safe_rcp(double d) {
    const double lim = 1e-25;
    double a = fabs(d) < lim ? copysign(lim, d) : d;
    return (float)(1.0 / a);
}

__device__ __forceinline__ WideRay make_wide_ray(V3 o, V3 dir, double ray_max) {
    WideRay r;
    r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
    r.idx = safe_rcp(dir.x); r.idy = safe_rcp(dir.y); r.idz = safe_rcp(dir.z);
    const float k = 1.1920929e-7f;  // 2^-23
    r.eox = (fabsf(r.ox) * k + 1e-30f) * fabsf(r.idx);
    r.eoy = (fabsf(r.oy) * k + 1e-30f) * fabsf(r.idy);
    r.eoz = (fabsf(r.oz) * k + 1e-30f) * fabsf(r.idz);
    r.tmax = __double2float_ru(ray_max);
    r.negx = sign_negative(dir.x); r.negy = sign_negative(dir.y); r.negz = sign_negative(dir.z);
    r.octinv = 7u - ((r.negx << 2) | (r.negy << 1) | r.negz);
    return r;
}

// byte `sel` of word -> float(2^23 + byte) in one PRMT, then an exact subtraction of 2^23
__device__ __forceinline__ float byte_to_float(uint32_t word, uint32_t sel) {
    const uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7650u | sel);
    return __uint_as_float(bits) - 8388608.0f;
}

struct AxisPlanes {
    float A, Bn, Bf;
};
__device__ __forceinline__ AxisPlanes axis_planes(float p, float o, float id, float eo, uint32_t ebyte) {
    const float scale = __uint_as_float(ebyte << 23);
    const float B = (p - o) * id;
    const float err = fmaf(fabsf(B), 2.3841858e-7f /*2^-22*/, eo);
    AxisPlanes r;
    r.A = scale * id;   // exact: scale is a power of two
    r.Bn = B - err;     // entry distances may only get smaller
    r.Bf = B + err;     // exit distances only larger
    return r;
}

// Traversal state of one ray through the 8-wide BVH, advanced one step at a time so that the lanes of a warp can be
// kept in the same phase (see run_wide_persistent in wavefront.cu):
//   node_step  pop the nearest pending interior child, test its 8 quantised child boxes, queue the leaf primitives hit
//   prim_step  run the f64 intersection test of ONE queued primitive
//   advance    when both queues are empty, pop the traversal stack (or finish)
// Leaf primitives are postponed (kept in `tg`, spilled to the stack when a newer group arrives) until enough lanes of
// the warp have primitive work, which is what keeps the expensive f64 tests from running one lane at a time.
template <bool ANY>
struct WideTraversal {
    V3 o, dir;
    double ray_max;
    WideRay r;
    uint2 ng;          // node group: x = first interior child, y = hit bits (31..24, octant ordered) | imask (7..0)
    uint2 tg;          // primitive group: x = first leaf primitive of the node, y = mask of primitives still to test
    uint2 stack[kWideStack];
    int sp;
    Hit hit;
    uint32_t best_prim;
    bool live;

    __device__ __forceinline__ void begin(V3 origin, V3 direction, double max_distance) {
        o = origin; dir = direction; ray_max = max_distance;
        r = make_wide_ray(origin, direction, max_distance);
        ng = make_uint2(0u, 0x80000000u);  // the root, as a one-child node group
        tg = make_uint2(0u, 0u);
        sp = 0;
        hit.slot = CRAY_NO_HIT; hit.t = max_distance; hit.u = 0.0; hit.v = 0.0;
        best_prim = 0;
        live = true;
    }
    __device__ __forceinline__ bool has_node_work() const { return (ng.y & 0xFF000000u) != 0u; }
    __device__ __forceinline__ bool has_prim_work() const { return tg.y != 0u; }

    __device__ __forceinline__ void node_step(const SceneView& s) {
        const uint32_t bit = 31u - __clz(ng.y);
        ng.y &= ~(1u << bit);
        const uint32_t slot = (bit - 24u) ^ r.octinv;
        const uint32_t child = ng.x + __popc(ng.y & 0xFFu & ((1u << slot) - 1u));
        if ((ng.y & 0xFF000000u) && sp < kWideStack) stack[sp++] = ng;

        const int4* raw = reinterpret_cast<const int4*>(s.wide_nodes + child);
        const int4 n0 = __ldg(raw), n1 = __ldg(raw + 1), n2 = __ldg(raw + 2), n3 = __ldg(raw + 3), n4 = __ldg(raw + 4);
        // n0 = px, py, pz, {ex,ey,ez,imask};  n1 = child_base, prim_base, meta[0..3], meta[4..7]
        // n2 = qlo_x[0..7], qlo_y[0..7];  n3 = qlo_z[0..7], qhi_x[0..7];  n4 = qhi_y[0..7], qhi_z[0..7]
        const uint32_t e_imask = (uint32_t)n0.w;
        const AxisPlanes X = axis_planes(__int_as_float(n0.x), r.ox, r.idx, r.eox, e_imask & 0xFFu);
        const AxisPlanes Y = axis_planes(__int_as_float(n0.y), r.oy, r.idy, r.eoy, (e_imask >> 8) & 0xFFu);
        const AxisPlanes Z = axis_planes(__int_as_float(n0.z), r.oz, r.idz, r.eoz, (e_imask >> 16) & 0xFFu);
        const uint32_t imask = e_imask >> 24;
        // entry planes: lower bounds for positive directions, upper bounds for negative ones
        const uint32_t nx0 = r.negx ? (uint32_t)n3.z : (uint32_t)n2.x, nx1 = r.negx ? (uint32_t)n3.w : (uint32_t)n2.y;
        const uint32_t fx0 = r.negx ? (uint32_t)n2.x : (uint32_t)n3.z, fx1 = r.negx ? (uint32_t)n2.y : (uint32_t)n3.w;
        const uint32_t ny0 = r.negy ? (uint32_t)n4.x : (uint32_t)n2.z, ny1 = r.negy ? (uint32_t)n4.y : (uint32_t)n2.w;
        const uint32_t fy0 = r.negy ? (uint32_t)n2.z : (uint32_t)n4.x, fy1 = r.negy ? (uint32_t)n2.w : (uint32_t)n4.y;
        const uint32_t nz0 = r.negz ? (uint32_t)n4.z : (uint32_t)n3.x, nz1 = r.negz ? (uint32_t)n4.w : (uint32_t)n3.y;
        const uint32_t fz0 = r.negz ? (uint32_t)n3.x : (uint32_t)n4.z, fz1 = r.negz ? (uint32_t)n3.y : (uint32_t)n4.w;
        const uint32_t meta_lo = (uint32_t)n1.z, meta_hi = (uint32_t)n1.w;

        uint32_t interior_hits = 0, prim_hits = 0;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
            const uint32_t sel = sl & 3;
            const uint32_t meta = ((sl < 4 ? meta_lo : meta_hi) >> (8 * sel)) & 0xFFu;
            const float tnx = fmaf(byte_to_float(sl < 4 ? nx0 : nx1, sel), X.A, X.Bn);
            const float tny = fmaf(byte_to_float(sl < 4 ? ny0 : ny1, sel), Y.A, Y.Bn);
            const float tnz = fmaf(byte_to_float(sl < 4 ? nz0 : nz1, sel), Z.A, Z.Bn);
            const float tfx = fmaf(byte_to_float(sl < 4 ? fx0 : fx1, sel), X.A, X.Bf);
            const float tfy = fmaf(byte_to_float(sl < 4 ? fy0 : fy1, sel), Y.A, Y.Bf);
            const float tfz = fmaf(byte_to_float(sl < 4 ? fz0 : fz1, sel), Z.A, Z.Bf);
            // relative slack 2^-21 for the rounding of idir and of the fmas: shrink the (non-negative) entry
            // distance, grow the exit distance (a negative exit distance only becomes more negative: still a miss)
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f)) * (1.0f - 4.7683716e-7f);
            const float tf = fminf(fminf(fminf(tfx, tfy), tfz) * (1.0f + 4.7683716e-7f), r.tmax);
            if (meta != 0u && tn <= tf) {
                if (meta == 0xE0u) interior_hits |= 1u << (24u + ((uint32_t)sl ^ r.octinv));
                else prim_hits |= ((1u << (meta >> 5)) - 1u) << (meta & 31u);
            }
        }
        if (prim_hits) {
            if (tg.y && sp < kWideStack) stack[sp++] = make_uint2(tg.x | 0x80000000u, tg.y);  // older group waits on the stack
            tg = make_uint2((uint32_t)n1.y, prim_hits);
        }
        ng = make_uint2((uint32_t)n1.x, interior_hits | imask);
    }

    // Tests one queued primitive.  Returns true if the ray is finished by it (any-hit mode found an occluder).
    __device__ __forceinline__ bool prim_step(const SceneView& s) {
        const uint32_t k = __ffs(tg.y) - 1u;
        tg.y &= tg.y - 1u;
        const uint32_t slot = tg.x + k;
        const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
        if (ANY) return leaf_prim_any(s, lp, o, dir, ray_max);
        double u, v;
        const int verdict = leaf_prim_candidate(s, lp, o, dir, ray_max, hit.slot != CRAY_NO_HIT, u, v);
        if (verdict == 1) {
            hit.t = ray_max; hit.u = u; hit.v = v; hit.slot = slot;
            best_prim = lp.prim;
            r.tmax = __double2float_ru(ray_max);
        } else if (verdict == 2 && lp.prim != best_prim) {
            // exact-t tie: keep whichever primitive the reference's traversal order reaches first
            if (reference_visits_first(s, s.rank_of_prim[lp.prim], s.rank_of_prim[best_prim], dir)) {
                hit.u = u; hit.v = v; hit.slot = slot;
                best_prim = lp.prim;
            }
        }
        return false;
    }

    // Refill the work queues from the stack.  Returns true when the traversal is complete.
    __device__ __forceinline__ bool advance() {
        if (has_node_work() || has_prim_work()) return false;
        if (sp == 0) return true;
        const uint2 e = stack[--sp];
        if (e.x & 0x80000000u) tg = make_uint2(e.x & 0x7FFFFFFFu, e.y);
        else ng = e;
        return false;
    }
};

// Single-ray driver (no warp cooperation): used where only a handful of rays are traced.
template <bool ANY>
__device__ __forceinline__ bool traverse_wide(const SceneView& s, V3 o, V3 dir, double ray_max, Hit& hit) {
    WideTraversal<ANY> t;
    t.begin(o, dir, ray_max);
    for (;;) {
        if (t.has_prim_work()) {
            if (t.prim_step(s)) return true;
        } else if (t.has_node_work()) {
            t.node_step(s);
        } else if (t.advance()) {
            break;
        }
    }
    hit = t.hit;
    return hit.slot != CRAY_NO_HIT;
}

}  // namespace cray
