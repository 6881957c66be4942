// BVH traversal device functions.
//
//  traverse_exact<ANY>: the reference's Bvh::intersect / Bvh::intersects (src/bvh.rs:58-147) on the flattened
//    binary SAH tree: f64 slab test with the reference's accept rule (bounds.rs:62-88) or-ed with
//    Bounds::contains, near/far push order by sign of dir[split_axis], strict `t < max_distance` acceptance.
//    Results are bit-identical to the reference by construction (including its false box misses).
//
//  Wide traversal (production): the 8-wide quantised BVH.  Boxes are tested in f32 with every rounding error
//    pushed outward (see "conservative slab test" below) so no primitive the exact f64 test would accept is ever
//    culled; the primitives whose own quantised box the ray enters are queued per warp and tested in f64 by all
//    32 lanes at once (WarpShared, prim_round_*).  Among primitives with bit-equal t the one the reference would
//    have visited first wins (reference_visits_first).
#pragma once
#include "shapes.cuh"

namespace cray {

constexpr int kExactStack = 96;
#ifndef CRAY_NODE_PAIR
#define CRAY_NODE_PAIR 0   // 1: two nodes per lane and step (node_step_pair)
#endif
// node groups: at most one entry per level, the builder rejects deeper trees; paired steps may hold kWideStackLimit more
constexpr int kWideStack = CRAY_NODE_PAIR ? 2 * kWideStackLimit : kWideStackLimit;

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

// Does `o` lie in the outer shell of a marked node box (bvh_build.hpp "planar contact"): outside the box -- Bounds::contains
// (bounds.rs:46-53) is false -- but by no more than `tol` on every axis, so that the slab entry distance of a ray of direction
// length <= tol / 1e-9 can be <= 1e-9?  Only such rays can be culled by the reference's box test where a conservative traversal
// would go on (bvh.rs:70,:117 with bounds.rs:62-88); the caller traces them with traverse_exact.  Walks only the part of the binary
// tree that holds marked nodes and whose grown boxes contain `o`.
__device__ __noinline__ bool origin_in_contact_shell(const SceneView& s, V3 o, double tol) {
    if (!s.bin_contact) return false;
    uint32_t stack[kExactStack];
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        const uint32_t flags = s.bin_contact[ni];
        if (!(flags & CONTACT_BELOW)) continue;
        const BinNode* np = s.bin_nodes + ni;
        const int4* raw = reinterpret_cast<const int4*>(np);
        BinNode n;
        int4* dst = reinterpret_cast<int4*>(&n);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = __ldg(raw + i);
        if (!(n.box.lo.x - tol <= o.x && n.box.lo.y - tol <= o.y && n.box.lo.z - tol <= o.z && n.box.hi.x + tol >= o.x && n.box.hi.y + tol >= o.y &&
              n.box.hi.z + tol >= o.z))
            continue;
        if ((flags & CONTACT_NODE) && !bounds_contains(n.box, o)) return true;
        if (n.axis != 3 && sp + 2 <= kExactStack) { stack[sp++] = n.b; stack[sp++] = n.a; }
    }
    return false;
}
// ... with the tolerance of a ray of direction `dir` (its largest component scales the reference's 1e-9)
__device__ __forceinline__ double contact_tol(V3 dir) { return kContactTol * fmax(1.0, fmax(fabs(dir.x), fmax(fabs(dir.y), fabs(dir.z)))); }

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
//   t~ = fma(K + q, A, B'),   A = 2^e * idir32,   B' = B - K*A,   B = (p - o32) * idir32,   K = 2^15
// (K + q is what one PRMT makes of the byte q: the f32 bit pattern 0x47000000 | q << 8).  Absolute error sources:
// rounding of p - o32 and of its product with idir32 (<= |B| * 2^-23), o rounded to f32 (<= |o| * 2^-24 * |idir|),
// rounding of B - err and of B' (<= 2^-23 * (|B| + K|A|)).  Their sum, with margin,
//   err = |B| * 2^-21 + |o32| * 2^-23 * |idir32| + |A| * 2^-7
// is subtracted from B for entry planes and added for exit planes once per node and axis.  What remains is a relative
// error of at most 2^-22 per distance (idir32 = RN(1/RN(d)), the final fma), covered by accepting a child when
//   max(entry distances, 0) <= min(exit distances) * (1 + 2^-19)   and   <= tmax * (1 + 2^-19).
// Child boxes themselves are rounded outward when quantised (bvh_build.cpp).
struct WideRay {
    float ox, oy, oz;
    float idx, idy, idz;
    float eox, eoy, eoz;   // |o32| * 2^-23 * |idir32|
    float tmax;            // (current max distance rounded up) * (1 + 2^-19)
    uint32_t octinv;       // 7 - octant, octant bit (4,2,1) set where dir (x,y,z) is negative
};

constexpr float kSlabSlack = 1.0f + 1.9073486e-6f;  // 1 + 2^-19

__device__ __forceinline__ float safe_rcp(double d) {
    float f = (float)d;
    if (fabsf(f) < 1e-25f) f = sign_negative(d) ? -1e-25f : 1e-25f;
    return __frcp_rn(f);
}
__device__ __forceinline__ float slab_tmax(double ray_max) { return __fmul_ru(__double2float_ru(ray_max), kSlabSlack); }

__device__ __forceinline__ WideRay make_wide_ray(V3 o, V3 dir, double ray_max) {
    WideRay r;
    r.ox = (float)o.x; r.oy = (float)o.y; r.oz = (float)o.z;
    r.idx = safe_rcp(dir.x); r.idy = safe_rcp(dir.y); r.idz = safe_rcp(dir.z);
    const float k = 1.1920929e-7f;  // 2^-23
    r.eox = (fabsf(r.ox) * k + 1e-30f) * fabsf(r.idx);
    r.eoy = (fabsf(r.oy) * k + 1e-30f) * fabsf(r.idy);
    r.eoz = (fabsf(r.oz) * k + 1e-30f) * fabsf(r.idz);
    r.tmax = slab_tmax(ray_max);
    const uint32_t oct = ((uint32_t)sign_negative(dir.x) << 2) | ((uint32_t)sign_negative(dir.y) << 1) | (uint32_t)sign_negative(dir.z);
    r.octinv = 7u - oct;
    return r;
}

// byte `sel` of word -> the f32 2^15 + byte, in one PRMT
__device__ __forceinline__ float byte_to_float_k(uint32_t word, uint32_t sel) {
    return __uint_as_float(__byte_perm(word, 0x47000000u, 0x7604u | (sel << 4)));
}

#ifndef CRAY_FFMA2
#define CRAY_FFMA2 1   // 0: scalar FFMA slab test (tuning builds)
#endif
// d = a * b + c on two f32 lanes at once (sm_100: fma.rn.f32x2, SASS FFMA2); each lane rounds exactly like fmaf
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

struct AxisPlanes {
    float A, Bn, Bf;
};
__device__ __forceinline__ AxisPlanes axis_planes(float p, float o, float id, float eo, uint32_t ebyte) {
    const float scale = __uint_as_float(ebyte << 23);
    AxisPlanes r;
    r.A = scale * id;   // exact: scale is a power of two
    const float B = (p - o) * id;
    const float err = fmaf(fabsf(B), 4.7683716e-7f /*2^-21*/, fmaf(fabsf(r.A), 7.8125e-3f /*2^-7*/, eo));
    const float KA = 32768.0f * r.A;  // exact
    r.Bn = (B - err) - KA;  // entry distances may only get smaller
    r.Bf = (B + err) - KA;  // exit distances only larger
    return r;
}

// ---- per-warp shared state of the persistent traversal kernel -----------------------------------------------------------
//
// Each lane owns one ray.  Its f64 origin / direction and its current closest hit live here, so that ANY lane of the warp
// can run a queued primitive test for it.  `queue` is a ring of pending tests (owner lane << 27 | leaf slot).
constexpr uint32_t kQueue = 512;        // >= 31 left over + 32 lanes x 8 leaf slots pushed by one node phase
constexpr uint32_t kSlotBits = 27;
constexpr uint32_t kSlotMask = (1u << kSlotBits) - 1u;

struct WarpShared {
    double ox[32], oy[32], oz[32], dx[32], dy[32], dz[32];
    double tmax[32];      // closest-hit: distance of the best hit so far (ray.max_distance); any-hit: the ray's max distance
    float tmax32[32];     // slab_tmax(tmax)
    uint32_t best[32];    // closest-hit: leaf slot of the best hit, CRAY_NO_HIT if none; any-hit: 1 once occluded
    uint32_t pend[32];    // queued tests of this lane's ray that have not run yet
    uint32_t tail;        // entries pushed so far (ring position of the next push)
    uint32_t _pad[3];
    uint32_t queue[kQueue];
};

__device__ __forceinline__ void warp_begin_ray(WarpShared& ws, unsigned lane, V3 o, V3 dir, double ray_max, uint32_t best_init) {
    ws.ox[lane] = o.x; ws.oy[lane] = o.y; ws.oz[lane] = o.z;
    ws.dx[lane] = dir.x; ws.dy[lane] = dir.y; ws.dz[lane] = dir.z;
    ws.tmax[lane] = ray_max;
    ws.tmax32[lane] = slab_tmax(ray_max);
    ws.best[lane] = best_init;
    ws.pend[lane] = 0u;
}

// ---- node phase ------------------------------------------------------------------------------------------------------------
//
// A node group (uint2) = { index of the first interior child of some node, hit interior children << 24 | imask }: the children of
// one node the ray still has to visit, nearest first (bit 31 = nearest in octant order).
struct NodeData {
    int4 n0, n1, n2, n3, n4;
    // n0 = px, py, pz, {ex,ey,ez,imask};  n1 = child_base, prim_base, {leafmask, pad}, pad
    // n2 = qlo_x[0..7], qlo_y[0..7];  n3 = qlo_z[0..7], qhi_x[0..7];  n4 = qhi_y[0..7], qhi_z[0..7]
};

// Removes the nearest pending child from the group and returns its node index.
__device__ __forceinline__ uint32_t pop_child(uint2& g, uint32_t octinv) {
    const uint32_t bit = 31u - __clz(g.y);
    g.y &= ~(1u << bit);
    const uint32_t slot = (bit - 24u) ^ octinv;
    return g.x + __popc(g.y & 0xFFu & ((1u << slot) - 1u));
}

__device__ __forceinline__ NodeData load_node(const SceneView& s, uint32_t index) {
    NodeData n;
#if CRAY_NODE96
    // sm_100: 256-bit global loads (LDG.E.ENL2.256), three per 96-byte node instead of five 128-bit ones per 80-byte node
    const char* raw = reinterpret_cast<const char*>(s.wide_nodes + index);
    int4 pad;
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(n.n0.x), "=r"(n.n0.y), "=r"(n.n0.z), "=r"(n.n0.w), "=r"(n.n1.x), "=r"(n.n1.y), "=r"(n.n1.z), "=r"(n.n1.w) : "l"(raw));
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(n.n2.x), "=r"(n.n2.y), "=r"(n.n2.z), "=r"(n.n2.w), "=r"(n.n3.x), "=r"(n.n3.y), "=r"(n.n3.z), "=r"(n.n3.w) : "l"(raw + 32));
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(n.n4.x), "=r"(n.n4.y), "=r"(n.n4.z), "=r"(n.n4.w), "=r"(pad.x), "=r"(pad.y), "=r"(pad.z), "=r"(pad.w) : "l"(raw + 64));
    (void)pad;
#else
    const int4* raw = reinterpret_cast<const int4*>(s.wide_nodes + index);
    n.n0 = __ldg(raw); n.n1 = __ldg(raw + 1); n.n2 = __ldg(raw + 2); n.n3 = __ldg(raw + 3); n.n4 = __ldg(raw + 4);
#endif
    return n;
}

// The conservative slab test of the ray against the 8 quantised child boxes of one node: bit s of the result = slot s is hit.
__device__ __forceinline__ uint32_t test_node(const NodeData& n, const WideRay& r) {
    const int4 n0 = n.n0, n2 = n.n2, n3 = n.n3, n4 = n.n4;
    const uint32_t e_imask = (uint32_t)n0.w;
    const AxisPlanes X = axis_planes(__int_as_float(n0.x), r.ox, r.idx, r.eox, e_imask & 0xFFu);
    const AxisPlanes Y = axis_planes(__int_as_float(n0.y), r.oy, r.idy, r.eoy, (e_imask >> 8) & 0xFFu);
    const AxisPlanes Z = axis_planes(__int_as_float(n0.z), r.oz, r.idz, r.eoz, (e_imask >> 16) & 0xFFu);
    // entry planes: lower bounds for positive directions, upper bounds for negative ones
    const bool negx = !(r.octinv & 4u), negy = !(r.octinv & 2u), negz = !(r.octinv & 1u);
    const uint32_t nx0 = negx ? (uint32_t)n3.z : (uint32_t)n2.x, nx1 = negx ? (uint32_t)n3.w : (uint32_t)n2.y;
    const uint32_t fx0 = negx ? (uint32_t)n2.x : (uint32_t)n3.z, fx1 = negx ? (uint32_t)n2.y : (uint32_t)n3.w;
    const uint32_t ny0 = negy ? (uint32_t)n4.x : (uint32_t)n2.z, ny1 = negy ? (uint32_t)n4.y : (uint32_t)n2.w;
    const uint32_t fy0 = negy ? (uint32_t)n2.z : (uint32_t)n4.x, fy1 = negy ? (uint32_t)n2.w : (uint32_t)n4.y;
    const uint32_t nz0 = negz ? (uint32_t)n4.z : (uint32_t)n3.x, nz1 = negz ? (uint32_t)n4.w : (uint32_t)n3.y;
    const uint32_t fz0 = negz ? (uint32_t)n3.x : (uint32_t)n4.z, fz1 = negz ? (uint32_t)n3.y : (uint32_t)n4.w;

    uint32_t hits = 0;
#if CRAY_FFMA2
    // two slots per packed FFMA2 (sm_100 fma.rn.f32x2): half the FMA issue slots of the scalar loop
    const float2 XA = make_float2(X.A, X.A), XBn = make_float2(X.Bn, X.Bn), XBf = make_float2(X.Bf, X.Bf);
    const float2 YA = make_float2(Y.A, Y.A), YBn = make_float2(Y.Bn, Y.Bn), YBf = make_float2(Y.Bf, Y.Bf);
    const float2 ZA = make_float2(Z.A, Z.A), ZBn = make_float2(Z.Bn, Z.Bn), ZBf = make_float2(Z.Bf, Z.Bf);
#pragma unroll
    for (int sl = 0; sl < 8; sl += 2) {
        const uint32_t s0 = sl & 3, s1 = (sl + 1) & 3;
        const float2 tnx = fma2(make_float2(byte_to_float_k(sl < 4 ? nx0 : nx1, s0), byte_to_float_k(sl < 4 ? nx0 : nx1, s1)), XA, XBn);
        const float2 tny = fma2(make_float2(byte_to_float_k(sl < 4 ? ny0 : ny1, s0), byte_to_float_k(sl < 4 ? ny0 : ny1, s1)), YA, YBn);
        const float2 tnz = fma2(make_float2(byte_to_float_k(sl < 4 ? nz0 : nz1, s0), byte_to_float_k(sl < 4 ? nz0 : nz1, s1)), ZA, ZBn);
        const float2 tfx = fma2(make_float2(byte_to_float_k(sl < 4 ? fx0 : fx1, s0), byte_to_float_k(sl < 4 ? fx0 : fx1, s1)), XA, XBf);
        const float2 tfy = fma2(make_float2(byte_to_float_k(sl < 4 ? fy0 : fy1, s0), byte_to_float_k(sl < 4 ? fy0 : fy1, s1)), YA, YBf);
        const float2 tfz = fma2(make_float2(byte_to_float_k(sl < 4 ? fz0 : fz1, s0), byte_to_float_k(sl < 4 ? fz0 : fz1, s1)), ZA, ZBf);
        const float tn0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, 0.0f)), tn1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, 0.0f));
        const float tf0 = fminf(fminf(fminf(tfx.x, tfy.x), tfz.x) * kSlabSlack, r.tmax);
        const float tf1 = fminf(fminf(fminf(tfx.y, tfy.y), tfz.y) * kSlabSlack, r.tmax);
        hits |= (tn0 <= tf0 ? 1u : 0u) << sl;
        hits |= (tn1 <= tf1 ? 1u : 0u) << (sl + 1);
    }
#else
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) {
        const uint32_t sel = sl & 3;
        const float tnx = fmaf(byte_to_float_k(sl < 4 ? nx0 : nx1, sel), X.A, X.Bn);
        const float tny = fmaf(byte_to_float_k(sl < 4 ? ny0 : ny1, sel), Y.A, Y.Bn);
        const float tnz = fmaf(byte_to_float_k(sl < 4 ? nz0 : nz1, sel), Z.A, Z.Bn);
        const float tfx = fmaf(byte_to_float_k(sl < 4 ? fx0 : fx1, sel), X.A, X.Bf);
        const float tfy = fmaf(byte_to_float_k(sl < 4 ? fy0 : fy1, sel), Y.A, Y.Bf);
        const float tfz = fmaf(byte_to_float_k(sl < 4 ? fz0 : fz1, sel), Z.A, Z.Bf);
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
        const float tf = fminf(fminf(fminf(tfx, tfy), tfz) * kSlabSlack, r.tmax);
        hits |= (tn <= tf ? 1u : 0u) << sl;
    }
#endif
    return hits;
}

// The interior children hit, as a node group: the bit of slot sl moves to position sl ^ octinv (nearest child in the highest bit).
__device__ __forceinline__ uint2 hit_group(const NodeData& n, uint32_t hits, uint32_t octinv) {
    const uint32_t imask = (uint32_t)n.n0.w >> 24;
    uint32_t ih = hits & imask;
    if (octinv & 1u) ih = ((ih & 0x55u) << 1) | ((ih & 0xAAu) >> 1);
    if (octinv & 2u) ih = ((ih & 0x33u) << 2) | ((ih & 0xCCu) >> 2);
    if (octinv & 4u) ih = ((ih & 0x0Fu) << 4) | ((ih & 0xF0u) >> 4);
    return make_uint2((uint32_t)n.n1.x, (ih << 24) | imask);
}

// Queues the primitives of the leaf slots in `lh` (one per slot, contiguous from prim_base in ascending slot order).
template <class WS>
__device__ __forceinline__ void push_prims(WS& ws, unsigned lane, uint32_t lh, uint32_t leafmask, uint32_t prim_base) {
    if (lh) {
        const uint32_t n = __popc(lh);
        uint32_t pos = atomicAdd(&ws.tail, n);
        ws.pend[lane] += n;
        const uint32_t tag = lane << kSlotBits;
        do {
            const uint32_t k = __ffs(lh) - 1u;
            lh &= lh - 1u;
            const uint32_t leaf_slot = prim_base + __popc(leafmask & ((1u << k) - 1u));
            ws.queue[pos & (kQueue - 1u)] = tag | leaf_slot;
            pos += 1u;
        } while (lh);
    }
}

// One interior node for one ray: pops the nearest pending child of the node group `ng`, tests its 8 quantised child boxes,
// leaves the interior children hit in `ng` (octant ordered) and queues the primitives whose box was hit.
template <class WS>
__device__ __forceinline__ uint32_t node_step(const SceneView& s, WS& ws, unsigned lane, const WideRay& r, uint2& ng, uint2* stack, int& sp) {
    const uint32_t child = pop_child(ng, r.octinv);
    if ((ng.y & 0xFF000000u) && sp < kWideStack) stack[sp++] = ng;
    const NodeData n = load_node(s, child);
    const uint32_t hits = test_node(n, r);
    ng = hit_group(n, hits, r.octinv);
    const uint32_t leafmask = (uint32_t)n.n1.z & 0xFFu, lh = hits & leafmask;
    push_prims(ws, lane, lh, leafmask, (uint32_t)n.n1.y);
    return (ng.y >> 24) | (lh << 8);   // (work counters of the stats build)
}

// Two nodes for one ray in one step (CRAY_NODE_PAIR): the nearest pending child A and the next one in visit order, B -- the
// second child of the same group or, if A was its last, the nearest child of the group on top of the stack.  Both records are
// requested before either is used, so a lane has two dependent-fetch latencies in flight instead of one.  Visit order stays
// depth-first, nearest first: A's children are visited before B's, B's before the rest of B's group.  B is visited a step early
// (before A's primitives could have shortened the ray) -- the box test it gets is the one it would have got in the same
// iteration anyway, since primitive rounds run after the node phase.
// Lanes without a second pending child run the B half on A's record with its result masked off (no divergence).
constexpr int kPairStackLimit = kWideStackLimit;   // a pair is only taken while sp + 2 <= this: the unpaired descent below needs
                                                   // at most one more entry per level, kWideStack = this + kWideStackLimit holds both
template <class WS>
__device__ __forceinline__ uint32_t node_step_pair(const SceneView& s, WS& ws, unsigned lane, const WideRay& r, uint2& ng, uint2* stack, int& sp) {
    const uint32_t a = pop_child(ng, r.octinv);
    uint32_t b = a;
    bool have_b = false;
    if (sp + 2 <= kPairStackLimit) {
        if (ng.y & 0xFF000000u) {
            b = pop_child(ng, r.octinv);
            have_b = true;
        } else if (sp > 0) {
            uint2 g = stack[sp - 1];
            b = pop_child(g, r.octinv);
            have_b = true;
            if (g.y & 0xFF000000u) stack[sp - 1] = g;
            else --sp;
        }
    }
    if ((ng.y & 0xFF000000u) && sp < kWideStack) stack[sp++] = ng;
    const NodeData na = load_node(s, a), nb = load_node(s, b);
    const uint32_t hits_a = test_node(na, r);
    uint32_t hits_b = have_b ? test_node(nb, r) : 0u;
    // primitives of both nodes in one pass: slots 0..7 = A's, 8..15 = B's
    const uint32_t lm_a = (uint32_t)na.n1.z & 0xFFu, lm_b = (uint32_t)nb.n1.z & 0xFFu;
    uint32_t lh = (hits_a & lm_a) | ((hits_b & lm_b) << 8);
    uint2 gb = hit_group(nb, hits_b, r.octinv);
    if (__popc(lh) > 15) {
        // the test queue holds 31 left-over entries + 15 per lane: a lane whose two nodes hit all 16 of their primitives puts B
        // back as a group of its own (visited again later)
        lh &= 0xFFu;
        hits_b = 0u;
        gb = make_uint2(b, 0x80000000u);
    }
    if ((gb.y & 0xFF000000u) && sp < kWideStack) stack[sp++] = gb;
    ng = hit_group(na, hits_a, r.octinv);
    const uint32_t outcome = (ng.y >> 24) | ((lh & 0xFFu) << 8) | (have_b ? 0x10000u | ((gb.y >> 24) << 24) | ((lh >> 8) ? 0x20000u : 0u) : 0u);
    if (lh) {
        const uint32_t n = __popc(lh);
        uint32_t pos = atomicAdd(&ws.tail, n);
        ws.pend[lane] += n;
        const uint32_t tag = lane << kSlotBits;
        do {
            const uint32_t k = __ffs(lh) - 1u;
            lh &= lh - 1u;
            const bool second = k >= 8u;
            const uint32_t leaf_slot = (second ? (uint32_t)nb.n1.y : (uint32_t)na.n1.y) + __popc((second ? lm_b : lm_a) & ((1u << (k & 7u)) - 1u));
            ws.queue[pos & (kQueue - 1u)] = tag | leaf_slot;
            pos += 1u;
        } while (lh);
    }
    return outcome;
}

// Runs the first min(count, 32) queued tests, one per lane (closest-hit flavour).  The lanes testing primitives for the same
// ray form a group (__match_any_sync); the group's smallest distance, if it beats the ray's current one, becomes the new best.
// Exact-t ties -- with the current best or inside the group -- take the slow path that asks which primitive the reference's
// traversal reaches first.
__device__ __forceinline__ void prim_round_closest(const SceneView& s, WarpShared& ws, unsigned lane, uint32_t head, uint32_t count) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool act = lane < count;
    const uint32_t e = act ? ws.queue[(head + lane) & (kQueue - 1u)] : 0u;
    const uint32_t owner = e >> kSlotBits, slot = e & kSlotMask;
    int verdict = 0;
    double t = 0.0;
    if (act) {
        const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
        const V3 o = mk(ws.ox[owner], ws.oy[owner], ws.oz[owner]);
        const V3 dir = mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]);
        const double rmax = ws.tmax[owner];
        double cand = rmax, u, v;
        verdict = leaf_prim_candidate(s, lp, o, dir, cand, ws.best[owner] != CRAY_NO_HIT, u, v, s.wide_prims + slot);
        t = verdict == 1 ? cand : rmax;
    }
    // accepted distances are positive f64: they order like their bit patterns, so one shared-memory atomicMin per
    // accepting lane leaves the ray's new closest distance in ws.tmax (a verdict-2 lane's t IS the current value)
    if (verdict == 1) atomicMin(reinterpret_cast<unsigned long long*>(&ws.tmax[owner]), (unsigned long long)__double_as_longlong(t));
    __syncwarp();
    const unsigned grp = __match_any_sync(FULL, act ? owner : 32u + lane);
    const bool win = verdict != 0 && ws.tmax[owner] == t;
    const unsigned wins = __ballot_sync(FULL, win) & grp;
    // exact-t ties: several winners in this round, or (verdict 2) a winner level with the best of an earlier round
    const bool tie = win && (__popc(wins) > 1 || verdict == 2);
    if (win && !tie) {
        ws.tmax32[owner] = slab_tmax(t);
        ws.best[owner] = slot;
    }
    if (act && lane == (unsigned)(__ffs(grp) - 1)) ws.pend[owner] -= __popc(grp);
    unsigned ties = __ballot_sync(FULL, tie);
    while (ties) {  // rare: one tied lane at a time, in lane order
        const unsigned l = __ffs(ties) - 1u;
        ties &= ties - 1u;
        if (lane == l) {
            // the first verdict-1 winner of a group displaces the (farther) previous best; everyone else is compared with
            // the primitive currently holding this distance
            bool take = verdict == 1 && lane == (unsigned)(__ffs(wins) - 1);
            if (!take) {  // (the rare path re-reads what it needs instead of keeping it in registers through every round)
                const uint32_t cur = ws.best[owner];
                const V3 dir = mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]);
                take = cur != slot && reference_visits_first(s, s.rank_of_prim[s.wide_prims[slot].prim], s.rank_of_prim[s.wide_prims[cur].prim], dir);
            }
            if (take) ws.best[owner] = slot;
            ws.tmax32[owner] = slab_tmax(t);
        }
        __syncwarp();
    }
    __syncwarp();
}

// Any-hit flavour: a test that finds an occluder marks the ray; tests queued for a ray already marked are skipped.
__device__ __forceinline__ void prim_round_any(const SceneView& s, WarpShared& ws, unsigned lane, uint32_t head, uint32_t count) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool act = lane < count;
    const uint32_t e = act ? ws.queue[(head + lane) & (kQueue - 1u)] : 0u;
    const uint32_t owner = e >> kSlotBits, slot = e & kSlotMask;
    if (act && ws.best[owner] == 0u) {
        const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
        const V3 o = mk(ws.ox[owner], ws.oy[owner], ws.oz[owner]);
        const V3 dir = mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]);
#if CRAY_PRIM_NOCOPY
        const bool occ = (lp.kind & 0xFFu) == PRIM_TRIANGLE ? leaf_prim_any(s, lp, o, dir, ws.tmax[owner]) : analytic_any_at(s, s.wide_prims + slot, o, dir, ws.tmax[owner]);
        if (occ) ws.best[owner] = 1u;
#else
        if (leaf_prim_any(s, lp, o, dir, ws.tmax[owner])) ws.best[owner] = 1u;
#endif
    }
    const unsigned grp = __match_any_sync(FULL, act ? owner : 32u + lane);
    if (act && lane == (unsigned)(__ffs(grp) - 1)) ws.pend[owner] -= __popc(grp);
    __syncwarp();
}

// ---- F32 mode: the same traversal, triangles tested in f32 ---------------------------------------------------
//
// Watertight ray / triangle test of Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection" (JCGT 2013): the vertices are
// translated to the ray origin, sheared so that the ray runs along +z of a permuted frame, and the three 2-D edge functions
// decide; a shared edge gets the same rounded products from both of its triangles (this file is compiled with --fmad=false), so
// no ray slips between two triangles of a closed mesh, and an edge function that rounds to zero is re-evaluated in f64, where
// the products of f32 values are exact.  The translation is done against the f64 origin split into an f32 head and an f32
// remainder, (v - head) - rest, so that geometry near the origin is resolved relative to its distance from the origin and not
// to the size of the scene's coordinates: what a ray leaving a surface needs.  Per ray (warp_begin_ray32): the permutation kz
// and the shear Sx, Sy, Sz.  Both sides of a triangle are hit (the reference does not cull, shape.rs:227-248).
//
// Shared memory is budgeted: at 8 CTAs per SM, 24 064 bytes per CTA is the most that still fits the 196 KB carve-out (the next
// step, 228 KB, leaves 28 KB of L1 and costs the closest-hit kernel 15 %).
struct WarpShared32 {
    float4 ra[32];        // origin head x, y, z | Sx
    float4 rb[32];        // origin remainder x, y, z | Sy
    float4 rc[32];        // Sz | kz (bits) | leaf slot of the triangle the ray leaves, CRAY_NO_HIT if none (bits) | closest distance so far
    double ox[32], oy[32], oz[32], dx[32], dy[32], dz[32];   // the f64 ray, for spheres, disks and confirmations
    double tmax[32];      // closest-hit: distance of the best hit (exact f64 for spheres / disks); any-hit: the ray's max distance
    float tmax32[32];     // slab-test bound
    float tmin[32];       // smallest distance accepted (kEpsilon32, or the caller's allowance for a ray of unknown provenance)
    uint32_t best[32];
    uint32_t pend[32];
    uint32_t tail;
    uint32_t _pad[3];
    uint32_t queue[kQueue];
};

constexpr float kEpsilon32 = 1e-9f;   // Ray::contains_distance's lower bound (ray.rs:26)

// S3 rays carry no record of the surface they start on, and in f32 a triangle lies up to an ulp of its coordinates off its f64
// plane: distances below 2^-19 of the origin's largest coordinate (in units of the direction's largest component) are taken for
// the ray's own surface.  The wavefront knows the triangle a ray leaves and skips exactly that one instead.
__device__ __forceinline__ float f32_unknown_origin_tmin(V3 o, V3 dir) {
    const float om = fmaxf(fmaxf(fabsf((float)o.x), fabsf((float)o.y)), fabsf((float)o.z));
    const float dm = fmaxf(fmaxf(fabsf((float)dir.x), fabsf((float)dir.y)), fabsf((float)dir.z));
    return fmaxf(kEpsilon32, 1.9073486e-6f * om / fmaxf(dm, 1e-30f));
}

// Closest hit with a finite max distance (S3 rays only): pulled in by 2^-18 relative, so that a surface AT the end of the ray is
// not found a few ulps in front of it.  (Shadow rays get the exact treatment, see prim_round_any32.)
__device__ __forceinline__ float f32_ray_max(double ray_max) {
    const float m = __double2float_rd(ray_max);
    return isinf(m) ? m : __fmul_rd(m, 1.0f - 3.8146973e-6f);
}

template <bool ANY>
__device__ __forceinline__ void warp_begin_ray32(WarpShared32& ws, unsigned lane, V3 o, V3 dir, double ray_max, uint32_t best_init, uint32_t self_slot, float tmin) {
    ws.tmin[lane] = tmin;
    ws.ox[lane] = o.x; ws.oy[lane] = o.y; ws.oz[lane] = o.z;
    ws.dx[lane] = dir.x; ws.dy[lane] = dir.y; ws.dz[lane] = dir.z;
    ws.tmax[lane] = ray_max;
    ws.tmax32[lane] = slab_tmax(ray_max);
    ws.best[lane] = best_init;
    ws.pend[lane] = 0u;
    const float hx = (float)o.x, hy = (float)o.y, hz = (float)o.z;
    const float fx = (float)dir.x, fy = (float)dir.y, fz = (float)dir.z;
    const float ax = fabsf(fx), ay = fabsf(fy), az = fabsf(fz);
    const uint32_t kz = ax > ay ? (ax > az ? 0u : 2u) : (ay > az ? 1u : 2u);
    // (kx, ky, kz) is a cyclic shift of (x, y, z)
    const float dk = kz == 0u ? fx : (kz == 1u ? fy : fz);
    const float dkx = kz == 0u ? fy : (kz == 1u ? fz : fx);
    const float dky = kz == 0u ? fz : (kz == 1u ? fx : fy);
    ws.ra[lane] = make_float4(hx, hy, hz, dkx / dk);
    ws.rb[lane] = make_float4((float)(o.x - (double)hx), (float)(o.y - (double)hy), (float)(o.z - (double)hz), dky / dk);
    ws.rc[lane] = make_float4(1.0f / dk, __uint_as_float(kz), __uint_as_float(self_slot), ANY ? __double2float_ru(ray_max) : f32_ray_max(ray_max));
}

__device__ __forceinline__ float& closest32(WarpShared32& ws, uint32_t owner) { return ws.rc[owner].w; }

struct Tri32Regs { float4 a, b, c; };   // v0.xyz v1.x | v1.yz v2.xy | v2.z prim kind pad
__device__ __forceinline__ Tri32Regs load_tri32(const Tri32* p) {
    const float4* src = reinterpret_cast<const float4*>(p);
    Tri32Regs r;
    r.a = __ldg(src); r.b = __ldg(src + 1); r.c = __ldg(src + 2);
    return r;
}

// true iff the ray of `owner` crosses the triangle; t = distance along the (un-normalised) direction
__device__ __forceinline__ bool tri32_hit(const Tri32Regs& tr, const WarpShared32& ws, uint32_t owner, float& t) {
    const float4 ra = ws.ra[owner], rb = ws.rb[owner], rc = ws.rc[owner];
    const uint32_t kz = __float_as_uint(rc.y);
    // ray-relative vertices
    const float a0 = (tr.a.x - ra.x) - rb.x, a1 = (tr.a.y - ra.y) - rb.y, a2 = (tr.a.z - ra.z) - rb.z;
    const float b0 = (tr.a.w - ra.x) - rb.x, b1 = (tr.b.x - ra.y) - rb.y, b2 = (tr.b.y - ra.z) - rb.z;
    const float c0 = (tr.b.z - ra.x) - rb.x, c1 = (tr.b.w - ra.y) - rb.y, c2 = (tr.c.x - ra.z) - rb.z;
    // permuted: (kx, ky, kz) = kz + 1, kz + 2, kz
    const float akx = kz == 0u ? a1 : (kz == 1u ? a2 : a0), aky = kz == 0u ? a2 : (kz == 1u ? a0 : a1), akz = kz == 0u ? a0 : (kz == 1u ? a1 : a2);
    const float bkx = kz == 0u ? b1 : (kz == 1u ? b2 : b0), bky = kz == 0u ? b2 : (kz == 1u ? b0 : b1), bkz = kz == 0u ? b0 : (kz == 1u ? b1 : b2);
    const float ckx = kz == 0u ? c1 : (kz == 1u ? c2 : c0), cky = kz == 0u ? c2 : (kz == 1u ? c0 : c1), ckz = kz == 0u ? c0 : (kz == 1u ? c1 : c2);
    const float Sx = ra.w, Sy = rb.w, Sz = rc.x;
    const float Ax = akx - Sx * akz, Ay = aky - Sy * akz;
    const float Bx = bkx - Sx * bkz, By = bky - Sy * bkz;
    const float Cx = ckx - Sx * ckz, Cy = cky - Sy * ckz;
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {  // on an edge to f32 precision: the products are exact in f64
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float T = U * (Sz * akz) + V * (Sz * bkz) + W * (Sz * ckz);
    t = T / det;
    return true;
}

// Closest-hit round of the F32 mode: as prim_round_closest, with the closest distance kept as an f32 bit pattern.  Equal
// distances: the lowest lane of a round wins, a later round's equal distance replaces it (no reference order in this mode).
__device__ __forceinline__ void prim_round_closest32(const SceneView& s, WarpShared32& ws, unsigned lane, uint32_t head, uint32_t count) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool act = lane < count;
    const uint32_t e = act ? ws.queue[(head + lane) & (kQueue - 1u)] : 0u;
    const uint32_t owner = e >> kSlotBits, slot = e & kSlotMask;
    bool hit = false;
    float t32 = 0.0f;
    double t64 = 0.0;
    if (act) {
        const Tri32Regs tr = load_tri32(s.wide_tris32 + slot);
        if ((__float_as_uint(tr.c.z) & 0xFFu) == PRIM_TRIANGLE) {
            float t;
            if (slot != __float_as_uint(ws.rc[owner].z) && tri32_hit(tr, ws, owner, t) && t > ws.tmin[owner] && t < closest32(ws, owner)) {
                hit = true; t32 = t; t64 = (double)t;
            }
        } else {  // spheres and disks: the f64 test of the parity mode on the f64 ray
            const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
            double cand = (double)closest32(ws, owner);
            if (analytic_candidate(s, lp, mk(ws.ox[owner], ws.oy[owner], ws.oz[owner]), mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]), cand, false) == 1) {
                hit = true; t64 = cand; t32 = __double2float_rd(cand);
            }
        }
    }
    if (hit) atomicMin(reinterpret_cast<unsigned*>(&closest32(ws, owner)), __float_as_uint(t32));
    __syncwarp();
    const unsigned grp = __match_any_sync(FULL, act ? owner : 32u + lane);
    const bool win = hit && closest32(ws, owner) == t32;
    const unsigned wins = __ballot_sync(FULL, win) & grp;
    if (win && lane == (unsigned)(__ffs(wins) - 1)) {
        ws.best[owner] = slot;
        ws.tmax[owner] = t64;
        ws.tmax32[owner] = slab_tmax(t64);
    }
    if (act && lane == (unsigned)(__ffs(grp) - 1)) ws.pend[owner] -= __popc(grp);
    __syncwarp();
}

// Any-hit round of the F32 mode.  A shadow ray ends 1e-9 short of its light sample (light.rs:124), and that sample lies ON a
// surface -- the light's own triangle for a mesh light -- whose f32 distance is uncertain by a few ulps divided by the cosine of
// the angle of incidence.  An f32 hit in the last 2^-10 of the ray is therefore confirmed by the f64 test of the parity mode
// against the exact max distance (out of line; a fraction of a percent of the tests), everything nearer is taken as it is.
__device__ __forceinline__ void prim_round_any32(const SceneView& s, WarpShared32& ws, unsigned lane, uint32_t head, uint32_t count) {
    const unsigned FULL = 0xFFFFFFFFu;
    const bool act = lane < count;
    const uint32_t e = act ? ws.queue[(head + lane) & (kQueue - 1u)] : 0u;
    const uint32_t owner = e >> kSlotBits, slot = e & kSlotMask;
    if (act && ws.best[owner] == 0u) {
        const Tri32Regs tr = load_tri32(s.wide_tris32 + slot);
        bool occ;
        if ((__float_as_uint(tr.c.z) & 0xFFu) == PRIM_TRIANGLE) {
            float t;
            occ = slot != __float_as_uint(ws.rc[owner].z) && tri32_hit(tr, ws, owner, t) && t > ws.tmin[owner] && t < closest32(ws, owner);
            if (occ && t > closest32(ws, owner) * 0.9990234375f)   // 1 - 2^-10 (never true for an infinite ray)
                occ = triangle_any_f64(s, slot, mk(ws.ox[owner], ws.oy[owner], ws.oz[owner]), mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]), ws.tmax[owner]);
        } else {
            const LeafPrim lp = load_leaf_prim(s.wide_prims + slot);
            occ = analytic_any(s, lp, mk(ws.ox[owner], ws.oy[owner], ws.oz[owner]), mk(ws.dx[owner], ws.dy[owner], ws.dz[owner]), ws.tmax[owner]);
        }
        if (occ) ws.best[owner] = 1u;
    }
    const unsigned grp = __match_any_sync(FULL, act ? owner : 32u + lane);
    if (act && lane == (unsigned)(__ffs(grp) - 1)) ws.pend[owner] -= __popc(grp);
    __syncwarp();
}

}  // namespace cray
