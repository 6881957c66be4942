// Kernels and entry points of the hot path.
//
//   S3  cray_trace_closest / cray_trace_any     one thread per ray, exact or wide traversal
//   S2  cray_estimate_li                         the wavefront below on an explicit (x, y, sample) list
//   S1  cray_render                              the wavefront below on every (pixel, sample) of a sample range
//
// Wavefront pipeline (replaces render / render_tile / render_pixel / estimate_Li, src/bin/craytracer.rs:148-291 and
// src/path_integrator.rs:41-215): a pool of path slots lives in HBM as structure-of-arrays; every iteration runs
//   k_generate  flush finished paths into the film, refill their slots with new camera rays (sampler + camera),
//               append every live slot to the extend queue
//   extend      closest-hit traversal of the queued rays (k_wide_persistent<false> / k_extend_exact)
//   k_shade     one path vertex: emission, light sample + shadow ray, BSDF sample, Russian roulette; appends the
//               slots whose light sample can contribute to the shadow queue
//   shadow      any-hit traversal of the queued shadow rays, adds the unoccluded contributions
// so all lanes keep working until the sample range is exhausted (path regeneration).
//
// Fast-mode traversal runs as a persistent kernel: every warp pulls rays from the queue with one atomic per refill and
// steps its 32 traversals in lock-step.  Node phases test the 8 quantised child boxes of one interior node per lane (f32)
// and push the primitives whose box was hit onto a per-warp queue in shared memory; as soon as 32 tests are queued the
// whole warp runs them, one (ray, primitive) pair per lane, in f64 -- so the expensive exact tests always execute with
// full SIMD width and 32 independent loads in flight.  Idle lanes are refilled once enough of them have finished, so no
// lane is parked behind the longest ray of a fixed assignment.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "scene_device.hpp"
#include "shading.cuh"
#include "traverse.cuh"

namespace cray {

enum : uint32_t { SLOT_EMPTY = 0, SLOT_ACTIVE = 1, SLOT_DONE = 2 };

struct Pool {  // structure-of-arrays over `capacity` path slots
    uint32_t capacity;
    double *ox, *oy, *oz, *dx, *dy, *dz;          // ray to extend (max_distance is always +inf: Ray::new)
    double* hit_t;                                 // result of the extend stage: distance and leaf slot (CRAY_NO_HIT = miss);
    uint32_t* hit_slot;                            // k_shade re-derives the barycentrics from them (same code, same bits)
    double *beta_r, *beta_g, *beta_b, *L_r, *L_g, *L_b, *prev_bsdf_pdf;
    double *sdx, *sdy, *sdz, *smax, *sc_r, *sc_g, *sc_b;  // pending shadow ray (origin = ox,oy,oz) and its contribution
    uint32_t *id, *pixel, *hash, *shuffled_rev;    // job-relative sample id, film offset, sampler state
    uint32_t* state;                               // state | bounces << 8 | specular << 16 | shadow_pending << 17 | bad << 18 | contact << 19 | bad L << 20
    // slot indices with a ray to extend / a shadow ray to test this iteration.  Rays for the wide traversal fill a queue from the
    // front; rays that start in a contact shell (state bit 19, fast and F32 modes) go to the reference-order traversal and fill the
    // same array from the back (entry k at capacity - 1 - k): the two never meet, there is at most one entry per slot.
    uint32_t *extend_queue, *shadow_queue;
    // the previous iteration's extend queue (the two arrays swap every iteration): once every sample has been generated, the next
    // queue is the survivors of this one (k_requeue) and nobody scans the whole pool any more
    uint32_t* extend_queue_prev;
    // the shade stage's queues: kKeyNone arrays of `capacity` entries, one per (shade class, shape kind) key and one for the misses
    uint32_t* class_queue;
};
constexpr uint32_t kStateContact = 1u << 19;
constexpr uint32_t kStateBadL = 1u << 20;   // the shadow kernel made L non-finite: the next vertex ends the path (path_integrator.rs:208)

struct Job {
    uint64_t seed;
    uint64_t n_total;           // samples in this job
    uint32_t sample_begin;
    uint32_t n_pixels;
    // full frame: the samples of the job are taken `sample_group` at a time -- ids run over the samples of a group fastest, then
    // over the pixels (in pixel_order), then over the groups; the last group may be shorter
    const uint32_t* pixel_order;
    uint32_t n_samples, sample_group;
    const uint32_t *lx, *ly, *ls; // explicit list (S2), or null
    double* film;               // W*H*3 f64 sums, or null
    double* out_rgb;            // per-sample radiance (S2), or null
    const uint32_t* sobol;
    int exact;                  // traversal mode
    int f32;                    // F32 mode: hit_t is an f32-precision distance, the hit is re-evaluated in f64 when shaded
};

struct Counters {
    unsigned long long next_id;
    unsigned long long closest_rays, shadow_rays, nan_samples;
    unsigned long long shadow_traced;               // shadow rays that were actually queued and traced (sum of the queue lengths)
    unsigned long long contact_rays;                // rays (closest + shadow) routed to the reference-order traversal in fast mode
    // per-iteration part, rolled into the totals above and cleared by k_begin_iteration
    unsigned long long n_extend, n_shadow;          // queue lengths (front part)
    unsigned long long extend_cursor, shadow_cursor;  // persistent-kernel fetch positions
    unsigned long long n_extend_contact, n_shadow_contact;  // queue lengths (back part)
    unsigned long long class_count[16];             // lengths of the shade stage's queues (k_shade_classify)
};

__device__ __forceinline__ uint32_t st_state(uint32_t s) { return s & 0xFFu; }
__device__ __forceinline__ uint32_t st_bounces(uint32_t s) { return (s >> 8) & 0xFFu; }

// A finished path goes into the film where it ends (craytracer.rs:177-188 accumulates the sample into its pixel) -- in the shade
// kernel that ended it or, if its last light sample is still being tested, in the shadow kernel -- and frees its slot.  A sample on
// which the reference would have panicked (path_integrator.rs:208-209 and the asserts of its callees) is dropped and counted.
__device__ __forceinline__ void flush_path(const Pool& p, const Job& job, Counters* counters, uint32_t i, double r, double g, double b, bool bad) {
    if (bad || !(isfinite(r) && isfinite(g) && isfinite(b))) {
        atomicAdd(&counters->nan_samples, 1ull);
    } else if (job.film) {
        double* px = job.film + 3ull * p.pixel[i];
        atomicAdd(px, r); atomicAdd(px + 1, g); atomicAdd(px + 2, b);
    }
    if (job.out_rgb) {
        double* dst = job.out_rgb + 3ull * p.id[i];
        dst[0] = r; dst[1] = g; dst[2] = b;
    }
    p.state[i] = SLOT_EMPTY;
}

// ---- ray sources for the persistent wide-BVH kernel -------------------------------------------------------------

struct ExtendSource {  // queued path rays -> Pool::hit_*
    Pool p;
    __device__ __forceinline__ uint32_t load(uint64_t idx, V3& o, V3& d, double& ray_max) const {
        const uint32_t i = p.extend_queue[idx];
        o = mk(p.ox[i], p.oy[i], p.oz[i]);
        d = mk(p.dx[i], p.dy[i], p.dz[i]);
        ray_max = inf_f64();
        return i;
    }
    // F32 mode: the leaf slot the ray leaves -- the hit of the path's previous vertex, CRAY_NO_HIT for a camera ray (k_generate)
    __device__ __forceinline__ uint32_t self_slot(uint32_t i) const { return p.hit_slot[i]; }
    __device__ __forceinline__ float t_min32(V3, V3) const { return kEpsilon32; }
    __device__ __forceinline__ void store_closest(const SceneView&, uint32_t i, uint32_t slot, double t) const {
        p.hit_slot[i] = slot;
        p.hit_t[i] = t;
    }
    __device__ __forceinline__ void store_any(uint32_t, bool) const {}
};

// The result of a path's shadow ray: the parked contribution is added when unoccluded; a path that had already ended (its state
// waited for this) is flushed.
__device__ __forceinline__ void settle_shadow(const Pool& p, const Job& job, Counters* counters, uint32_t i, bool occluded) {
    const uint32_t st = p.state[i];
    if (st_state(st) == SLOT_DONE) {
        double r = p.L_r[i], g = p.L_g[i], b = p.L_b[i];
        if (!occluded) { r += p.sc_r[i]; g += p.sc_g[i]; b += p.sc_b[i]; }
        flush_path(p, job, counters, i, r, g, b, (st >> 18) & 1u);
    } else if (!occluded) {
        const double r = p.L_r[i] + p.sc_r[i], g = p.L_g[i] + p.sc_g[i], b = p.L_b[i] + p.sc_b[i];
        p.L_r[i] = r; p.L_g[i] = g; p.L_b[i] = b;
        // (only the matte class reads L at the next vertex; the others learn of a non-finite L from the state word)
        if (!(isfinite(r) && isfinite(g) && isfinite(b))) p.state[i] = st | kStateBadL;
    }
}

struct ShadowSource {  // queued shadow rays -> L += contribution when unoccluded (path_integrator.rs:141-163)
    Pool p;
    Job job;
    Counters* counters;
    __device__ __forceinline__ uint32_t load(uint64_t idx, V3& o, V3& d, double& ray_max) const {
        const uint32_t i = p.shadow_queue[idx];
        o = mk(p.ox[i], p.oy[i], p.oz[i]);
        d = mk(p.sdx[i], p.sdy[i], p.sdz[i]);
        ray_max = p.smax[i];
        return i;
    }
    __device__ __forceinline__ void store_closest(const SceneView&, uint32_t, uint32_t, double) const {}
    __device__ __forceinline__ void store_any(uint32_t i, bool occluded) const { settle_shadow(p, job, counters, i, occluded); }
};

// `h.u, h.v` are only consulted when `have_uv`; otherwise a triangle's barycentrics are re-derived from (slot, t).
// `f32`: F32 mode, h.t has f32 precision -- the triangle found is evaluated again in f64 (t, u, v of its plane if the f64 test
// itself rejects a grazing hit).
__device__ __forceinline__ void write_hit(const SceneView& s, const LeafPrim* prims, const cray_ray& r, Hit h, bool have_uv, bool found, uint64_t i,
                                          cray_hit* hits, cray_surface* surf, bool f32 = false) {
    cray_hit out;
    out._pad = 0;
    cray_surface sf;
    memset(&sf, 0, sizeof(sf));
    if (found) {
        const V3 o = mk(r.origin[0], r.origin[1], r.origin[2]), d = mk(r.direction[0], r.direction[1], r.direction[2]);
        const LeafPrim lp = load_leaf_prim(prims + h.slot);
        const bool tri = (lp.kind & 0xFFu) == PRIM_TRIANGLE;
        if (tri && !have_uv) {
            double t2 = h.t;
            if (!triangle_eval(lp.d, o, d, t2, h.u, h.v) && f32) triangle_eval_unchecked(lp.d, o, d, t2, h.u, h.v);
            if (f32) h.t = t2;
        }
        V3 loc, nrm;
        double tu, tv;
        surface_at(s, lp, o, d, h.t, h.u, h.v, loc, nrm, tu, tv);
        out.prim = lp.prim;
        out.t = h.t;
        out.u = tri ? h.u : tu;  // triangles: Moeller-Trumbore barycentrics; spheres / disks: surface uv
        out.v = tri ? h.v : tv;
        sf.location[0] = loc.x; sf.location[1] = loc.y; sf.location[2] = loc.z;
        sf.normal[0] = nrm.x; sf.normal[1] = nrm.y; sf.normal[2] = nrm.z;
        sf.uv[0] = tu; sf.uv[1] = tv;
    } else {
        out.prim = CRAY_NO_HIT;
        out.t = 0.0; out.u = 0.0; out.v = 0.0;
    }
    hits[i] = out;
    if (surf) surf[i] = sf;
}

struct RayArraySource {  // S3: caller-provided cray_ray records
    const cray_ray* rays;
    cray_hit* hits;
    cray_surface* surf;
    uint8_t* occluded;
    bool f32;
    const uint32_t* index;  // the rays of this launch (k_classify_rays), or null for all of them
    __device__ __forceinline__ uint32_t self_slot(uint32_t) const { return CRAY_NO_HIT; }
    __device__ __forceinline__ float t_min32(V3 o, V3 d) const { return f32_unknown_origin_tmin(o, d); }
    __device__ __forceinline__ uint32_t load(uint64_t idx, V3& o, V3& d, double& ray_max) const {
        const uint32_t i = index ? index[idx] : (uint32_t)idx;
        const cray_ray r = rays[i];
        o = mk(r.origin[0], r.origin[1], r.origin[2]);
        d = mk(r.direction[0], r.direction[1], r.direction[2]);
        ray_max = r.max_distance;
        return i;
    }
    __device__ __forceinline__ void store_closest(const SceneView& s, uint32_t i, uint32_t slot, double t) const {
        Hit h;
        h.slot = slot; h.t = t; h.u = 0.0; h.v = 0.0;
        write_hit(s, s.wide_prims, rays[i], h, false, slot != CRAY_NO_HIT, i, hits, surf, f32);
    }
    __device__ __forceinline__ void store_any(uint32_t i, bool occ) const { occluded[i] = occ ? 1 : 0; }
};

// ---- persistent wide-BVH traversal ----------------------------------------------------------------------------------
//
// One warp = 32 concurrent traversals.  Per loop iteration:
//   refill  if at least tune.refill_lanes lanes are idle (or all are), one atomicAdd claims that many queue entries
//   node    every lane with a pending interior child visits it (node_step: 8 quantised box tests in f32, hit primitives
//           are pushed onto the warp's queue)
//   prims   while 32 tests are queued -- or fewer, when no lane has node work left or tune.wait_lanes lanes can do nothing
//           but wait for their tests -- the warp runs them, one per lane (prim_round_*)
//   finish  a lane whose node stack is empty and whose tests have all run writes its result and becomes idle
struct WideTuning {
    int refill_lanes;  // refill once this many lanes are idle
    int wait_lanes;    // run a partial primitive round once this many lanes are waiting on queued tests
};

#ifndef CRAY_WIDE_MIN_BLOCKS
#define CRAY_WIDE_MIN_BLOCKS 5
#endif
constexpr int kWideMinBlocks = CRAY_WIDE_MIN_BLOCKS;
#ifndef CRAY_WIDE_MIN_BLOCKS_ANY
#define CRAY_WIDE_MIN_BLOCKS_ANY 9
#endif
constexpr int kWideMinBlocksAny = CRAY_WIDE_MIN_BLOCKS_ANY;   // the any-hit instantiation carries less state   // CTAs of 128 threads per SM the register allocation is held to
#ifndef CRAY_WIDE_MIN_BLOCKS_F32
#define CRAY_WIDE_MIN_BLOCKS_F32 8
#endif
constexpr int kWideMinBlocksF32 = CRAY_WIDE_MIN_BLOCKS_F32;   // F32 mode: no f64 triangle test to hold registers for

enum : int { LANE_IDLE = 0, LANE_LIVE = 1, LANE_DRAIN = 2 };
#ifndef CRAY_NODE_REPS
#define CRAY_NODE_REPS 6
#endif
constexpr int kNodeReps = CRAY_NODE_REPS;   // node phases per iteration of the persistent loop

// Tuning build (make VARIANT=stats EXTRA=-DCRAY_WIDE_STATS=1): work counters of the traversal kernels, read with
// cray_debug_wide_stats.  [0] rays  [1] warp iterations  [2] node steps (lanes)  [3] node phases (warps)
// [4] primitive tests (lanes)  [5] primitive rounds (warps)  [6] refills (warps)  [7] idle-lane iterations
// [8] node steps that hit nothing  [9] interior children hit  [10] node steps that queued no primitive  [11] node steps taken
// while the lane's ray already had a hit (closest) 
#ifndef CRAY_WIDE_STATS
#define CRAY_WIDE_STATS 0
#endif
#if CRAY_WIDE_STATS
__device__ unsigned long long g_wide_stats[2][12];
#define WIDE_STAT(k, v) stat[k] += (unsigned long long)(v)
#else
#define WIDE_STAT(k, v)
#endif

template <bool ANY, class Source, bool F32 = false>
__global__ void __launch_bounds__(128, F32 ? kWideMinBlocksF32 : (ANY ? kWideMinBlocksAny : kWideMinBlocks)) k_wide_persistent(const __grid_constant__ SceneView s, const __grid_constant__ Source src, const unsigned long long* __restrict__ n_ptr, unsigned long long* cursor, WideTuning tune) {
    using WS = std::conditional_t<F32, WarpShared32, WarpShared>;
    __shared__ WS shared[4];
    WS& ws = shared[threadIdx.x >> 5];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long n = *n_ptr;
    if (lane == 0) ws.tail = 0u;
    ws.pend[lane] = 0u;
    __syncwarp();
    uint32_t head = 0;  // queue entries consumed so far (warp-uniform)
#if CRAY_WIDE_STATS
    unsigned long long stat[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    WideRay r;
    uint2 ng = make_uint2(0u, 0u);
    uint2 stack[kWideStack];
    int sp = 0;
    int state = LANE_IDLE;
    uint32_t id = 0;
    bool exhausted = false;
    for (;;) {
        const unsigned idle = __ballot_sync(FULL, state == LANE_IDLE);
        WIDE_STAT(1, 1);
        WIDE_STAT(7, __popc(idle));  // counters are warp-uniform; lane 0 publishes them
        if (!exhausted && (idle == FULL || __popc(idle) >= tune.refill_lanes)) {
            const int want = __popc(idle);
            WIDE_STAT(6, 1);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(cursor, (unsigned long long)want);
            base = __shfl_sync(FULL, base, 0);
            if (state == LANE_IDLE) {
                const unsigned long long idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < n) {
                    V3 o, d;
                    double ray_max;
                    id = src.load(idx, o, d, ray_max);
                    r = make_wide_ray(o, d, ray_max);
                    if constexpr (F32) warp_begin_ray32<ANY>(ws, lane, o, d, ray_max, ANY ? 0u : CRAY_NO_HIT, src.self_slot(id), src.t_min32(o, d));
                    else warp_begin_ray(ws, lane, o, d, ray_max, ANY ? 0u : CRAY_NO_HIT);
                    ng = make_uint2(0u, 0x80000000u);  // the root, as a one-child node group
                    sp = 0;
                    state = LANE_LIVE;
                }
            }
            WIDE_STAT(0, __popc(__ballot_sync(FULL, state == LANE_LIVE) & idle));
            if (base + want >= n) exhausted = true;
        }
        if (__ballot_sync(FULL, state != LANE_IDLE) == 0u) {
            if (exhausted) break;
            continue;
        }
        // node phase(s)
#pragma unroll 1
        for (int rep = 0;;) {
        bool stepping = false;
        if (state == LANE_LIVE) {
            if (!(ng.y & 0xFF000000u) && sp > 0) ng = stack[--sp];
            stepping = (ng.y & 0xFF000000u) != 0u;
            uint32_t outcome = 0xFFFFFFFFu;
#if CRAY_NODE_PAIR
            if (stepping) outcome = node_step_pair(s, ws, lane, r, ng, stack, sp);
#else
            if (stepping) outcome = node_step(s, ws, lane, r, ng, stack, sp);
#endif
#if CRAY_WIDE_STATS
            {
                // outcome: interior hits | leaf hits << 8 of node A; paired steps add bit 16 = a node B was visited, bit 17 = B queued
                // primitives, bits 24.. = B's interior hits
                const unsigned m = __activemask();
                const bool second = CRAY_NODE_PAIR && stepping && (outcome & 0x10000u);
                const bool a_empty = stepping && (outcome & 0xFFFFu) == 0u, b_empty = second && (outcome >> 24) == 0u && !(outcome & 0x20000u);
                WIDE_STAT(8, __popc(__ballot_sync(m, a_empty)) + __popc(__ballot_sync(m, b_empty)));
                WIDE_STAT(10, __popc(__ballot_sync(m, stepping && ((outcome >> 8) & 0xFFu) == 0u)) + __popc(__ballot_sync(m, second && !(outcome & 0x20000u))));
                WIDE_STAT(11, __popc(__ballot_sync(m, stepping && ws.best[lane] != (ANY ? 0u : CRAY_NO_HIT))) * (1 + 0) + __popc(__ballot_sync(m, second && ws.best[lane] != (ANY ? 0u : CRAY_NO_HIT))));
                unsigned ihs = stepping ? __popc(outcome & 0xFFu) + (second ? __popc(outcome >> 24) : 0u) : 0u;
                for (int o = 16; o; o >>= 1) ihs += __shfl_xor_sync(m, ihs, o);
                WIDE_STAT(9, ihs);
                WIDE_STAT(2, __popc(__ballot_sync(m, second)));   // (the first node of each step is counted below)
            }
#else
            (void)outcome;
#endif
            // a node group with nothing left to visit is replaced from the stack right away: the (local-memory) load has the
            // rest of the iteration to arrive instead of stalling the next node step
            if (!(ng.y & 0xFF000000u) && sp > 0) ng = stack[--sp];
        }
#if CRAY_WIDE_STATS
        {
            const unsigned m = __ballot_sync(FULL, stepping);
            WIDE_STAT(2, __popc(m));
            WIDE_STAT(3, m != 0u);
        }
#endif
        __syncwarp();
        // CRAY_NODE_REPS > 1: further node phases before the warp-wide bookkeeping below (its ballots, the finish test, the refill
        // test cost about a third of a node step), as long as no full round of tests is queued -- which also keeps the queue within
        // its 31 + 32 x 8 entries -- and some lane still has a node to visit
        if (++rep >= kNodeReps) break;
        if (*(volatile uint32_t*)&ws.tail - head >= 32u) break;
        if (!__any_sync(FULL, state == LANE_LIVE && ((ng.y & 0xFF000000u) || sp > 0))) break;
        }
        // primitive rounds
        uint32_t count = *(volatile uint32_t*)&ws.tail - head;
        if (count) {
            const bool node_work = state == LANE_LIVE && ((ng.y & 0xFF000000u) || sp > 0);
            const unsigned node_lanes = __ballot_sync(FULL, node_work);
            const unsigned waiting = __ballot_sync(FULL, state != LANE_IDLE && !node_work);
            while (count >= 32u || (count > 0u && (node_lanes == 0u || __popc(waiting) >= tune.wait_lanes))) {
                const uint32_t take = count < 32u ? count : 32u;
                if constexpr (F32 && ANY) prim_round_any32(s, ws, lane, head, take);
                else if constexpr (F32) prim_round_closest32(s, ws, lane, head, take);
                else if constexpr (ANY) prim_round_any(s, ws, lane, head, take);
                else prim_round_closest(s, ws, lane, head, take);
                head += take;
                count -= take;
                WIDE_STAT(4, take);
                WIDE_STAT(5, 1);
            }
            r.tmax = ws.tmax32[lane];
        }
        // finish
        if (state != LANE_IDLE) {
            const uint32_t pend = ws.pend[lane];
            if constexpr (ANY) {
                if (state == LANE_LIVE) {
                    if (ws.best[lane]) { src.store_any(id, true); state = LANE_DRAIN; }
                    else if (!(ng.y & 0xFF000000u) && sp == 0 && pend == 0u) { src.store_any(id, false); state = LANE_IDLE; }
                }
                if (state == LANE_DRAIN && pend == 0u) state = LANE_IDLE;  // its stale tests have left the queue
            } else {
                if (!(ng.y & 0xFF000000u) && sp == 0 && pend == 0u) {
                    src.store_closest(s, id, ws.best[lane], ws.tmax[lane]);
                    state = LANE_IDLE;
                }
            }
        }
    }
#if CRAY_WIDE_STATS
    if (lane == 0)
        for (int k = 0; k < 12; ++k) atomicAdd(&g_wide_stats[ANY ? 1 : 0][k], stat[k]);
#endif
}

// ---- reference-order kernels (the reference's binary BVH, one thread per ray, grid-stride) ---------------------------------
//
// They serve the exact mode (every ray) and, inside the fast mode, the rays that start in a contact shell (bvh_build.hpp "planar
// contact"): only this traversal reproduces the reference's box-test culls for them.  `index_back`: the rays of the launch are
// listed from the back of an array of `n_all` entries (k_classify_rays); null: rays [0, n).
constexpr unsigned kExactBlocks = 148 * 8;

__global__ void __launch_bounds__(128) k_trace_closest_exact(SceneView s, const cray_ray* __restrict__ rays, const unsigned long long* __restrict__ n_ptr, const uint32_t* __restrict__ index_back,
                                                              uint64_t n_all, cray_hit* __restrict__ hits, cray_surface* __restrict__ surf) {
    const uint64_t n = *n_ptr;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = index_back ? index_back[n_all - 1 - q] : q;
        const cray_ray r = rays[i];
        Hit h;
        const bool found = traverse_exact<false>(s, mk(r.origin[0], r.origin[1], r.origin[2]), mk(r.direction[0], r.direction[1], r.direction[2]), r.max_distance, h);
        write_hit(s, s.bin_prims, r, h, true, found, i, hits, surf);
    }
}

__global__ void __launch_bounds__(128) k_trace_any_exact(SceneView s, const cray_ray* __restrict__ rays, const unsigned long long* __restrict__ n_ptr, const uint32_t* __restrict__ index_back,
                                                          uint64_t n_all, uint8_t* __restrict__ occluded) {
    const uint64_t n = *n_ptr;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = index_back ? index_back[n_all - 1 - q] : q;
        const cray_ray r = rays[i];
        Hit h;
        occluded[i] = traverse_exact<true>(s, mk(r.origin[0], r.origin[1], r.origin[2]), mk(r.direction[0], r.direction[1], r.direction[2]), r.max_distance, h) ? 1 : 0;
    }
}

// S3, fast mode, scenes with marked boxes: splits the batch into the rays the wide traversal may take (listed from the front of
// `list`) and the rays that start in a contact shell (from the back).  counts[0] = front length, counts[2] = back length.
__global__ void __launch_bounds__(128) k_classify_rays(SceneView s, const cray_ray* __restrict__ rays, uint64_t n, uint32_t* __restrict__ list, unsigned long long* counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cray_ray r = rays[i];
    const bool contact = origin_in_contact_shell(s, mk(r.origin[0], r.origin[1], r.origin[2]), contact_tol(mk(r.direction[0], r.direction[1], r.direction[2])));
    const unsigned m = __activemask();
    const unsigned mc = __ballot_sync(m, contact), mw = m & ~mc;
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long bw = 0, bc = 0;
    if (mw && lane == (unsigned)(__ffs(mw) - 1)) bw = atomicAdd(&counts[0], (unsigned long long)__popc(mw));
    if (mc && lane == (unsigned)(__ffs(mc) - 1)) bc = atomicAdd(&counts[2], (unsigned long long)__popc(mc));
    if (mw) bw = __shfl_sync(m, bw, __ffs(mw) - 1);
    if (mc) bc = __shfl_sync(m, bc, __ffs(mc) - 1);
    if (contact) list[n - 1 - (bc + __popc(mc & ((1u << lane) - 1u)))] = (uint32_t)i;
    else list[bw + __popc(mw & ((1u << lane) - 1u))] = (uint32_t)i;
}

// ---- wavefront kernels --------------------------------------------------------------------------------------

__device__ __forceinline__ void warp_count(unsigned long long* counter, bool pred) {
    const unsigned mask = __ballot_sync(__activemask(), pred);
    if (pred && (threadIdx.x & 31) == (unsigned)(__ffs(mask) - 1)) atomicAdd(counter, (unsigned long long)__popc(mask));
}

// Warp-aggregated append of `value` to a device queue (one atomic per converged group of lanes).
__device__ __forceinline__ void queue_append(unsigned long long* counter, uint32_t* queue, uint32_t value) {
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31u, leader = __ffs(m) - 1u;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// ... and to the back part of a queue of `capacity` entries (entry k sits at capacity - 1 - k)
__device__ __forceinline__ void queue_append_back(unsigned long long* counter, uint32_t* queue, uint32_t capacity, uint32_t value) {
    const unsigned m = __activemask();
    const unsigned lane = threadIdx.x & 31u, leader = __ffs(m) - 1u;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    queue[capacity - 1u - (uint32_t)(base + __popc(m & ((1u << lane) - 1u)))] = value;
}

// Camera::sample + generate_ray camera.rs:131-162
__device__ __forceinline__ void camera_ray(const DevCamera& c, double fu, double fv, double lu, double lv, uint32_t x, uint32_t y, V3& o, V3& d) {
    const double dx = 2.0 * fu - 1.0, dy = 2.0 * fv - 1.0;
    const V3 pr = mk((double)x + dx, (double)y + dy, 0.0);
    const double(*m)[4] = c.camera_from_raster;
    V3 pc = mk(m[0][0] * pr.x + m[0][1] * pr.y + m[0][2] * pr.z + m[0][3], m[1][0] * pr.x + m[1][1] * pr.y + m[1][2] * pr.z + m[1][3],
               m[2][0] * pr.x + m[2][1] * pr.y + m[2][2] * pr.z + m[2][3]);
    pc = pc / (m[3][0] * pr.x + m[3][1] * pr.y + m[3][2] * pr.z + m[3][3]);
    V3 ro = pc, rd = c.perspective ? normalized(pc - mk(0.0, 0.0, 0.0)) : mk(0.0, 0.0, 1.0);
    if (c.lens_radius != 0.0) {
        const double lens_x = 2.0 * lu - 1.0, lens_y = 2.0 * lv - 1.0;
        const V3 p_lens = mk(lens_x * c.lens_radius, lens_y * c.lens_radius, 0.0);
        const V3 p_focal = ro + rd * (c.focal_distance / rd.z);
        ro = p_lens;
        rd = normalized(p_focal - p_lens);
    }
    o = xf_point(c.world_from_camera, ro);
    d = xf_vector(c.world_from_camera, rd);
}

// Block-wide exclusive prefix sum of `value` over the 256 threads (ascending thread order) and its total.  `s_warp` = 8 words of
// shared memory; ends with a barrier, so it can be called again right away.
__device__ __forceinline__ uint32_t block_scan_256(uint32_t value, uint32_t* s_warp, uint32_t& total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = value;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (unsigned)d) incl += up;
    }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (unsigned w = 0; w < 8; ++w) {
        const uint32_t c = s_warp[w];
        before += w < warp ? c : 0u;
        all += c;
    }
    __syncthreads();
    total = all;
    return before + incl - value;
}

// One block = 2048 consecutive path slots, eight per thread (2, 4 and 8 measured: generate 6.7 / 4.8 / 4.2 ms per 256-spp frame) (the scan of the slot states is what this kernel does most: a block
// per 256 slots spent its time in barriers).  The slots to refill are compacted, so that the camera-ray code (SipHash, four Sobol
// values, f64 camera transform) runs on full warps instead of on the scattered lanes whose path happened to end -- and the sample
// ids and the extend-queue range of the whole block are each claimed with ONE atomic.  (Finished paths were flushed where they
// ended, flush_path; a slot still found DONE here is flushed as a fallback.)
#ifndef CRAY_GEN_SLOTS
#define CRAY_GEN_SLOTS 8
#endif
constexpr uint32_t kGenSlots = CRAY_GEN_SLOTS, kGenBlock = 256u * kGenSlots;
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ SceneView s, const __grid_constant__ Pool p, const __grid_constant__ Job job, Counters* counters) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_list[kGenBlock];
    __shared__ unsigned long long s_base[3];
    const uint32_t first = (blockIdx.x * 256u + threadIdx.x) * kGenSlots;
    uint32_t st[kGenSlots];
    uint32_t n_mine = 0;   // empty slots among this thread's
#pragma unroll
    for (uint32_t k = 0; k < kGenSlots; ++k) {
        const uint32_t i = first + k;
        st[k] = i < p.capacity ? p.state[i] : 0xFFu;  // 0xFF: no slot
        if (st_state(st[k]) == SLOT_DONE) {
            flush_path(p, job, counters, i, p.L_r[i], p.L_g[i], p.L_b[i], (st[k] >> 18) & 1u);
            st[k] = SLOT_EMPTY;
        }
        n_mine += st_state(st[k]) == SLOT_EMPTY ? 1u : 0u;
    }
    // the empty slots of this block, compacted in ascending order
    uint32_t n_empty;
    const uint32_t my_rank = block_scan_256(n_mine, s_warp, n_empty);
    {
        uint32_t r = my_rank;
#pragma unroll
        for (uint32_t k = 0; k < kGenSlots; ++k)
            if (st_state(st[k]) == SLOT_EMPTY) s_list[r++] = first + k;
    }
    if (threadIdx.x == 0 && n_empty) s_base[0] = atomicAdd(&counters->next_id, (unsigned long long)n_empty);
    __syncthreads();
    const unsigned long long id_base = n_empty ? s_base[0] : 0ull;
    for (uint32_t j = threadIdx.x; j < n_empty; j += 256u) {
        const unsigned long long id = id_base + j;
        if (id >= job.n_total) break;
        const uint32_t slot = s_list[j];
        uint32_t x, y, si;
        if (job.lx) { x = job.lx[id]; y = job.ly[id]; si = job.ls[id]; }
        else {
            const unsigned long long pass = (unsigned long long)job.n_pixels * job.sample_group;
            const uint32_t group = (uint32_t)(id / pass), rem = (uint32_t)(id % pass);
            const uint32_t in_group = min(job.sample_group, job.n_samples - group * job.sample_group);
            const uint32_t po = job.pixel_order[rem / in_group];
            x = po & 0xFFFFu; y = po >> 16;
            si = job.sample_begin + group * job.sample_group + rem % in_group;
        }
        PixelSampler smp;
        smp.start_pixel(job.seed, x, y, si);
        const double fu = smp.sample_1d(job.sobol), fv = smp.sample_1d(job.sobol);
        const double lu = smp.sample_1d(job.sobol), lv = smp.sample_1d(job.sobol);
        V3 o, d;
        camera_ray(s.camera, fu, fv, lu, lv, x, y, o, d);
        p.ox[slot] = o.x; p.oy[slot] = o.y; p.oz[slot] = o.z;
        p.dx[slot] = d.x; p.dy[slot] = d.y; p.dz[slot] = d.z;
        p.beta_r[slot] = 1.0; p.beta_g[slot] = 1.0; p.beta_b[slot] = 1.0;
        p.L_r[slot] = 0.0; p.L_g[slot] = 0.0; p.L_b[slot] = 0.0;
        p.prev_bsdf_pdf[slot] = 0.0;
        p.hit_slot[slot] = CRAY_NO_HIT;  // a camera ray leaves no surface (F32 mode's self-intersection rule)
        p.id[slot] = (uint32_t)id;
        p.pixel[slot] = x + y * s.camera.width;
        p.hash[slot] = smp.hash;
        p.shuffled_rev[slot] = smp.shuffled_rev;
        p.state[slot] = SLOT_ACTIVE | (1u << 16);  // bounces = 0, is_specular_bounce = true (path_integrator.rs:50)
    }
    // every live slot has a ray to extend this iteration: the block's slots go to the queue in ascending order -- to its back
    // part if the ray starts in a contact shell (the shade stage found out) and needs the reference-order traversal
    uint32_t live_mask = 0, contact_mask = 0;
    {
        uint32_t r = my_rank;
#pragma unroll
        for (uint32_t k = 0; k < kGenSlots; ++k) {
            const bool empty = st_state(st[k]) == SLOT_EMPTY, active = st_state(st[k]) == SLOT_ACTIVE;
            const bool live = active || (empty && id_base + r < job.n_total);
            r += empty ? 1u : 0u;
            const bool contact = active && (st[k] & kStateContact);
            live_mask |= (live && !contact ? 1u : 0u) << k;
            contact_mask |= (contact ? 1u : 0u) << k;
        }
    }
    uint32_t totals;
    const uint32_t ranks = block_scan_256(__popc(live_mask) | (__popc(contact_mask) << 16), s_warp, totals);
    const uint32_t n_wide = totals & 0xFFFFu, n_contact = totals >> 16;
    if (threadIdx.x == 0 && n_wide) s_base[1] = atomicAdd(&counters->n_extend, (unsigned long long)n_wide);
    if (threadIdx.x == 32 && n_contact) s_base[2] = atomicAdd(&counters->n_extend_contact, (unsigned long long)n_contact);
    __syncthreads();
    uint32_t qrank = ranks & 0xFFFFu, crank = ranks >> 16;
#pragma unroll
    for (uint32_t k = 0; k < kGenSlots; ++k) {
        if (contact_mask & (1u << k)) p.extend_queue[p.capacity - 1u - (uint32_t)(s_base[2] + crank++)] = first + k;
        else if (live_mask & (1u << k)) p.extend_queue[s_base[1] + qrank++] = first + k;
    }
}

// k_generate's stand-in once the job has no samples left to start: the iteration's extend queue = the slots of the previous
// one whose path is still alive, in the same order, found by walking that queue (n_front entries at its front, n_contact at its
// back) instead of the whole pool -- by then most slots are empty and stay so.  (A path that ended waiting for its shadow ray
// was flushed by the shadow kernel; one still found DONE is flushed here, as in k_generate.)
__global__ void __launch_bounds__(256) k_requeue(const __grid_constant__ Pool p, const __grid_constant__ Job job, Counters* counters, uint64_t n_front, uint64_t n_contact) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base[2];
    const uint64_t first = ((uint64_t)blockIdx.x * 256u + threadIdx.x) * kGenSlots, n_prev = n_front + n_contact;
    uint32_t slot[kGenSlots];
    uint32_t live_mask = 0, contact_mask = 0;
#pragma unroll
    for (uint32_t k = 0; k < kGenSlots; ++k) {
        const uint64_t q = first + k;
        if (q >= n_prev) break;
        const uint32_t i = p.extend_queue_prev[q < n_front ? (uint32_t)q : p.capacity - 1u - (uint32_t)(q - n_front)];
        const uint32_t st = p.state[i];
        slot[k] = i;
        if (st_state(st) == SLOT_DONE) flush_path(p, job, counters, i, p.L_r[i], p.L_g[i], p.L_b[i], (st >> 18) & 1u);
        else if (st_state(st) == SLOT_ACTIVE) {
            if (st & kStateContact) contact_mask |= 1u << k;
            else live_mask |= 1u << k;
        }
    }
    uint32_t totals;
    const uint32_t ranks = block_scan_256(__popc(live_mask) | (__popc(contact_mask) << 16), s_warp, totals);
    const uint32_t n_wide = totals & 0xFFFFu, n_back = totals >> 16;
    if (threadIdx.x == 0 && n_wide) s_base[0] = atomicAdd(&counters->n_extend, (unsigned long long)n_wide);
    if (threadIdx.x == 32 && n_back) s_base[1] = atomicAdd(&counters->n_extend_contact, (unsigned long long)n_back);
    __syncthreads();
    uint32_t qrank = ranks & 0xFFFFu, crank = ranks >> 16;
#pragma unroll
    for (uint32_t k = 0; k < kGenSlots; ++k) {
        if (contact_mask & (1u << k)) p.extend_queue[p.capacity - 1u - (uint32_t)(s_base[1] + crank++)] = slot[k];
        else if (live_mask & (1u << k)) p.extend_queue[s_base[0] + qrank++] = slot[k];
    }
}

// Start of a wavefront iteration: rolls the queue lengths of the previous one into the run's totals and clears the per-iteration
// counters (one thread).
__global__ void k_begin_iteration(Counters* c) {
    c->shadow_traced += c->n_shadow + c->n_shadow_contact;
    c->contact_rays += c->n_extend_contact + c->n_shadow_contact;
    c->n_extend = 0; c->n_shadow = 0; c->extend_cursor = 0; c->shadow_cursor = 0; c->n_extend_contact = 0; c->n_shadow_contact = 0;
    // every vertex on a surface makes the reference's Scene::intersects call (path_integrator.rs:141): one counted shadow ray each
    for (uint32_t k = 0; k < 12u; ++k) c->shadow_rays += c->class_count[k];
    for (int k = 0; k < 16; ++k) c->class_count[k] = 0;
}

// BACK: the contact rays of the fast mode (back part of the queue); their hits are reported as wide leaf slots, which is what
// k_shade reads in that mode.
template <bool BACK>
__global__ void __launch_bounds__(128) k_extend_exact(SceneView s, Pool p, const unsigned long long* __restrict__ n_ptr) {
    const uint64_t n = *n_ptr;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = p.extend_queue[BACK ? p.capacity - 1u - (uint32_t)q : (uint32_t)q];
        Hit h;
        traverse_exact<false>(s, mk(p.ox[i], p.oy[i], p.oz[i]), mk(p.dx[i], p.dy[i], p.dz[i]), inf_f64(), h);
        if (BACK && h.slot != CRAY_NO_HIT) h.slot = s.wide_slot_of_prim[s.bin_prims[h.slot].prim];
        p.hit_slot[i] = h.slot;
        p.hit_t[i] = h.t;
    }
}

// The shade stage: one iteration of the `while` loop of estimate_Li (path_integrator.rs:54-212) for every path of the iteration.
//
//   k_shade_classify   every block counting-sorts its 256 consecutive queue entries by (shade class, shape kind) of what they hit
//                      and appends each group to that key's queue (one atomic per block and key).  Within a key the entries stay in
//                      runs of ascending slots, so the structure-of-arrays path state is still read and written sector-wise.
//   k_shade_class<C>   one kernel per shade class (matte, glass, plastic, metal) over its three shape queues: each carries only its
//                      own material family's code and registers (a conductor vertex runs no light sampling code at all)
//   k_shade_miss       the paths that left the scene
#ifndef CRAY_SHADE_THREADS
#define CRAY_SHADE_THREADS 256
#endif
constexpr uint32_t kShadeThreads = CRAY_SHADE_THREADS;
constexpr uint32_t kShadeKeys = 16;   // class (matte, glass, plastic, metal) x shape kind (3); 12 = miss; 13 = no path
constexpr uint32_t kKeyMiss = 12u, kKeyNone = 13u;

#ifndef CRAY_CLASSIFY_ITEMS
#define CRAY_CLASSIFY_ITEMS 4
#endif
constexpr uint32_t kClassifyItems = CRAY_CLASSIFY_ITEMS, kClassifyBlock = kShadeThreads * kClassifyItems;   // queue entries per thread / block

__global__ void __launch_bounds__(kShadeThreads) k_shade_classify(const __grid_constant__ SceneView s, const __grid_constant__ Pool p, const __grid_constant__ Job job, Counters* counters) {
    // cells in (key, round, warp) order: within a key the block's entries keep their queue order
    constexpr uint32_t kWarps = kShadeThreads / 32, kRows = kClassifyItems * kWarps, kCells = kShadeKeys * kRows;
    static_assert(kCells == 2 * kShadeThreads, "two cells per thread in the scan");
    __shared__ uint32_t s_cell[kCells];
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_order[kClassifyBlock];
    __shared__ uint8_t s_key[kClassifyBlock];
    __shared__ uint32_t s_start[kShadeKeys + 1];
    __shared__ unsigned long long s_base[kShadeKeys];
    // the iteration's rays: the wide traversal's (front of the queue), then the reference-order traversal's (back of it)
    const uint64_t n_front = counters->n_extend;
    const uint64_t n_extend = n_front + counters->n_extend_contact;
    const uint64_t block_base = (uint64_t)blockIdx.x * kClassifyBlock;
    if (block_base >= n_extend) return;  // whole block idle
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t mine[kClassifyItems], key[kClassifyItems], within[kClassifyItems];
#pragma unroll
    for (uint32_t r = 0; r < kClassifyItems; ++r) {
        const uint64_t q = block_base + r * kShadeThreads + threadIdx.x;
        mine[r] = 0xFFFFFFFFu; key[r] = kKeyNone;
        if (q < n_extend) {
            mine[r] = p.extend_queue[q < n_front ? (uint32_t)q : p.capacity - 1u - (uint32_t)(q - n_front)];
            const uint32_t my_slot = p.hit_slot[mine[r]];
            key[r] = kKeyMiss;
            if (my_slot != CRAY_NO_HIT) {
                const LeafPrim* rec = (job.exact ? s.bin_prims : s.wide_prims) + my_slot;
                const uint32_t kind = __ldg(&rec->kind);
                key[r] = ((kind >> 8) & 3u) * 3u + (kind & 3u);
            }
        }
    }
    s_cell[2u * threadIdx.x] = 0u; s_cell[2u * threadIdx.x + 1u] = 0u;
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < kClassifyItems; ++r) {
        const unsigned grp = __match_any_sync(0xFFFFFFFFu, key[r]);
        if (lane == (unsigned)(__ffs(grp) - 1)) s_cell[key[r] * kRows + r * kWarps + warp] = __popc(grp);
        within[r] = __popc(grp & ((1u << lane) - 1u));
    }
    __syncthreads();
    {   // exclusive scan of the cells: two per thread
        const uint32_t c0 = s_cell[2u * threadIdx.x], c1 = s_cell[2u * threadIdx.x + 1u];
        uint32_t total;
        const uint32_t excl = block_scan_256(c0 + c1, s_warp, total);
        s_cell[2u * threadIdx.x] = excl; s_cell[2u * threadIdx.x + 1u] = excl + c0;
    }
    __syncthreads();
    if (threadIdx.x < kShadeKeys) s_start[threadIdx.x] = s_cell[threadIdx.x * kRows];   // first entry of each key in the sorted block
    if (threadIdx.x == 0) s_start[kShadeKeys] = kClassifyBlock;
#pragma unroll
    for (uint32_t r = 0; r < kClassifyItems; ++r) {
        const uint32_t pos = s_cell[key[r] * kRows + r * kWarps + warp] + within[r];
        s_order[pos] = mine[r];
        s_key[pos] = (uint8_t)key[r];
    }
    __syncthreads();
    // one atomic per key claims the block's range in that key's queue
    if (threadIdx.x < kKeyNone) {
        const uint32_t cnt = s_start[threadIdx.x + 1] - s_start[threadIdx.x];
        s_base[threadIdx.x] = cnt ? atomicAdd(&counters->class_count[threadIdx.x], (unsigned long long)cnt) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t r = 0; r < kClassifyItems; ++r) {
        const uint32_t pos = r * kShadeThreads + threadIdx.x;
        const uint32_t k = s_key[pos];
        if (k < kKeyNone) p.class_queue[(uint64_t)k * p.capacity + s_base[k] + (pos - s_start[k])] = s_order[pos];
    }
}

// One path vertex on a primitive of shade class CLS.
template <uint32_t CLS>
__device__ __forceinline__ void shade_vertex(const SceneView& s, const Pool& p, const Job& job, Counters* counters, uint32_t i, uint32_t slot) {
    constexpr uint32_t LOBES = lobes_of_class(CLS);
    uint32_t st = p.state[i];
    const uint32_t bounces = st_bounces(st);
    const bool is_specular_bounce = (st >> 16) & 1u;
    bool bad = (st >> 18) & 1u;
    const V3 ro = mk(p.ox[i], p.oy[i], p.oz[i]), rd = mk(p.dx[i], p.dy[i], p.dz[i]);
    const V3 w_o = neg(rd);
    // Only an emitter adds to the path's radiance at the vertex itself (the light sample is parked for the shadow kernel), and an
    // emitter is always of the matte class: the other classes neither read nor write L unless the path ends here.
    constexpr bool kEmits = CLS == CRAY_MAT_MATTE;
    Color3 L = mkc(0.0, 0.0, 0.0);
    if (kEmits) L = mkc(p.L_r[i], p.L_g[i], p.L_b[i]);
    bool L_changed = false;
    Color3 beta = mkc(p.beta_r[i], p.beta_g[i], p.beta_b[i]);

    // the path ends here: into the film right away, unless its light sample is still to be tested (then the shadow kernel flushes)
    auto finish = [&](bool shadow_pending) {
        if (!shadow_pending) {
            if (!kEmits) L = mkc(p.L_r[i], p.L_g[i], p.L_b[i]);
            flush_path(p, job, counters, i, L.r, L.g, L.b, bad);
            return;
        }
        if (L_changed) { p.L_r[i] = L.r; p.L_g[i] = L.g; p.L_b[i] = L.b; }
        p.state[i] = SLOT_DONE | (bounces << 8) | (1u << 17) | ((bad ? 1u : 0u) << 18);
    };

    const LeafPrim lp = load_leaf_prim((job.exact ? s.bin_prims : s.wide_prims) + slot);
    V3 location, normal;
    double tu, tv;
    double hit_t;
    double bu = 0.0, bv = 0.0;
    if ((lp.kind & 0xFFu) == PRIM_TRIANGLE) {
        // the distance and barycentrics the traversal computed for this hit (same code, same bits: the stored distance is not even
        // read).  F32 mode: the f32 traversal chose the triangle, t / u / v are the f64 values of that choice
        if (!triangle_eval(lp.d, ro, rd, hit_t, bu, bv)) {
            hit_t = p.hit_t[i];
            if (job.f32) triangle_eval_unchecked(lp.d, ro, rd, hit_t, bu, bv);
        }
    } else {
        hit_t = p.hit_t[i];
    }
    // Primitive's material / light binding: triangles carry it in the first sector of their shading record
    int32_t material_index, area_light;
    if ((lp.kind & 0xFFu) == PRIM_TRIANGLE) {
        const int2 ids = __ldg(reinterpret_cast<const int2*>(&s.tri_shade[lp.prim].material));
        material_index = ids.x; area_light = ids.y;
    } else {
        const cray_primitive_desc prim = s.prims[lp.prim];
        material_index = prim.material; area_light = prim.area_light;
    }
    const DevMaterial& material = s.materials[material_index];
    // every lobe perfectly specular (glass, metal): known at compile time for those classes
    const bool all_delta = CLS == CRAY_MAT_GLASS || CLS == CRAY_MAT_METAL ? true : (CLS == CRAY_MAT_MATTE ? false : material.all_delta != 0u);
    surface_at(s, lp, ro, rd, hit_t, bu, bv, location, normal, tu, tv, material.needs_uv != 0u);

    // PathSegmentSamples::from path_integrator.rs:25-36 -- dimensions 4 + 8 * bounces ... (evaluated where consumed)
    const VertexSamples vs{job.sobol, p.shuffled_rev[i], p.hash[i], 4u + 8u * bounces};

    // emission (:106-126); an emitting primitive carries the black matte (primitive.rs:43-46), so it is only ever met by the matte class
    if (CLS == CRAY_MAT_MATTE && area_light >= 0) {
        const DevLight& light = s.lights[area_light];
        const Color3 Le = mkc(light.color[0], light.color[1], light.color[2]);
        if (!is_black(Le)) {
            if (is_specular_bounce) {
                L = L + beta * Le;
            } else {
                const double light_pdf = light_pdf_li(s, light, location, normal, w_o).value * light_pick_pdf(s, (uint32_t)area_light);
                const double weight = power_heuristic(light_pdf, p.prev_bsdf_pdf[i]);
                L = L + beta * Le * weight;
            }
            L_changed = true;
        }
    }

    // next-event estimation (:129-164): the shadow ray is traced by the shadow stage, the contribution is parked.
    // Where every lobe of the material is perfectly specular, Material::f is black for every direction, so the light sample
    // cannot contribute: it is not drawn (the reference's shadow ray, :141, is still counted).
    bool shadow_pending = false;   // (the reference's shadow ray of this vertex is counted from the queue lengths, k_begin_iteration)
    // Both rays of this vertex start at `location`.  If it lies in the outer shell of a marked node box, only the reference-order
    // traversal reproduces what the reference's box test does to them (bvh_build.hpp "planar contact"); directions are unit vectors.
    const bool contact = !job.exact && (lp.kind & kKindContact) && origin_in_contact_shell(s, location, kContactTol);
    if (!all_delta) {
        double light_sampler_pdf;
        // a single light is picked whatever the sample value is (the search of light.rs:203-211 ends at 0 for every u < 1)
        const uint32_t light_index = light_pick(s, s.n_lights > 1 ? vs.get(VertexSamples::LIGHT_INDEX) : 0.0, light_sampler_pdf);
        const DevLight& light = s.lights[light_index];
        const LightSample ls = light_sample_li(s, light, vs, location, normal, bad);
        Color3 contribution = mkc(0.0, 0.0, 0.0);
        const Color3 f = material_f<LOBES>(s, material, w_o, ls.w_i, normal, tu, tv);
        const double cos_theta = fabs(dot(ls.w_i, normal));
        if (!ls.pdf.delta) {
            if (ls.pdf.value > 0.0) {
                const double light_pdf = ls.pdf.value * light_sampler_pdf;
                const PdfValue bp = material_pdf(material, w_o, ls.w_i, normal);
                const double bsdf_pdf = bp.delta ? 0.0 : bp.value;
                const double weight = power_heuristic(light_pdf, bsdf_pdf);
                contribution = beta * ls.Li * f * cos_theta * weight / light_pdf;
            }
        } else {
            contribution = beta * ls.Li * f * cos_theta / light_sampler_pdf;
        }
        // traced only when an unoccluded result could change L
        if (!is_black(contribution) || !is_finite3(contribution)) {
            shadow_pending = true;
            if (contact) queue_append_back(&counters->n_shadow_contact, p.shadow_queue, p.capacity, i);
            else queue_append(&counters->n_shadow, p.shadow_queue, i);
            p.sdx[i] = ls.w_i.x; p.sdy[i] = ls.w_i.y; p.sdz[i] = ls.w_i.z;
            p.smax[i] = ls.shadow_max;
            p.sc_r[i] = contribution.r; p.sc_g[i] = contribution.g; p.sc_b[i] = contribution.b;
        }
    }
    // shadow ray and continuation ray both start at the hit location (no offset, :191)
    p.ox[i] = location.x; p.oy[i] = location.y; p.oz[i] = location.z;

    // BSDF sample (:167-195)
    SurfaceSample ss;
    if (!material_sample<LOBES>(s, material, vs, w_o, normal, tu, tv, ss, bad)) { finish(shadow_pending); return; }
    if (is_black(ss.f)) { finish(shadow_pending); return; }
    const double cos_theta = fabs(dot(ss.w_i, normal));
    const double bsdf_pdf = ss.pdf.delta ? 1.0 : ss.pdf.value;
    if (bsdf_pdf == 0.0) { finish(shadow_pending); return; }
    beta = beta * ss.f * cos_theta;
    if (bsdf_pdf != 1.0) beta = beta / bsdf_pdf;   // (x / 1.0 is x, bit for bit: perfectly specular lobes skip three divisions)

    // Russian roulette (:197-206)
    if (bounces > 0) {
        const double max_beta_component = rmax(beta.r, rmax(beta.g, beta.b));
        if (max_beta_component < 1.0) {
            const double q = 1.0 - max_beta_component;
            if (vs.get(VertexSamples::ROULETTE) < q) { finish(shadow_pending); return; }
            beta = beta / (1.0 - q);
        }
    }
    // assert!(L.is_finite()); assert!(beta.is_finite()) (:208-209).  L still lacks this vertex's light sample, which
    // the shadow stage adds; k_generate re-checks L when the path is flushed.
    if ((kEmits ? !is_finite3(L) : (st & kStateBadL) != 0u) || !is_finite3(beta)) { bad = true; finish(shadow_pending); return; }

    const uint32_t next_bounces = bounces + 1;
    if (next_bounces < s.max_depth && !is_black(beta)) {  // loop condition (:54)
        if (L_changed) { p.L_r[i] = L.r; p.L_g[i] = L.g; p.L_b[i] = L.b; }
        p.dx[i] = ss.w_i.x; p.dy[i] = ss.w_i.y; p.dz[i] = ss.w_i.z;
        p.beta_r[i] = beta.r; p.beta_g[i] = beta.g; p.beta_b[i] = beta.b;
        if (!ss.is_specular) p.prev_bsdf_pdf[i] = bsdf_pdf;   // (read by the next vertex only after a non-specular bounce)
        p.state[i] = SLOT_ACTIVE | (next_bounces << 8) | ((ss.is_specular ? 1u : 0u) << 16) | ((shadow_pending ? 1u : 0u) << 17) | ((bad ? 1u : 0u) << 18) |
                     (contact ? kStateContact : 0u);
    } else if (shadow_pending) {
        if (L_changed) { p.L_r[i] = L.r; p.L_g[i] = L.g; p.L_b[i] = L.b; }
        p.state[i] = SLOT_DONE | (next_bounces << 8) | (1u << 17) | ((bad ? 1u : 0u) << 18);
    } else {
        if (!kEmits) L = mkc(p.L_r[i], p.L_g[i], p.L_b[i]);
        flush_path(p, job, counters, i, L.r, L.g, L.b, bad);
    }
}

// Resident CTAs per SM the register allocation of each class kernel is held to (3 = 80 registers: measured against 2, 4 and 6 --
// shade 24.5 ms per 256-spp dragon frame against 25.1, 26.5 and 29.4; the spills of the 64-register build cost more than its occupancy gains)
#ifndef CRAY_SHADE_BLOCKS_MATTE
#define CRAY_SHADE_BLOCKS_MATTE 3
#endif
#ifndef CRAY_SHADE_BLOCKS_GLASS
#define CRAY_SHADE_BLOCKS_GLASS 3
#endif
#ifndef CRAY_SHADE_BLOCKS_PLASTIC
#define CRAY_SHADE_BLOCKS_PLASTIC 3
#endif
#ifndef CRAY_SHADE_BLOCKS_METAL
#define CRAY_SHADE_BLOCKS_METAL 3
#endif
__host__ __device__ constexpr int shade_blocks_of_class(uint32_t cls) {
    return cls == CRAY_MAT_MATTE ? CRAY_SHADE_BLOCKS_MATTE : cls == CRAY_MAT_GLASS ? CRAY_SHADE_BLOCKS_GLASS : cls == CRAY_MAT_PLASTIC ? CRAY_SHADE_BLOCKS_PLASTIC : CRAY_SHADE_BLOCKS_METAL;
}

template <uint32_t CLS>
__global__ void __launch_bounds__(kShadeThreads, shade_blocks_of_class(CLS)) k_shade_class(const __grid_constant__ SceneView s, const __grid_constant__ Pool p, const __grid_constant__ Job job, Counters* counters) {
    // the three shape queues of this class, walked as one range
    const uint64_t n0 = counters->class_count[CLS * 3u], n1 = counters->class_count[CLS * 3u + 1u], n2 = counters->class_count[CLS * 3u + 2u];
    const uint64_t n = n0 + n1 + n2;
    for (uint64_t t = (uint64_t)blockIdx.x * kShadeThreads + threadIdx.x; t < n; t += (uint64_t)gridDim.x * kShadeThreads) {
        const uint32_t shape = t < n0 ? 0u : (t < n0 + n1 ? 1u : 2u);
        const uint64_t local = t - (shape == 0u ? 0ull : (shape == 1u ? n0 : n0 + n1));
        const uint32_t i = p.class_queue[(uint64_t)(CLS * 3u + shape) * p.capacity + local];
        shade_vertex<CLS>(s, p, job, counters, i, p.hit_slot[i]);
    }
}

// The paths that left the scene (path_integrator.rs:60-90).  Paths start pixel by pixel, so the lanes of a warp mostly end on the
// same pixel: the radiances of each run of lanes on one pixel are summed in the warp first, and the run gets one set of film
// atomics instead of up to 32 on the same three addresses (the order of the f64 film sums is free either way).
__global__ void __launch_bounds__(kShadeThreads) k_shade_miss(const __grid_constant__ SceneView s, const __grid_constant__ Pool p, const __grid_constant__ Job job, Counters* counters) {
    const uint64_t n = counters->class_count[kKeyMiss];
    const unsigned lane = threadIdx.x & 31u;
    for (uint64_t base = (uint64_t)blockIdx.x * kShadeThreads + (threadIdx.x & ~31u); base < n; base += (uint64_t)gridDim.x * kShadeThreads) {
        const uint64_t t = base + lane;
        const bool valid = t < n;
        uint32_t i = 0, pixel = 0xFFFFFFFFu;
        Color3 L = mkc(0.0, 0.0, 0.0);
        bool good = false;
        if (valid) {
            i = p.class_queue[(uint64_t)kKeyMiss * p.capacity + t];
            const uint32_t st = p.state[i];
            const bool is_specular_bounce = (st >> 16) & 1u;
            const bool bad = (st >> 18) & 1u;
            L = mkc(p.L_r[i], p.L_g[i], p.L_b[i]);
            for (uint32_t li = 0; li < s.n_lights; ++li) {
                const DevLight& light = s.lights[li];
                if (light.kind != CRAY_LIGHT_INFINITE) continue;  // Light::Le is black for every other kind (light.rs:161-168)
                // (throughput and pdf are read only here: a scene without an infinite light flushes its escaped paths from L alone)
                const Color3 beta = mkc(p.beta_r[i], p.beta_g[i], p.beta_b[i]);
                const Color3 Le = mkc(light.color[0], light.color[1], light.color[2]);
                if (is_specular_bounce) {
                    L = L + beta * Le;
                } else if (!is_black(Le)) {
                    const double light_pdf = (kFrac1Pi / 4.0) * light_pick_pdf(s, li);
                    const double weight = power_heuristic(light_pdf, p.prev_bsdf_pdf[i]);
                    L = L + beta * Le * weight;
                }
            }
            good = !bad && is_finite3(L);
            if (!job.film || job.out_rgb || !good) {   // per-sample output, or a dropped sample: the plain way
                flush_path(p, job, counters, i, L.r, L.g, L.b, bad);
                good = false;
            } else {
                pixel = p.pixel[i];
                p.state[i] = SLOT_EMPTY;
            }
        }
        // runs of neighbouring lanes on the same pixel: a segmented sum over each run, its first lane adds it to the film
        const uint32_t key = good ? pixel : 0xFFFFFFFFu;
        const uint32_t before = __shfl_up_sync(0xFFFFFFFFu, key, 1);
        const unsigned heads = __ballot_sync(0xFFFFFFFFu, lane == 0u || key != before);
        const unsigned later = lane == 31u ? 0u : heads & ~((2u << lane) - 1u);
        const unsigned run_end = later ? (unsigned)__ffs(later) - 2u : 31u;   // last lane of this lane's run
        double r = L.r, g = L.g, b = L.b;
#pragma unroll
        for (unsigned d = 1; d < 32u; d <<= 1) {
            const double r2 = __shfl_down_sync(0xFFFFFFFFu, r, d), g2 = __shfl_down_sync(0xFFFFFFFFu, g, d), b2 = __shfl_down_sync(0xFFFFFFFFu, b, d);
            if (lane + d <= run_end) { r += r2; g += g2; b += b2; }
        }
        if (good && ((heads >> lane) & 1u)) {
            double* px = job.film + 3ull * pixel;
            atomicAdd(px, r); atomicAdd(px + 1, g); atomicAdd(px + 2, b);
        }
    }
}

template <bool BACK>
__global__ void __launch_bounds__(128) k_shadow_exact(SceneView s, Pool p, Job job, Counters* counters, const unsigned long long* __restrict__ n_ptr) {
    const uint64_t n = *n_ptr;
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = p.shadow_queue[BACK ? p.capacity - 1u - (uint32_t)q : (uint32_t)q];
        Hit h;
        const bool occluded = traverse_exact<true>(s, mk(p.ox[i], p.oy[i], p.oz[i]), mk(p.sdx[i], p.sdy[i], p.sdz[i]), p.smax[i], h);
        settle_shadow(p, job, counters, i, occluded);
    }
}

__global__ void k_film_to_f32(const double* __restrict__ film, float* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)film[i];
}

// ---- host side ------------------------------------------------------------------------------------------------

struct PoolStorage {
    Pool pool{};
    void* slab = nullptr;
    Counters* d_counters = nullptr;
    Counters* h_counters = nullptr;  // pinned
    double* d_film = nullptr;
    uint64_t film_elems = 0;
    float* d_film_f32 = nullptr;      // staging of cray_render's host film
    unsigned persistent_blocks = 0;   // SMs x resident CTAs of the persistent traversal kernel
    WideTuning tune{12, 8};
    unsigned shadow_blocks = 0;       // the any-hit instantiation needs fewer registers: its own occupancy
    unsigned f32_blocks = 0, f32_shadow_blocks = 0;  // F32 mode instantiations
    unsigned shade_blocks = 0;        // grid of the per-class shade kernels (grid-stride loops over their queues)
    unsigned long long* d_trace_counters = nullptr;  // {n, cursor, n of the reference-order launch} for the S3 entry points
    uint32_t* d_trace_list = nullptr;                // S3 ray lists of k_classify_rays (grow only)
    uint64_t trace_list_capacity = 0;
    bool initialised = false;
    std::vector<cudaEvent_t> timers;                 // stage boundaries of every iteration of one render, reused across calls
};

// Resident CTAs per SM of a persistent traversal kernel -- and the shared-memory carve-out asked for is exactly what those
// CTAs need, so that the rest of the 256 KB stays L1.  Left to itself the driver sizes the carve-out for the occupancy shared
// memory alone would allow (closest-hit kernel: 132 KB for 7 CTAs where registers admit 5 -- 35 KB of L1 given away).
// CRAY_CARVEOUT=0 keeps the driver's choice (tuning).
template <class Kernel>
int resident_blocks(Kernel kernel, int* per_sm) {
    CRAY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kernel, 128, 0));
    *per_sm = std::max(*per_sm, 1);
    const char* e = std::getenv("CRAY_CARVEOUT");
    if (e && std::atoi(e) == 0) return CRAY_OK;
    cudaFuncAttributes attr{};
    CRAY_CUDA(cudaFuncGetAttributes(&attr, kernel));
    const size_t need = (size_t)*per_sm * (attr.sharedSizeBytes + 1024);  // + the 1 KB the system reserves per CTA
    const int percent = (int)std::min<size_t>(100, (need * 100 + 228 * 1024 - 1) / (228 * 1024));
    CRAY_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, percent));
    return CRAY_OK;
}

constexpr int kPoolNoMemory = -100;   // (internal) the device could not hold a pool of the requested capacity
int ensure_pool(cray_scene* sc, uint32_t capacity) {
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    if (!ps) { ps = new PoolStorage(); sc->pool = ps; }
    if (!ps->initialised) {  // (set once everything below has succeeded: a retry after a failure starts over)
        if (!ps->d_counters) CRAY_CUDA(cudaMalloc(&ps->d_counters, sizeof(Counters)));
        if (!ps->h_counters) CRAY_CUDA(cudaMallocHost(&ps->h_counters, sizeof(Counters)));
        int sms = 0;
        CRAY_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device));
        int per_sm = 0;
        int rc = resident_blocks(k_wide_persistent<false, ExtendSource>, &per_sm);
        if (rc != CRAY_OK) return rc;
        ps->persistent_blocks = (unsigned)(sms * per_sm);
        if ((rc = resident_blocks(k_wide_persistent<true, ShadowSource>, &per_sm)) != CRAY_OK) return rc;
        ps->shadow_blocks = (unsigned)(sms * per_sm);
        if ((rc = resident_blocks(k_wide_persistent<false, ExtendSource, true>, &per_sm)) != CRAY_OK) return rc;
        ps->f32_blocks = (unsigned)(sms * per_sm);
        if ((rc = resident_blocks(k_wide_persistent<true, RayArraySource, true>, &per_sm)) != CRAY_OK) return rc;
        ps->f32_shadow_blocks = (unsigned)(sms * per_sm);
        // the S3 instantiations share the budgets of their wavefront twins
        int unused = 0;
        if ((rc = resident_blocks(k_wide_persistent<false, RayArraySource>, &unused)) != CRAY_OK) return rc;
        if ((rc = resident_blocks(k_wide_persistent<true, RayArraySource>, &unused)) != CRAY_OK) return rc;
        if ((rc = resident_blocks(k_wide_persistent<false, RayArraySource, true>, &unused)) != CRAY_OK) return rc;
        if (const char* e = std::getenv("CRAY_REFILL_LANES")) ps->tune.refill_lanes = std::max(1, std::min(32, std::atoi(e)));
        if (const char* e = std::getenv("CRAY_WAIT_LANES")) ps->tune.wait_lanes = std::max(1, std::min(33, std::atoi(e)));
        if (const char* e = std::getenv("CRAY_BLOCKS_PER_SM")) ps->persistent_blocks = ps->shadow_blocks = ps->f32_blocks = ps->f32_shadow_blocks = (unsigned)(sms * std::max(1, std::atoi(e)));
        ps->shade_blocks = (unsigned)(sms * 16);
        ps->initialised = true;
    }
    if (ps->pool.capacity >= capacity) return CRAY_OK;
    if (ps->slab) { cudaDeviceSynchronize(); cudaFree(ps->slab); ps->slab = nullptr; }  // (re-)allocation is rare: grow only
    const size_t n = capacity;
    const size_t n_f64 = 21, n_u32 = 9 + 13;  // (+ the 13 queues of the shade stage)
    const size_t bytes = n * (n_f64 * 8 + n_u32 * 4);
    if (cudaMalloc(&ps->slab, bytes) != cudaSuccess) {   // the caller retries with a smaller pool
        cudaGetLastError();
        ps->slab = nullptr;
        ps->pool.capacity = 0;
        if (capacity <= (1u << 20)) { set_error("out of device memory for the path pool"); return CRAY_E_CUDA; }
        return kPoolNoMemory;
    }
    // cleared before anyone can touch it, whichever stream the caller renders on
    CRAY_CUDA(cudaMemsetAsync(ps->slab, 0, bytes, sc->stream));
    CRAY_CUDA(cudaStreamSynchronize(sc->stream));
    double* f = static_cast<double*>(ps->slab);
    Pool& p = ps->pool;
    double** fields[] = {&p.ox, &p.oy, &p.oz, &p.dx, &p.dy, &p.dz, &p.hit_t, &p.beta_r, &p.beta_g, &p.beta_b,
                         &p.L_r, &p.L_g, &p.L_b, &p.prev_bsdf_pdf, &p.sdx, &p.sdy, &p.sdz, &p.smax, &p.sc_r, &p.sc_g, &p.sc_b};
    size_t k = 0;
    for (double** fp : fields) { *fp = f + k * n; ++k; }
    uint32_t* u = reinterpret_cast<uint32_t*>(f + n_f64 * n);
    uint32_t** ufields[] = {&p.hit_slot, &p.id, &p.pixel, &p.hash, &p.shuffled_rev, &p.state, &p.extend_queue, &p.extend_queue_prev, &p.shadow_queue};
    k = 0;
    for (uint32_t** up : ufields) { *up = u + k * n; ++k; }
    p.class_queue = u + k * n;
    p.capacity = capacity;
    return CRAY_OK;
}

// Runs the wavefront until every sample of `job` has been flushed.
int run_wavefront(cray_scene* sc, Job job, uint32_t capacity, cudaStream_t stream, cray_render_stats* stats) {
    int rc = ensure_pool(sc, capacity);
    while (rc == kPoolNoMemory && capacity > (1u << 20)) {   // a device short of memory renders with fewer paths in flight
        capacity = (capacity + 1u) / 2u;
        rc = ensure_pool(sc, capacity);
    }
    if (rc == kPoolNoMemory) { set_error("out of device memory for the path pool"); return CRAY_E_CUDA; }
    if (rc != CRAY_OK) return rc;
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    Pool pool = ps->pool;
    pool.capacity = capacity;
    Counters* dc = ps->d_counters;
    CRAY_CUDA(cudaMemsetAsync(pool.state, 0, sizeof(uint32_t) * capacity, stream));
    CRAY_CUDA(cudaMemsetAsync(dc, 0, sizeof(Counters), stream));
    struct Events {  // destroyed on every return path
        cudaEvent_t e0 = nullptr, e1 = nullptr, gen_done = nullptr;
        ~Events() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (gen_done) cudaEventDestroy(gen_done); }
    } ev;
    CRAY_CUDA(cudaEventCreate(&ev.e0)); CRAY_CUDA(cudaEventCreate(&ev.e1));
    CRAY_CUDA(cudaEventCreateWithFlags(&ev.gen_done, cudaEventDisableTiming));
    CRAY_CUDA(cudaEventRecord(ev.e0, stream));
    const unsigned g256 = (capacity + kGenBlock - 1) / kGenBlock;   // k_generate: 1024 slots per block
    const unsigned gx = (unsigned)std::min<uint64_t>(kExactBlocks, (capacity + 127) / 128);
    const unsigned gp = ps->persistent_blocks, gs = ps->shadow_blocks;
    // fast mode: rays that start in a contact shell go to the reference-order kernels (only scenes with marked primitives have any)
    const bool contact = !job.exact && sc->info.contact_primitives > 0;
    uint64_t iterations = 0, launches = 0, closest = 0;
    // once every sample of the job has been started (known from the counters of the previous iteration), k_requeue replaces k_generate
    bool exhausted = false;
    uint64_t prev_front = 0, prev_contact = 0;
    const bool timed = stats != nullptr;
    // The host never waits for traversal or shading: it enqueues the whole iteration, then waits only for the iteration's
    // k_generate (long finished by the time the GPU works through extend / shade / shadow) to learn whether any path is
    // still alive.  The launches behind the last, empty generate find empty queues and return at once.
    constexpr size_t kMarks = 5;  // per iteration: before generate | extend | shade | shadow | after shadow
    for (size_t iter = 0;; ++iter) {
        if (timed) {
            while (ps->timers.size() < kMarks * (iter + 1)) {
                cudaEvent_t t;
                CRAY_CUDA(cudaEventCreate(&t));
                ps->timers.push_back(t);
            }
            CRAY_CUDA(cudaEventRecord(ps->timers[kMarks * iter], stream));
        }
        k_begin_iteration<<<1, 1, 0, stream>>>(dc);
        std::swap(pool.extend_queue, pool.extend_queue_prev);
        if (!exhausted) k_generate<<<g256, 256, 0, stream>>>(sc->view, pool, job, dc);
        else k_requeue<<<(unsigned)((prev_front + prev_contact + kGenBlock - 1) / kGenBlock), 256, 0, stream>>>(pool, job, dc, prev_front, prev_contact);
        CRAY_CUDA(cudaMemcpyAsync(ps->h_counters, dc, sizeof(Counters), cudaMemcpyDeviceToHost, stream));
        CRAY_CUDA(cudaEventRecord(ev.gen_done, stream));
        if (timed) CRAY_CUDA(cudaEventRecord(ps->timers[kMarks * iter + 1], stream));
        if (job.exact) k_extend_exact<false><<<gx, 128, 0, stream>>>(sc->view, pool, &dc->n_extend);
        else if (job.f32) k_wide_persistent<false, ExtendSource, true><<<ps->f32_blocks, 128, 0, stream>>>(sc->view, ExtendSource{pool}, &dc->n_extend, &dc->extend_cursor, ps->tune);
        else k_wide_persistent<false, ExtendSource><<<gp, 128, 0, stream>>>(sc->view, ExtendSource{pool}, &dc->n_extend, &dc->extend_cursor, ps->tune);
        if (contact) k_extend_exact<true><<<gx, 128, 0, stream>>>(sc->view, pool, &dc->n_extend_contact);
        if (timed) CRAY_CUDA(cudaEventRecord(ps->timers[kMarks * iter + 2], stream));
        {
            const unsigned blocks = (capacity + kShadeThreads - 1) / kShadeThreads;
            const unsigned strided = std::min(blocks, ps->shade_blocks);   // the class kernels walk their queues with a grid stride
            // (the queue's length lives on the device; it cannot exceed the previous iteration's once no new paths start)
            const uint64_t most = exhausted ? prev_front + prev_contact : capacity;
            k_shade_classify<<<(unsigned)((most + kClassifyBlock - 1) / kClassifyBlock), kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
            k_shade_class<CRAY_MAT_METAL><<<strided, kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
            k_shade_class<CRAY_MAT_MATTE><<<strided, kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
            k_shade_class<CRAY_MAT_PLASTIC><<<strided, kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
            k_shade_class<CRAY_MAT_GLASS><<<strided, kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
            k_shade_miss<<<strided, kShadeThreads, 0, stream>>>(sc->view, pool, job, dc);
        }
        if (timed) CRAY_CUDA(cudaEventRecord(ps->timers[kMarks * iter + 3], stream));
        // at most one shadow ray per shaded vertex; the queue lengths live on the device
        if (job.exact) k_shadow_exact<false><<<gx, 128, 0, stream>>>(sc->view, pool, job, dc, &dc->n_shadow);
        // (F32 mode traces its shadow rays with the f64 any-hit kernel: measured faster than the f32 one, 20.9 against 24.0 ms per
        // 256-spp dragon frame -- the f32 instantiation's larger shared-memory footprint leaves it less L1 -- and exact at the
        // light's end of the ray.  The f32 any-hit kernel serves cray_trace_any.)
        else k_wide_persistent<true, ShadowSource><<<gs, 128, 0, stream>>>(sc->view, ShadowSource{pool, job, dc}, &dc->n_shadow, &dc->shadow_cursor, ps->tune);
        if (contact) k_shadow_exact<true><<<gx, 128, 0, stream>>>(sc->view, pool, job, dc, &dc->n_shadow_contact);
        if (timed) CRAY_CUDA(cudaEventRecord(ps->timers[kMarks * iter + 4], stream));
        launches += contact ? 12 : 10;
        CRAY_CUDA(cudaEventSynchronize(ev.gen_done));
        const uint64_t live = ps->h_counters->n_extend + ps->h_counters->n_extend_contact;
        if (live == 0) break;
        prev_front = ps->h_counters->n_extend; prev_contact = ps->h_counters->n_extend_contact;
        exhausted = ps->h_counters->next_id >= job.n_total;
        closest += live;
        iterations += 1;
    }
    CRAY_CUDA(cudaEventRecord(ev.e1, stream));
    CRAY_CUDA(cudaEventSynchronize(ev.e1));
    CRAY_CUDA(cudaGetLastError());
    float ms = 0;
    CRAY_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    if (stats) {
        double part[4] = {0.0, 0.0, 0.0, 0.0};  // generate, extend, shade, shadow
        for (size_t i = 0; i < iterations; ++i)
            for (size_t k = 0; k < 4; ++k) {
                float t = 0;
                CRAY_CUDA(cudaEventElapsedTime(&t, ps->timers[kMarks * i + k], ps->timers[kMarks * i + k + 1]));
                part[k] += t;
            }
        stats->samples = job.n_total;
        stats->closest_rays = closest;
        stats->shadow_rays = ps->h_counters->shadow_rays;
        stats->shadow_rays_traced = ps->h_counters->shadow_traced;
        stats->contact_rays = ps->h_counters->contact_rays;
        stats->nan_samples = ps->h_counters->nan_samples;
        stats->iterations = iterations;
        stats->kernel_launches = launches;
        stats->render_ms = ms;
        stats->generate_ms = part[0];
        stats->trace_ms = part[1];
        stats->shade_ms = part[2];
        stats->shadow_ms = part[3];
    }
    return CRAY_OK;
}

int check_mode(const cray_scene* sc, int mode) {
    if (!sc) { set_error("null scene"); return CRAY_E_INVALID; }
    if (mode != CRAY_TRAVERSE_EXACT && mode != CRAY_TRAVERSE_FAST && mode != CRAY_TRAVERSE_F32) { set_error("unknown traversal mode"); return CRAY_E_INVALID; }
    if (mode == CRAY_TRAVERSE_F32 && !(sc->build_flags & CRAY_BUILD_F32)) { set_error("scene was created without CRAY_BUILD_F32"); return CRAY_E_INVALID; }
    if (mode == CRAY_TRAVERSE_FAST && !(sc->build_flags & CRAY_BUILD_FAST)) { set_error("scene was created without CRAY_BUILD_FAST"); return CRAY_E_INVALID; }
    return CRAY_OK;
}

}  // namespace cray

using namespace cray;

extern "C" {

// Tuning builds only (CRAY_WIDE_STATS): copies and clears the traversal work counters, [0..11] closest-hit, [12..23] any-hit.
int cray_debug_wide_stats(unsigned long long* out16) {
#if CRAY_WIDE_STATS
    unsigned long long zero[24] = {};
    CRAY_CUDA(cudaDeviceSynchronize());
    CRAY_CUDA(cudaMemcpyFromSymbol(out16, g_wide_stats, sizeof(zero)));
    CRAY_CUDA(cudaMemcpyToSymbol(g_wide_stats, zero, sizeof(zero)));
    return CRAY_OK;
#else
    (void)out16;
    set_error("built without CRAY_WIDE_STATS");
    return CRAY_E_UNSUPPORTED;
#endif
}

void cray_pool_release(cray_scene* sc) {
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    if (!ps) return;
    if (ps->slab) cudaFree(ps->slab);
    if (ps->d_counters) cudaFree(ps->d_counters);
    if (ps->h_counters) cudaFreeHost(ps->h_counters);
    if (ps->d_film) cudaFree(ps->d_film);
    if (ps->d_film_f32) cudaFree(ps->d_film_f32);
    if (ps->d_trace_counters) cudaFree(ps->d_trace_counters);
    if (ps->d_trace_list) cudaFree(ps->d_trace_list);
    for (cudaEvent_t ev : ps->timers) cudaEventDestroy(ev);
    delete ps;
    sc->pool = nullptr;
    cudaGetLastError();  // a failed release must not surface as the error of an unrelated later call
}

static int launch_trace(cray_scene* sc, int mode, bool any, const cray_ray* d_rays, uint64_t n, cray_hit* d_hits, cray_surface* d_surf, uint8_t* d_occluded, cudaStream_t stream) {
    CRAY_CUDA(cudaSetDevice(sc->device));
    if (n > 0xFFFFFFFFull) { set_error("more than 2^32 rays in one call"); return CRAY_E_INVALID; }
    int rc = ensure_pool(sc, 1);
    if (rc != CRAY_OK) return rc;
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    if (!ps->d_trace_counters) CRAY_CUDA(cudaMalloc(&ps->d_trace_counters, 3 * sizeof(unsigned long long)));
    unsigned long long* dc = ps->d_trace_counters;  // {rays of the wide launch, its cursor, rays of the reference-order launch}
    const unsigned exact_blocks = (unsigned)std::min<uint64_t>(kExactBlocks, (n + 127) / 128);
    if (mode == CRAY_TRAVERSE_EXACT) {
        const unsigned long long init[3] = {0ull, 0ull, n};
        CRAY_CUDA(cudaMemcpyAsync(dc, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        if (any) k_trace_any_exact<<<exact_blocks, 128, 0, stream>>>(sc->view, d_rays, dc + 2, nullptr, n, d_occluded);
        else k_trace_closest_exact<<<exact_blocks, 128, 0, stream>>>(sc->view, d_rays, dc + 2, nullptr, n, d_hits, d_surf);
    } else {
        const bool f32 = mode == CRAY_TRAVERSE_F32;
        // fast / F32 mode on a scene with marked node boxes: the rays that start in a contact shell are traced in reference order
        const bool split = sc->view.bin_contact != nullptr;
        uint32_t* list = nullptr;
        if (split) {
            if (ps->trace_list_capacity < n) {
                if (ps->d_trace_list) { CRAY_CUDA(cudaStreamSynchronize(stream)); cudaFree(ps->d_trace_list); ps->d_trace_list = nullptr; ps->trace_list_capacity = 0; }
                CRAY_CUDA(cudaMalloc(&ps->d_trace_list, n * sizeof(uint32_t)));
                ps->trace_list_capacity = n;
            }
            list = ps->d_trace_list;
            CRAY_CUDA(cudaMemsetAsync(dc, 0, 3 * sizeof(unsigned long long), stream));
            k_classify_rays<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc->view, d_rays, n, list, dc);
        } else {
            const unsigned long long init[3] = {n, 0ull, 0ull};
            CRAY_CUDA(cudaMemcpyAsync(dc, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        }
        const RayArraySource src{d_rays, d_hits, d_surf, d_occluded, f32, list};
        const unsigned resident = f32 ? (any ? ps->f32_shadow_blocks : ps->f32_blocks) : (any ? ps->shadow_blocks : ps->persistent_blocks);
        const unsigned blocks = (unsigned)std::min<uint64_t>(resident, (n + 127) / 128);
        if (f32 && any) k_wide_persistent<true, RayArraySource, true><<<blocks, 128, 0, stream>>>(sc->view, src, dc, dc + 1, ps->tune);
        else if (f32) k_wide_persistent<false, RayArraySource, true><<<blocks, 128, 0, stream>>>(sc->view, src, dc, dc + 1, ps->tune);
        else if (any) k_wide_persistent<true, RayArraySource><<<blocks, 128, 0, stream>>>(sc->view, src, dc, dc + 1, ps->tune);
        else k_wide_persistent<false, RayArraySource><<<blocks, 128, 0, stream>>>(sc->view, src, dc, dc + 1, ps->tune);
        if (split) {
            if (any) k_trace_any_exact<<<exact_blocks, 128, 0, stream>>>(sc->view, d_rays, dc + 2, list, n, d_occluded);
            else k_trace_closest_exact<<<exact_blocks, 128, 0, stream>>>(sc->view, d_rays, dc + 2, list, n, d_hits, d_surf);
        }
    }
    CRAY_CUDA(cudaGetLastError());
    return CRAY_OK;
}

int cray_trace_closest_device(cray_scene* sc, int mode, const cray_ray* d_rays, uint64_t n, cray_hit* d_hits, cray_surface* d_surf, void* stream) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (n == 0) return CRAY_OK;
    if (!d_rays || !d_hits) { set_error("null buffer"); return CRAY_E_INVALID; }
    return launch_trace(sc, mode, false, d_rays, n, d_hits, d_surf, nullptr, (cudaStream_t)stream);
}

int cray_trace_any_device(cray_scene* sc, int mode, const cray_ray* d_rays, uint64_t n, uint8_t* d_occluded, void* stream) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (n == 0) return CRAY_OK;
    if (!d_rays || !d_occluded) { set_error("null buffer"); return CRAY_E_INVALID; }
    return launch_trace(sc, mode, true, d_rays, n, nullptr, nullptr, d_occluded, (cudaStream_t)stream);
}

int cray_trace_closest(cray_scene* sc, int mode, const cray_ray* rays, uint64_t n, cray_hit* hits, cray_surface* surf) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (n == 0) return CRAY_OK;
    if (!rays || !hits) { set_error("null buffer"); return CRAY_E_INVALID; }
    CRAY_CUDA(cudaSetDevice(sc->device));
    cray_ray* d_rays = nullptr;
    cray_hit* d_hits = nullptr;
    cray_surface* d_surf = nullptr;
    CRAY_CUDA(cudaMalloc(&d_rays, n * sizeof(cray_ray)));
    cudaError_t e = cudaMalloc(&d_hits, n * sizeof(cray_hit));
    if (e == cudaSuccess && surf) e = cudaMalloc(&d_surf, n * sizeof(cray_surface));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, n * sizeof(cray_ray), cudaMemcpyHostToDevice, sc->stream);
    if (e == cudaSuccess) {
        rc = cray_trace_closest_device(sc, mode, d_rays, n, d_hits, d_surf, sc->stream);
        if (rc == CRAY_OK) e = cudaMemcpyAsync(hits, d_hits, n * sizeof(cray_hit), cudaMemcpyDeviceToHost, sc->stream);
        if (rc == CRAY_OK && e == cudaSuccess && surf) e = cudaMemcpyAsync(surf, d_surf, n * sizeof(cray_surface), cudaMemcpyDeviceToHost, sc->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sc->stream);
    }
    cudaFree(d_rays); cudaFree(d_hits); cudaFree(d_surf);
    if (e != cudaSuccess) return cuda_fail(e, "cray_trace_closest");
    return rc;
}

int cray_trace_any(cray_scene* sc, int mode, const cray_ray* rays, uint64_t n, uint8_t* occluded) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (n == 0) return CRAY_OK;
    if (!rays || !occluded) { set_error("null buffer"); return CRAY_E_INVALID; }
    CRAY_CUDA(cudaSetDevice(sc->device));
    cray_ray* d_rays = nullptr;
    uint8_t* d_occ = nullptr;
    CRAY_CUDA(cudaMalloc(&d_rays, n * sizeof(cray_ray)));
    cudaError_t e = cudaMalloc(&d_occ, n);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rays, rays, n * sizeof(cray_ray), cudaMemcpyHostToDevice, sc->stream);
    if (e == cudaSuccess) {
        rc = cray_trace_any_device(sc, mode, d_rays, n, d_occ, sc->stream);
        if (rc == CRAY_OK) e = cudaMemcpyAsync(occluded, d_occ, n, cudaMemcpyDeviceToHost, sc->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sc->stream);
    }
    cudaFree(d_rays); cudaFree(d_occ);
    if (e != cudaSuccess) return cuda_fail(e, "cray_trace_any");
    return rc;
}

int cray_estimate_li(cray_scene* sc, int mode, uint64_t seed, const uint32_t* x, const uint32_t* y, const uint32_t* sample_index, uint64_t n, double* rgb) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (n == 0) return CRAY_OK;
    if (!x || !y || !sample_index || !rgb) { set_error("null buffer"); return CRAY_E_INVALID; }
    if (n > 0xFFFFFFFFull) { set_error("too many samples in one call"); return CRAY_E_INVALID; }
    CRAY_CUDA(cudaSetDevice(sc->device));
    uint32_t* d_list = nullptr;
    double* d_rgb = nullptr;
    CRAY_CUDA(cudaMalloc(&d_list, 3 * n * sizeof(uint32_t)));
    cudaError_t e = cudaMalloc(&d_rgb, 3 * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_list, x, n * 4, cudaMemcpyHostToDevice, sc->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_list + n, y, n * 4, cudaMemcpyHostToDevice, sc->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_list + 2 * n, sample_index, n * 4, cudaMemcpyHostToDevice, sc->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_rgb, 0, 3 * n * sizeof(double), sc->stream);
    if (e == cudaSuccess) {
        Job job{};
        job.seed = seed; job.n_total = n;
        job.lx = d_list; job.ly = d_list + n; job.ls = d_list + 2 * n;
        job.out_rgb = d_rgb;
        job.sobol = sc->d_sobol;
        job.exact = mode == CRAY_TRAVERSE_EXACT;
        job.f32 = mode == CRAY_TRAVERSE_F32;
        const uint32_t capacity = (uint32_t)std::min<uint64_t>(n, 1u << 20);
        rc = run_wavefront(sc, job, capacity, sc->stream, nullptr);
        if (rc == CRAY_OK) e = cudaMemcpy(rgb, d_rgb, 3 * n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_list); cudaFree(d_rgb);
    if (e != cudaSuccess) return cuda_fail(e, "cray_estimate_li");
    return rc;
}

int cray_render_device(cray_scene* sc, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float* d_rgb_sum, void* stream_, cray_render_stats* stats) {
    int rc = check_mode(sc, mode);
    if (rc != CRAY_OK) return rc;
    if (!d_rgb_sum || sample_end < sample_begin) { set_error("bad arguments"); return CRAY_E_INVALID; }
    if (sample_end > 65536u) { set_error("sample index beyond the sampler's 2^16 points (sobol_burley uses 16 index bits)"); return CRAY_E_INVALID; }
    CRAY_CUDA(cudaSetDevice(sc->device));
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : sc->stream;
    const uint64_t n_pixels = (uint64_t)sc->info.width * sc->info.height;
    const uint64_t n_total = n_pixels * (sample_end - sample_begin);
    if (n_total > 0xFFFFFFFFull) { set_error("more than 2^32 samples in one call: split the sample range"); return CRAY_E_INVALID; }
    rc = ensure_pool(sc, 1);
    if (rc != CRAY_OK) return rc;
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    if (ps->film_elems < n_pixels * 3) {
        if (ps->d_film) cudaFree(ps->d_film);
        CRAY_CUDA(cudaMalloc(&ps->d_film, n_pixels * 3 * sizeof(double)));
        ps->film_elems = n_pixels * 3;
    }
    CRAY_CUDA(cudaMemsetAsync(ps->d_film, 0, n_pixels * 3 * sizeof(double), stream));
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_total > 0) {
        Job job{};
        job.seed = seed; job.n_total = n_total;
        job.sample_begin = sample_begin;
        job.n_pixels = (uint32_t)n_pixels;
        job.pixel_order = sc->d_pixel_order;
        job.n_samples = sample_end - sample_begin;
        job.sample_group = job.n_samples;   // default: all samples of a pixel together
        if (const char* e = std::getenv("CRAY_SAMPLE_GROUP")) job.sample_group = (uint32_t)std::max(1, std::atoi(e));
        job.sample_group = std::min(job.sample_group, job.n_samples);
        job.film = ps->d_film;
        job.sobol = sc->d_sobol;
        job.exact = mode == CRAY_TRAVERSE_EXACT;
        job.f32 = mode == CRAY_TRAVERSE_F32;
        // path slots in flight: up to 2^28 (256 B of path state and queues per slot: a third of the 180 GB for the 245.8 M samples of the
        // headline frame, which then is ONE wave -- every launch as long as it can be, the tails of the persistent traversal kernels
        // amortised; measured 2^22 .. 2^28, profiles/r2d_pool_sweep.txt; a device short of memory gets a smaller pool, run_wavefront);
        // CRAY_POOL_LOG2 overrides it for tuning
        uint32_t pool_log2 = 28;
        if (const char* e = std::getenv("CRAY_POOL_LOG2")) pool_log2 = (uint32_t)std::max(10, std::min(28, std::atoi(e)));
        const uint32_t capacity = (uint32_t)std::min<uint64_t>(n_total, 1ull << pool_log2);
        rc = run_wavefront(sc, job, capacity, stream, stats);
        if (rc != CRAY_OK) return rc;
    }
    k_film_to_f32<<<(unsigned)((n_pixels * 3 + 255) / 256), 256, 0, stream>>>(ps->d_film, d_rgb_sum, n_pixels * 3);
    if (stats) stats->kernel_launches += 1;
    CRAY_CUDA(cudaStreamSynchronize(stream));
    CRAY_CUDA(cudaGetLastError());
    return CRAY_OK;
}

// The scene's own device film (W*H*3 f32), allocated on first use; null on failure (see cray_last_error).
float* cray_scene_film_f32(cray_scene* sc) {
    if (!sc || cudaSetDevice(sc->device) != cudaSuccess) { set_error("bad scene / device"); return nullptr; }
    if (ensure_pool(sc, 1) != CRAY_OK) return nullptr;
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    if (!ps->d_film_f32) {  // film size is fixed per scene
        const cudaError_t e = cudaMalloc(&ps->d_film_f32, (size_t)sc->info.width * sc->info.height * 3 * sizeof(float));
        if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc(film)"); return nullptr; }
    }
    return ps->d_film_f32;
}

int cray_render(cray_scene* sc, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end, float* rgb_sum, cray_render_stats* stats) {
    if (!sc || !rgb_sum) { set_error("bad arguments"); return CRAY_E_INVALID; }
    CRAY_CUDA(cudaSetDevice(sc->device));
    const uint64_t n = (uint64_t)sc->info.width * sc->info.height * 3;
    if (!cray_scene_film_f32(sc)) return CRAY_E_CUDA;
    auto* ps = static_cast<PoolStorage*>(sc->pool);
    int rc = cray_render_device(sc, mode, seed, sample_begin, sample_end, ps->d_film_f32, sc->stream, stats);
    if (rc != CRAY_OK) return rc;
    CRAY_CUDA(cudaMemcpyAsync(rgb_sum, ps->d_film_f32, n * sizeof(float), cudaMemcpyDeviceToHost, sc->stream));
    CRAY_CUDA(cudaStreamSynchronize(sc->stream));
    return CRAY_OK;
}

}  // extern "C"
