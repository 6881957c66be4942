// The reference's binary SAH tree (src/bvh.rs:234-336 with src/util.rs partition_by) built on the GPU, decision for decision.
//
// Every reduction of BvhNode::from_sah_splitting is a min / max or a count, the bucket index and the costs are a fixed sequence of
// f64 operations on them (this file is compiled with --fmad=false like the rest of the library), and the two-pointer partition
// swaps the k-th misplaced item of the left part with the k-th misplaced item of the right part counted from the back -- so the
// tree does not depend on how the work is scheduled, and a GPU build returns the same nodes, leaf order and split axes as the
// host's (tests/test_gpu_bvh_build.py compares them node for node).
//
//   phase A  level-synchronous over all segments (= nodes being split) of more than kSmall items: per level one pass that
//            accumulates node and centroid bounds (ordered-integer atomics, aggregated per warp and per block), one that fills the
//            twelve buckets, a per-segment cost evaluation, and the partition (flags, two exclusive scans, pairing, swap)
//   phase B  one thread per segment of at most kSmall items runs the reference's sequential recursion on it and writes the
//            sub-tree in local pre-order
//   assembly the (small) top of the tree is numbered in pre-order on the host; one kernel writes the final node array
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bvh_build.hpp"

namespace cray {
namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kSmall = 128;             // segments up to this size are finished by one thread each
constexpr uint32_t kBuckets = 12;            // bvh.rs:235
constexpr double kTraversalCost = 1.0 / 8.0; // bvh.rs:236
constexpr uint32_t kMaxLeaf = 4;             // bvh.rs:237
enum : uint32_t { KIND_ACTIVE = 0, KIND_SPLIT = 1, KIND_SMALL = 2 };
enum : uint32_t { ERR_NONE = 0, ERR_AREA = 1, ERR_COST = 2, ERR_EMPTY_SIDE = 3, ERR_PAIRS = 4, ERR_CAPACITY = 5 };

struct GNode {  // a node of the top of the tree
    double lo[3], hi[3];
    uint32_t begin, count;
    uint32_t left, right;   // SPLIT: child node ids; SMALL: left = number of nodes of its sub-tree
    uint32_t axis, kind;
};

struct Scratch {  // per active segment of the current level
    unsigned long long acc[12];                       // node box lo (min) x3, hi (max) x3, centroid box lo x3, hi x3 as ordered keys
    unsigned long long blo[kBuckets][3], bhi[kBuckets][3];
    unsigned int bcount[kBuckets];
    uint32_t node, best, mid;
    int axis;
    double cmin, cext, total_area;
};

// doubles ordered like unsigned integers
__host__ __device__ inline unsigned long long key_of(double x) {
    unsigned long long b;
#if defined(__CUDA_ARCH__)
    b = (unsigned long long)__double_as_longlong(x);
#else
    std::memcpy(&b, &x, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ inline double value_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}
constexpr unsigned long long kKeyMaxInit = 0ull, kKeyMinInit = ~0ull;

__device__ inline double centroid_of(const BuildItem& it, int axis) { return (it.lo[axis] + it.hi[axis]) * 0.5; }
__device__ inline double surface_area(const double* lo, const double* hi) {  // bounds.rs:29-32
    const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0 * (dx * dy + dy * dz + dz * dx);
}
__device__ inline int maximum_extent(const double* lo, const double* hi) {   // bounds.rs:36-45
    const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx > dy && dx > dz) return 0;
    if (dy > dz) return 1;
    return 2;
}
__device__ inline uint32_t bucket_of(double c, double cmin, double cext) {
    const double offset = (c - cmin) / cext;                    // Bounds::offset bounds.rs:55-61
    const double scaled = (double)kBuckets * offset;
    unsigned long long raw = 0;                                 // `as usize`: saturating, NaN -> 0
    if (scaled > 0.0) raw = scaled >= 18446744073709551616.0 ? ~0ull : (unsigned long long)scaled;
    return (uint32_t)(raw < kBuckets - 1 ? raw : kBuckets - 1);
}

// ---- phase A --------------------------------------------------------------------------------------------------------------

__global__ void k_init_scratch(Scratch* sc, const uint32_t* active, uint32_t na) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= na) return;
    Scratch& x = sc[s];
    for (int k = 0; k < 12; ++k) x.acc[k] = (k % 6) < 3 ? kKeyMinInit : kKeyMaxInit;
    for (uint32_t b = 0; b < kBuckets; ++b) {
        for (int a = 0; a < 3; ++a) { x.blo[b][a] = kKeyMinInit; x.bhi[b][a] = kKeyMaxInit; }
        x.bcount[b] = 0;
    }
    x.node = active[s];
}

// node bounds and centroid bounds of every active segment
__global__ void __launch_bounds__(256) k_accumulate(const BuildItem* __restrict__ items, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ slot_of, Scratch* sc, uint32_t n) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t slot = kNone;
    unsigned long long key[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) key[k] = (k % 6) < 3 ? kKeyMinInit : kKeyMaxInit;
    if (i < n) {
        const uint32_t s = seg[i];
        if (s != kNone) {
            slot = slot_of[s];
            const BuildItem it = items[i];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double c = centroid_of(it, a);
                key[a] = key_of(it.lo[a]); key[3 + a] = key_of(it.hi[a]);
                key[6 + a] = key_of(c); key[9 + a] = key_of(c);
            }
        }
    }
    // the usual case: every item of the warp that takes part belongs to one segment -> one set of atomics per warp
    const unsigned members = __ballot_sync(0xFFFFFFFFu, slot != kNone);
    if (members == 0u) return;
    const uint32_t first = __shfl_sync(0xFFFFFFFFu, slot, __ffs(members) - 1);
    const bool uniform = __all_sync(0xFFFFFFFFu, slot == kNone || slot == first);
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            unsigned long long v = key[k];
#pragma unroll
            for (int off = 16; off; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, off);
                v = (k % 6) < 3 ? (o < v ? o : v) : (o > v ? o : v);
            }
            if (lane == 0) {
                if ((k % 6) < 3) atomicMin(&sc[first].acc[k], v);
                else atomicMax(&sc[first].acc[k], v);
            }
        }
    } else if (slot != kNone) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            if ((k % 6) < 3) atomicMin(&sc[slot].acc[k], key[k]);
            else atomicMax(&sc[slot].acc[k], key[k]);
        }
    }
}

// bounds -> node box, split axis, centroid range (bvh.rs:239-253)
__global__ void k_params(Scratch* sc, GNode* nodes, uint32_t na, uint32_t* error) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= na) return;
    Scratch& x = sc[s];
    GNode& nd = nodes[x.node];
    double clo[3], chi[3];
    for (int a = 0; a < 3; ++a) {
        nd.lo[a] = value_of(x.acc[a]); nd.hi[a] = value_of(x.acc[3 + a]);
        clo[a] = value_of(x.acc[6 + a]); chi[a] = value_of(x.acc[9 + a]);
    }
    x.total_area = surface_area(nd.lo, nd.hi);
    if (!(x.total_area > 0.0)) atomicMax(error, (uint32_t)ERR_AREA);
    x.axis = maximum_extent(clo, chi);
    x.cmin = clo[x.axis];
    x.cext = chi[x.axis] - clo[x.axis];
}

// bucket of every item of an active segment, bucket bounds and counts (bvh.rs:255-279)
__global__ void __launch_bounds__(256) k_buckets(BuildItem* __restrict__ items, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ slot_of, Scratch* sc, uint32_t n) {
    __shared__ unsigned long long s_lo[kBuckets][3], s_hi[kBuckets][3];
    __shared__ unsigned int s_count[kBuckets];
    __shared__ uint32_t s_slot;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (threadIdx.x < kBuckets) {
        for (int a = 0; a < 3; ++a) { s_lo[threadIdx.x][a] = kKeyMinInit; s_hi[threadIdx.x][a] = kKeyMaxInit; }
        s_count[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) s_slot = kNone;
    __syncthreads();
    uint32_t slot = kNone;
    if (i < n) {
        const uint32_t s = seg[i];
        if (s != kNone) slot = slot_of[s];
    }
    // the block's accumulators serve the segment of its lowest participating item; items of other segments (at most the tail of
    // the block, where a segment ends) go to global memory directly
    if (slot != kNone) atomicMin(&s_slot, slot == kNone ? kNone : (uint32_t)threadIdx.x);
    __syncthreads();
    const uint32_t leader = s_slot;   // thread index of the first participating item, or kNone
    __shared__ uint32_t s_main;
    if (threadIdx.x == leader) s_main = slot;
    __syncthreads();
    if (leader == kNone) return;
    const uint32_t main_slot = s_main;
    if (slot != kNone) {
        const Scratch& x = sc[slot];
        BuildItem it = items[i];
        const uint32_t b = bucket_of(centroid_of(it, x.axis), x.cmin, x.cext);
        items[i].bucket = b;
        if (slot == main_slot) {
            for (int a = 0; a < 3; ++a) { atomicMin(&s_lo[b][a], key_of(it.lo[a])); atomicMax(&s_hi[b][a], key_of(it.hi[a])); }
            atomicAdd(&s_count[b], 1u);
        } else {
            Scratch& g = sc[slot];
            for (int a = 0; a < 3; ++a) { atomicMin(&g.blo[b][a], key_of(it.lo[a])); atomicMax(&g.bhi[b][a], key_of(it.hi[a])); }
            atomicAdd(&g.bcount[b], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kBuckets && s_count[threadIdx.x]) {
        Scratch& g = sc[main_slot];
        const uint32_t b = threadIdx.x;
        for (int a = 0; a < 3; ++a) { atomicMin(&g.blo[b][a], s_lo[b][a]); atomicMax(&g.bhi[b][a], s_hi[b][a]); }
        atomicAdd(&g.bcount[b], s_count[b]);
    }
}

// The cost of splitting after each bucket, the cheapest one, the size of the left part (bvh.rs:281-318).
__device__ inline uint32_t choose_split(const double (*blo)[3], const double (*bhi)[3], const unsigned int* bcount, double total_area, bool& finite, double& best_cost) {
    double costs[kBuckets - 1];
    finite = true;
    for (uint32_t i = 0; i + 1 < kBuckets; ++i) {
        double cost = kTraversalCost;
        for (int part = 0; part < 2; ++part) {
            const uint32_t lo = part == 0 ? 0 : i + 1, hi = part == 0 ? i + 1 : kBuckets;
            double mlo[3], mhi[3];
            bool some = false;
            unsigned long long count = 0;
            for (uint32_t k = lo; k < hi; ++k) {
                if (!bcount[k]) continue;
                if (!some) { for (int a = 0; a < 3; ++a) { mlo[a] = blo[k][a]; mhi[a] = bhi[k][a]; } some = true; }
                else for (int a = 0; a < 3; ++a) { mlo[a] = fmin(mlo[a], blo[k][a]); mhi[a] = fmax(mhi[a], bhi[k][a]); }
                count += bcount[k];
            }
            if (some) cost += (double)count * surface_area(mlo, mhi) / total_area;
        }
        if (!isfinite(cost)) finite = false;
        costs[i] = cost;
    }
    uint32_t best = 0;
    for (uint32_t i = 0; i + 1 < kBuckets; ++i)
        if (costs[i] < costs[best]) best = i;
    best_cost = costs[best];
    return best;
}

__global__ void k_cost(Scratch* sc, const GNode* nodes, uint32_t na, uint32_t* error) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= na) return;
    Scratch& x = sc[s];
    double blo[kBuckets][3], bhi[kBuckets][3];
    for (uint32_t b = 0; b < kBuckets; ++b)
        for (int a = 0; a < 3; ++a) { blo[b][a] = value_of(x.blo[b][a]); bhi[b][a] = value_of(x.bhi[b][a]); }
    bool finite;
    double best_cost;
    x.best = choose_split(blo, bhi, x.bcount, x.total_area, finite, best_cost);
    if (!finite) atomicMax(error, (uint32_t)ERR_COST);
    uint32_t m = 0;
    for (uint32_t b = 0; b <= x.best; ++b) m += x.bcount[b];
    x.mid = m;
    // (a segment of more than kSmall > MAX_LEAF_PRIMITIVES items never becomes a leaf, bvh.rs:314-316)
    if (m == 0 || m == nodes[x.node].count) atomicMax(error, (uint32_t)ERR_EMPTY_SIDE);
}

// partition_by (util.rs:4-26): which items sit on the wrong side
__global__ void __launch_bounds__(256) k_flags(const BuildItem* __restrict__ items, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ slot_of, const Scratch* __restrict__ sc,
                                                const GNode* __restrict__ nodes, uint32_t* f_left, uint32_t* f_right, uint32_t n) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i > n) return;
    uint32_t fl = 0, fr = 0;
    if (i < n) {
        const uint32_t s = seg[i];
        if (s != kNone) {
            const Scratch& x = sc[slot_of[s]];
            const bool pass = items[i].bucket <= x.best;
            const uint32_t pos = i - nodes[s].begin;
            fl = pos < x.mid && !pass;
            fr = pos >= x.mid && pass;
        }
    }
    f_left[i] = fl; f_right[i] = fr;   // (entry n stays 0: the scans' entry n is then the total)
}
__global__ void __launch_bounds__(256) k_collect(const uint32_t* __restrict__ f_left, const uint32_t* __restrict__ f_right, const uint32_t* __restrict__ s_left, const uint32_t* __restrict__ s_right,
                                                  uint32_t* pos_left, uint32_t* pos_right, uint32_t n) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    if (f_left[i]) pos_left[s_left[i]] = i;
    if (f_right[i]) pos_right[s_right[i]] = i;
}
// the k-th misplaced item of a segment's left part (ascending) changes places with the k-th misplaced item of its right part
// counted from the back
__global__ void __launch_bounds__(256) k_swap(BuildItem* items, const uint32_t* __restrict__ seg, const GNode* __restrict__ nodes, const uint32_t* __restrict__ s_left, const uint32_t* __restrict__ s_right,
                                               const uint32_t* __restrict__ pos_left, const uint32_t* __restrict__ pos_right, uint32_t n, uint32_t* error) {
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= s_left[n]) return;
    const uint32_t i = pos_left[j];
    const GNode& nd = nodes[seg[i]];
    const uint32_t base_l = s_left[nd.begin], k_pairs = s_left[nd.begin + nd.count] - base_l;
    const uint32_t base_r = s_right[nd.begin];
    if (s_right[nd.begin + nd.count] - base_r != k_pairs) { atomicMax(error, (uint32_t)ERR_PAIRS); return; }
    const uint32_t partner = pos_right[base_r + (k_pairs - 1u - (j - base_l))];
    const BuildItem a = items[i], b = items[partner];
    items[i] = b; items[partner] = a;
}

// the two children of every segment split at this level
__global__ void k_children(const Scratch* __restrict__ sc, GNode* nodes, uint32_t na, uint32_t* node_count, uint32_t node_capacity, uint32_t* next_active, uint32_t* n_next, uint32_t* slot_of,
                           uint32_t* small_list, uint32_t* n_small, uint32_t* error) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= na) return;
    const Scratch& x = sc[s];
    GNode& nd = nodes[x.node];
    const uint32_t base = atomicAdd(node_count, 2u);
    if (base + 2u > node_capacity) { atomicMax(error, (uint32_t)ERR_CAPACITY); return; }
    nd.left = base; nd.right = base + 1u; nd.axis = (uint32_t)x.axis; nd.kind = KIND_SPLIT;
    for (int side = 0; side < 2; ++side) {
        GNode c{};
        c.begin = side == 0 ? nd.begin : nd.begin + x.mid;
        c.count = side == 0 ? x.mid : nd.count - x.mid;
        c.left = c.right = kNone; c.axis = 3;
        const uint32_t id = base + side;
        if (c.count > kSmall) {
            c.kind = KIND_ACTIVE;
            const uint32_t slot = atomicAdd(n_next, 1u);
            next_active[slot] = id;
            slot_of[id] = slot;
        } else {
            c.kind = KIND_SMALL;
            small_list[atomicAdd(n_small, 1u)] = id;
        }
        nodes[id] = c;
    }
}
// every item follows its segment into the child that now holds it (or leaves phase A with it)
__global__ void __launch_bounds__(256) k_reassign(uint32_t* seg, const GNode* __restrict__ nodes, uint32_t n) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = seg[i];
    if (s == kNone) return;
    const GNode& nd = nodes[s];
    const uint32_t child = (i - nd.begin) < nodes[nd.left].count ? nd.left : nd.right;
    seg[i] = nodes[child].kind == KIND_ACTIVE ? child : kNone;
}

// ---- phase B: the reference's sequential recursion on one small segment per thread ------------------------------------------

__global__ void __launch_bounds__(64) k_small_subtrees(BuildItem* items, GNode* nodes, const uint32_t* __restrict__ small_list, uint32_t n_small, BinNode* local_nodes, uint32_t* error) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_small) return;
    GNode& top = nodes[small_list[t]];
    BinNode* out = local_nodes + 2ull * top.begin;   // a sub-tree over c items has at most 2c - 1 nodes
    struct Frame { uint32_t begin, n, parent, side; };
    Frame stack[kSmall + 2];   // depth-first, left child first: at most one pending right child per level
    int sp = 0;
    stack[sp++] = {top.begin, top.count, kNone, 0u};
    uint32_t emitted = 0;
    while (sp > 0) {
        const Frame f = stack[--sp];
        const uint32_t me = emitted++;
        if (f.parent != kNone) {
            if (f.side == 0) out[f.parent].a = me;
            else { out[f.parent].b = me; out[f.parent].right_first = f.begin; }
        }
        BuildItem* its = items + f.begin;
        double lo[3], hi[3], clo[3], chi[3];
        for (uint32_t i = 0; i < f.n; ++i) {
            for (int a = 0; a < 3; ++a) {
                const double c = centroid_of(its[i], a);
                if (i == 0) { lo[a] = its[i].lo[a]; hi[a] = its[i].hi[a]; clo[a] = c; chi[a] = c; }
                else { lo[a] = fmin(lo[a], its[i].lo[a]); hi[a] = fmax(hi[a], its[i].hi[a]); clo[a] = fmin(clo[a], c); chi[a] = fmax(chi[a], c); }
            }
        }
        BinNode nd{};
        nd.box.lo = mk(lo[0], lo[1], lo[2]); nd.box.hi = mk(hi[0], hi[1], hi[2]);
        auto leaf = [&]() { nd.a = f.begin; nd.b = f.n; nd.axis = 3u; nd.right_first = 0u; out[me] = nd; };
        if (f.n <= 1) { leaf(); continue; }
        const double total_area = surface_area(lo, hi);
        if (!(total_area > 0.0)) { atomicMax(error, (uint32_t)ERR_AREA); leaf(); continue; }
        const int axis = maximum_extent(clo, chi);
        const double cmin = clo[axis], cext = chi[axis] - clo[axis];
        double blo[kBuckets][3], bhi[kBuckets][3];
        unsigned int bcount[kBuckets];
        for (uint32_t b = 0; b < kBuckets; ++b) bcount[b] = 0;
        for (uint32_t i = 0; i < f.n; ++i) {
            const uint32_t b = bucket_of(centroid_of(its[i], axis), cmin, cext);
            its[i].bucket = b;
            if (!bcount[b]) { for (int a = 0; a < 3; ++a) { blo[b][a] = its[i].lo[a]; bhi[b][a] = its[i].hi[a]; } }
            else for (int a = 0; a < 3; ++a) { blo[b][a] = fmin(blo[b][a], its[i].lo[a]); bhi[b][a] = fmax(bhi[b][a], its[i].hi[a]); }
            bcount[b] += 1;
        }
        bool finite;
        double best_cost;
        const uint32_t best = choose_split(blo, bhi, bcount, total_area, finite, best_cost);
        if (!finite) atomicMax(error, (uint32_t)ERR_COST);
        if ((double)f.n <= best_cost && f.n <= kMaxLeaf) { leaf(); continue; }
        // partition_by util.rs:4-26 with pred = bucket <= best
        uint32_t left = 0, right = f.n - 1;
        while (left != right) {
            while (left < right && its[left].bucket <= best) left += 1;
            while (right > left && !(its[right].bucket <= best)) right -= 1;
            const BuildItem tmp = its[left]; its[left] = its[right]; its[right] = tmp;
        }
        const uint32_t mid = its[left].bucket <= best ? left + 1 : left;
        if (mid == 0 || mid == f.n) { atomicMax(error, (uint32_t)ERR_EMPTY_SIDE); leaf(); continue; }
        nd.axis = (uint32_t)axis;
        out[me] = nd;
        stack[sp++] = {f.begin + mid, f.n - mid, me, 1u};   // popped after the whole left sub-tree: pre-order
        stack[sp++] = {f.begin, mid, me, 0u};
    }
    top.left = emitted;
}

// ---- assembly -----------------------------------------------------------------------------------------------------------

// `pre[id]`: pre-order index of top node `id` in the final array
__global__ void k_emit(const GNode* __restrict__ nodes, const uint32_t* __restrict__ pre, uint32_t n_top, const BinNode* __restrict__ local_nodes, BinNode* out) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_top) return;
    const GNode& nd = nodes[id];
    const uint32_t at = pre[id];
    if (nd.kind == KIND_SPLIT) {
        BinNode b{};
        b.box.lo = mk(nd.lo[0], nd.lo[1], nd.lo[2]); b.box.hi = mk(nd.hi[0], nd.hi[1], nd.hi[2]);
        b.a = pre[nd.left]; b.b = pre[nd.right]; b.axis = nd.axis; b.right_first = nodes[nd.right].begin;
        out[at] = b;
    } else if (nd.kind == KIND_SMALL) {
        const BinNode* src = local_nodes + 2ull * nd.begin;
        for (uint32_t j = 0; j < nd.left; ++j) {
            BinNode b = src[j];
            if (b.axis != 3u) { b.a += at; b.b += at; }   // (leaf ranges and right_first are absolute ranks already)
            out[at + j] = b;
        }
    }
}
__global__ void k_prim_order(const BuildItem* __restrict__ items, uint32_t* order, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[i] = items[i].prim;
}

struct DeviceBuffers {   // freed on every return path
    std::vector<void*> ptrs;
    template <class T>
    bool alloc(T** p, size_t count) {
        void* q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return false; }
        ptrs.push_back(q);
        *p = static_cast<T*>(q);
        return true;
    }
    ~DeviceBuffers() { for (void* q : ptrs) cudaFree(q); }
};

std::mutex g_release_mutex;
std::thread g_release_thread;

}  // namespace

void finish_device_build_release() {
    std::lock_guard<std::mutex> lock(g_release_mutex);
    if (g_release_thread.joinable()) g_release_thread.join();
}

bool build_reference_bvh_gpu(const BuildItem* host_items, size_t n_items, int device, RefBvh& out, std::string& why) {
    if (n_items < 2 * kSmall || n_items > 0x7FFFFFFFull) { why = "too few primitives for the device build"; return false; }
    PhaseTimer timer;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); why = "no CUDA device"; return false; }
    int previous = 0;
    cudaGetDevice(&previous);
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); why = "cudaSetDevice failed"; return false; }
    struct Restore { int dev; ~Restore() { cudaSetDevice(dev); } } restore{previous};
    cudaFree(nullptr);   // (creates the device context if this is the process's first CUDA work)
    timer.mark("gpu build: device context");
    const uint32_t n = (uint32_t)n_items;
    // the host arrays the tree is downloaded into are sized (and their pages touched) while the device works
    struct Toucher {
        std::thread th;
        ~Toucher() { if (th.joinable()) th.join(); }
    } toucher;
    toucher.th = std::thread([&out, n] { out.nodes.resize(2ull * n); out.prim_order.resize(n); });
    const uint32_t node_capacity = (uint32_t)std::min<uint64_t>(2ull * n, (uint64_t)n / 16 + 4096);   // top of the tree only
    DeviceBuffers buf;
    BuildItem* d_items; uint32_t *d_seg, *d_slot, *d_active, *d_next, *d_small, *d_fl, *d_fr, *d_sl, *d_sr, *d_pl, *d_pr, *d_counters, *d_pre, *d_order;
    GNode* d_nodes; Scratch* d_scratch; BinNode *d_local, *d_out;
    void* d_temp = nullptr;
    size_t temp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n + 1));
    const uint32_t max_active = n / kSmall + 2;
    bool ok = buf.alloc(&d_items, n) && buf.alloc(&d_seg, n) && buf.alloc(&d_slot, node_capacity) && buf.alloc(&d_active, max_active) && buf.alloc(&d_next, max_active) &&
              buf.alloc(&d_small, node_capacity) && buf.alloc(&d_fl, n + 1) && buf.alloc(&d_fr, n + 1) && buf.alloc(&d_sl, n + 1) && buf.alloc(&d_sr, n + 1) && buf.alloc(&d_pl, n) &&
              buf.alloc(&d_pr, n) && buf.alloc(&d_counters, 8) && buf.alloc(&d_nodes, node_capacity) && buf.alloc(&d_scratch, max_active) && buf.alloc(&d_local, 2ull * n) &&
              buf.alloc(&d_order, n) && buf.alloc((char**)&d_temp, temp_bytes);
    if (!ok) { why = "out of device memory for the BVH build"; return false; }
    auto check = [&](const char* what) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { why = std::string(what) + ": " + cudaGetErrorString(e); return false; }
        return true;
    };
    // counters: [0] nodes, [1] next level's segments, [2] small segments, [3] error
    cudaMemcpy(d_items, host_items, sizeof(BuildItem) * n, cudaMemcpyHostToDevice);
    cudaMemset(d_seg, 0, sizeof(uint32_t) * n);            // every item starts in the root segment (node 0)
    {
        GNode root{};
        root.begin = 0; root.count = n; root.left = root.right = kNone; root.axis = 3; root.kind = KIND_ACTIVE;
        const uint32_t zero = 0, init[8] = {1, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpy(d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice);
        cudaMemcpy(d_active, &zero, 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_slot, &zero, 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_counters, init, sizeof(init), cudaMemcpyHostToDevice);
    }
    if (!check("upload")) return false;
    timer.mark("gpu build: upload");
    const unsigned item_blocks = (n + 255) / 256, item_blocks1 = (n + 1 + 255) / 256;
    uint32_t na = 1, levels = 0;
    while (na > 0) {
        if (++levels > 4096 || na > max_active) { why = "device BVH build: level / segment limit"; return false; }
        const unsigned seg_blocks = (na + 127) / 128;
        k_init_scratch<<<seg_blocks, 128>>>(d_scratch, d_active, na);
        k_accumulate<<<item_blocks, 256>>>(d_items, d_seg, d_slot, d_scratch, n);
        k_params<<<seg_blocks, 128>>>(d_scratch, d_nodes, na, d_counters + 3);
        k_buckets<<<item_blocks, 256>>>(d_items, d_seg, d_slot, d_scratch, n);
        k_cost<<<seg_blocks, 128>>>(d_scratch, d_nodes, na, d_counters + 3);
        k_flags<<<item_blocks1, 256>>>(d_items, d_seg, d_slot, d_scratch, d_nodes, d_fl, d_fr, n);
        cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_fl, d_sl, (int)(n + 1));
        cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_fr, d_sr, (int)(n + 1));
        k_collect<<<item_blocks, 256>>>(d_fl, d_fr, d_sl, d_sr, d_pl, d_pr, n);
        k_swap<<<item_blocks, 256>>>(d_items, d_seg, d_nodes, d_sl, d_sr, d_pl, d_pr, n, d_counters + 3);
        cudaMemsetAsync(d_counters + 1, 0, 4);
        k_children<<<seg_blocks, 128>>>(d_scratch, d_nodes, na, d_counters, node_capacity, d_next, d_counters + 1, d_slot, d_small, d_counters + 2, d_counters + 3);
        k_reassign<<<item_blocks, 256>>>(d_seg, d_nodes, n);
        uint32_t h[4] = {};
        cudaMemcpy(h, d_counters, sizeof(h), cudaMemcpyDeviceToHost);
        if (!check("level")) return false;
        if (h[3] != ERR_NONE) { why = "device BVH build: the reference's builder would fail here (code " + std::to_string(h[3]) + ")"; return false; }
        na = h[1];
        std::swap(d_active, d_next);
    }
    timer.mark("gpu build: top levels");
    uint32_t h[4] = {};
    cudaMemcpy(h, d_counters, sizeof(h), cudaMemcpyDeviceToHost);
    const uint32_t n_top = h[0], n_small = h[2];
    k_small_subtrees<<<(n_small + 63) / 64, 64>>>(d_items, d_nodes, d_small, n_small, d_local, d_counters + 3);
    std::vector<GNode> top(n_top);
    cudaMemcpy(top.data(), d_nodes, sizeof(GNode) * n_top, cudaMemcpyDeviceToHost);
    cudaMemcpy(h, d_counters, sizeof(h), cudaMemcpyDeviceToHost);
    if (!check("sub-trees")) return false;
    if (h[3] != ERR_NONE) { why = "device BVH build: the reference's builder would fail here (code " + std::to_string(h[3]) + ")"; return false; }
    timer.mark("gpu build: sub-trees");
    // pre-order numbering of the top of the tree on the host (it is small): sizes bottom-up, offsets top-down
    std::vector<uint32_t> size(n_top, 0), pre(n_top, 0), order;
    order.reserve(n_top);
    {
        std::vector<uint32_t> stack{0u};
        while (!stack.empty()) {   // a pre-order walk; its reverse visits children before parents
            const uint32_t id = stack.back();
            stack.pop_back();
            order.push_back(id);
            if (top[id].kind == KIND_SPLIT) { stack.push_back(top[id].right); stack.push_back(top[id].left); }
        }
        for (size_t k = order.size(); k-- > 0;) {
            const GNode& g = top[order[k]];
            size[order[k]] = g.kind == KIND_SPLIT ? 1u + size[g.left] + size[g.right] : g.left;
        }
        pre[0] = 0;
        for (uint32_t id : order)
            if (top[id].kind == KIND_SPLIT) { pre[top[id].left] = pre[id] + 1u; pre[top[id].right] = pre[id] + 1u + size[top[id].left]; }
    }
    const uint32_t total_nodes = size[0];
    if (!buf.alloc(&d_pre, n_top) || !buf.alloc(&d_out, total_nodes)) { why = "out of device memory for the BVH build"; return false; }
    cudaMemcpy(d_pre, pre.data(), sizeof(uint32_t) * n_top, cudaMemcpyHostToDevice);
    k_emit<<<(n_top + 127) / 128, 128>>>(d_nodes, d_pre, n_top, d_local, d_out);
    k_prim_order<<<item_blocks, 256>>>(d_items, d_order, n);
    toucher.th.join();
    out.nodes.resize(total_nodes);   // (shrinks: a tree over n items has at most 2n - 1 nodes)
    cudaMemcpy(out.nodes.data(), d_out, sizeof(BinNode) * total_nodes, cudaMemcpyDeviceToHost);
    cudaMemcpy(out.prim_order.data(), d_order, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
    if (!check("assembly")) return false;
    timer.mark("gpu build: assembly + download");
    // the ~6 GB of build buffers are released behind the caller's back (cudaFree of the first build of a process was measured at
    // 0.85 s): finish_device_build_release() waits for it
    finish_device_build_release();
    {
        std::vector<void*> ptrs;
        ptrs.swap(buf.ptrs);
        std::lock_guard<std::mutex> lock(g_release_mutex);
        g_release_thread = std::thread([ptrs, device] {
            if (cudaSetDevice(device) == cudaSuccess)
                for (void* q : ptrs) cudaFree(q);
            cudaGetLastError();
        });
    }
    return true;
}

}  // namespace cray
