// See bvh_build.hpp.
#include "bvh_build.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "host_math.hpp"

namespace cray {

namespace {

constexpr uint32_t kBuckets = 12;            // bvh.rs:235
constexpr double kTraversalCost = 1.0 / 8.0; // bvh.rs:236
constexpr size_t kMaxLeaf = 4;               // bvh.rs:237
constexpr uint32_t kDeferred = 0xFFFFFFFFu;

using Item = BuildItem;

inline Box3 item_box(const Item& it) { return {mk(it.lo[0], it.lo[1], it.lo[2]), mk(it.hi[0], it.hi[1], it.hi[2])}; }
inline double centroid_axis(const Item& it, int axis) { return (it.lo[axis] + it.hi[axis]) * 0.5; }

struct Accum {  // running Bounds sum, bounds.rs:91-108 / :129-137
    double lo[3], hi[3];
    bool some = false;
    void add(const double* l, const double* h) {
        if (!some) { std::memcpy(lo, l, 24); std::memcpy(hi, h, 24); some = true; return; }
        for (int k = 0; k < 3; ++k) { lo[k] = rmin(lo[k], l[k]); hi[k] = rmax(hi[k], h[k]); }
    }
    Box3 box() const { return {mk(lo[0], lo[1], lo[2]), mk(hi[0], hi[1], hi[2])}; }
};

struct SubTree {
    std::vector<BinNode> nodes;        // local pre-order; deferred placeholders have axis == kDeferred, a = task id
    std::vector<uint32_t> order;
    std::string error;
};
struct Task {
    Item* items;
    size_t n;
    SubTree tree;
};

// Runs fn(chunk, begin, end) over [0, n) split into `threads` contiguous chunks, one thread each.
template <class F>
void parallel_chunks(size_t n, unsigned threads, F fn) {
    threads = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, n));
    const size_t step = (n + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (unsigned c = 1; c < threads; ++c) {
        const size_t b = std::min(n, c * step), e = std::min(n, b + step);
        pool.emplace_back([=] { fn(c, b, e); });
    }
    fn(0u, (size_t)0, std::min(n, step));
    for (auto& th : pool) th.join();
}

constexpr size_t kParallelNode = 1u << 15;  // nodes with at least this many primitives are split by all threads

struct Builder {
    size_t grain;                 // subtrees at or below this size are deferred to the pool (0 = never defer)
    std::vector<Task>* tasks;
    unsigned threads = 1;         // > 1: the passes over the items of a large node run on this many threads
    bool all_axes = false;        // false: the reference's rule (split along the longest centroid axis, bvh.rs:253); true: the cheapest
                                  // of the three axes' best splits (build_quality_bvh: the tree the wide BVH is collapsed from)

    uint32_t leaf(SubTree& t, const Item* items, size_t n, const Box3& box) {
        uint32_t idx = (uint32_t)t.nodes.size();
        t.nodes.push_back({box, (uint32_t)t.order.size(), (uint32_t)n, 3u, 0u});
        for (size_t i = 0; i < n; ++i) t.order.push_back(items[i].prim);
        return idx;
    }

    // BvhNode::from_sah_splitting bvh.rs:234-336
    uint32_t split(SubTree& t, Item* items, size_t n) {
        if (grain && n <= grain && tasks) {
            uint32_t idx = (uint32_t)t.nodes.size();
            tasks->push_back({items, n, {}});
            t.nodes.push_back({Box3{}, (uint32_t)tasks->size() - 1, 0u, kDeferred, 0u});
            return idx;
        }
        // Every reduction of this function is a min / max or a count, so chunking it over threads changes no bit of the tree.
        const bool parallel = threads > 1 && n >= kParallelNode;
        Accum all, cent;
        auto accumulate = [&](Accum& a, Accum& c3, size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                a.add(items[i].lo, items[i].hi);
                double c[3] = {centroid_axis(items[i], 0), centroid_axis(items[i], 1), centroid_axis(items[i], 2)};
                c3.add(c, c);
            }
        };
        if (parallel) {
            std::vector<Accum> pa(threads), pc(threads);
            parallel_chunks(n, threads, [&](unsigned c, size_t b, size_t e) {
                Accum a, c3;  // thread-local: neighbouring slots of pa / pc share cache lines
                accumulate(a, c3, b, e);
                pa[c] = a; pc[c] = c3;
            });
            for (unsigned c = 0; c < threads; ++c) {
                if (pa[c].some) all.add(pa[c].lo, pa[c].hi);
                if (pc[c].some) cent.add(pc[c].lo, pc[c].hi);
            }
        } else {
            accumulate(all, cent, 0, n);
        }
        const Box3 bounds = all.box();
        if (n <= 1) return leaf(t, items, n, bounds);

        const double total_surface_area = box_surface_area(bounds);
        if (!(total_surface_area > 0.0)) {
            if (t.error.empty()) t.error = "Encountered primitives with no surface area";
            return leaf(t, items, n, bounds);
        }
        const Box3 cb = cent.box();
        int axis = box_maximum_extent(cb);
        double cmin = cb.lo[axis], cext = cb.hi[axis] - cb.lo[axis];

        struct Buckets {
            Accum box[kBuckets];
            size_t count[kBuckets] = {};
        };
        Buckets bk;
        auto fill = [&](Buckets& dst, size_t b0, size_t e0) {
            for (size_t i = b0; i < e0; ++i) {
                const double offset = (centroid_axis(items[i], axis) - cmin) / cext;                 // Bounds::offset bounds.rs:55-61
                const uint64_t raw = as_usize((double)kBuckets * offset);                             // `as usize`, NaN -> 0
                const uint32_t b = (uint32_t)std::min<uint64_t>(raw, kBuckets - 1);
                items[i].bucket = b;
                dst.box[b].add(items[i].lo, items[i].hi);
                dst.count[b] += 1;
            }
        };
        auto fill_all = [&]() {
            bk = Buckets{};
            if (parallel) {
                std::vector<Buckets> part(threads);
                parallel_chunks(n, threads, [&](unsigned c, size_t b0, size_t e0) {
                    Buckets local;
                    fill(local, b0, e0);
                    part[c] = local;
                });
                for (unsigned c = 0; c < threads; ++c)
                    for (uint32_t b = 0; b < kBuckets; ++b) {
                        if (part[c].box[b].some) bk.box[b].add(part[c].box[b].lo, part[c].box[b].hi);
                        bk.count[b] += part[c].count[b];
                    }
            } else {
                fill(bk, 0, n);
            }
        };
        double costs[kBuckets - 1];
        uint32_t best = 0;
        auto evaluate = [&]() {
            Accum* bucket_box = bk.box;
            size_t* bucket_count = bk.count;
            for (uint32_t i = 0; i + 1 < kBuckets; ++i) {
                double cost = kTraversalCost;
                for (int part = 0; part < 2; ++part) {
                    const uint32_t lo = part == 0 ? 0 : i + 1, hi = part == 0 ? i + 1 : kBuckets;
                    Accum merged;
                    size_t count = 0;
                    for (uint32_t k = lo; k < hi; ++k)
                        if (bucket_box[k].some) { merged.add(bucket_box[k].lo, bucket_box[k].hi); count += bucket_count[k]; }
                    if (merged.some) cost += (double)count * box_surface_area(merged.box()) / total_surface_area;
                }
                if (!std::isfinite(cost) && t.error.empty()) t.error = "SAH cost is not finite";
                costs[i] = cost;
            }
            best = 0;
            for (uint32_t i = 0; i + 1 < kBuckets; ++i)
                if (costs[i] < costs[best]) best = i;
        };
        if (all_axes) {
            // the cheapest split over the three axes (an axis along which every centroid coincides cannot split)
            int best_axis = -1;
            double best_cost = 0.0;
            for (int a = 0; a < 3; ++a) {
                if (!(cb.hi[a] - cb.lo[a] > 0.0)) continue;
                axis = a; cmin = cb.lo[a]; cext = cb.hi[a] - cb.lo[a];
                fill_all();
                evaluate();
                size_t left = 0;
                for (uint32_t b = 0; b <= best; ++b) left += bk.count[b];
                if (left == 0 || left == n) continue;
                if (best_axis < 0 || costs[best] < best_cost) { best_axis = a; best_cost = costs[best]; }
            }
            axis = best_axis < 0 ? box_maximum_extent(cb) : best_axis;
            cmin = cb.lo[axis]; cext = cb.hi[axis] - cb.lo[axis];
        }
        fill_all();
        evaluate();
        size_t* bucket_count = bk.count;

        if ((double)n <= costs[best] && n <= kMaxLeaf) return leaf(t, items, n, bounds);

        // partition_by util.rs:4-26 with pred = bucket <= best
        size_t mid;
        if (parallel) {
            // The sequential two-pointer loop never moves an item that already sits on its side; it swaps the k-th item that
            // fails the predicate inside the first m positions (ascending) with the k-th item that passes it behind them
            // (descending), m = number of passing items.  The same pairs, found with per-chunk counts and swapped in parallel:
            size_t m = 0;
            for (uint32_t b = 0; b <= best; ++b) m += bucket_count[b];
            std::vector<size_t> nf(threads + 1, 0), nt(threads + 1, 0);
            parallel_chunks(n, threads, [&](unsigned c, size_t b0, size_t e0) {
                size_t f = 0, tr = 0;
                for (size_t i = b0; i < e0; ++i) {
                    const bool pass = items[i].bucket <= best;
                    if (i < m) f += !pass;
                    else tr += pass;
                }
                nf[c + 1] = f; nt[c + 1] = tr;
            });
            for (unsigned c = 0; c < threads; ++c) { nf[c + 1] += nf[c]; nt[c + 1] += nt[c]; }
            const size_t k = nf[threads];
            std::vector<uint32_t> fpos(k), tpos(k);
            parallel_chunks(n, threads, [&](unsigned c, size_t b0, size_t e0) {
                size_t f = nf[c], tr = nt[c];
                for (size_t i = b0; i < e0; ++i) {
                    const bool pass = items[i].bucket <= best;
                    if (i < m) { if (!pass) fpos[f++] = (uint32_t)i; }
                    else if (pass) tpos[tr++] = (uint32_t)i;
                }
            });
            if (nt[threads] == k)
                parallel_chunks(k, threads, [&](unsigned, size_t b0, size_t e0) {
                    for (size_t j = b0; j < e0; ++j) std::swap(items[fpos[j]], items[tpos[k - 1 - j]]);
                });
            mid = m;
        } else {
            size_t left = 0, right = n - 1;
            while (left != right) {
                while (left < right && items[left].bucket <= best) left += 1;
                while (right > left && !(items[right].bucket <= best)) right -= 1;
                std::swap(items[left], items[right]);
            }
            mid = items[left].bucket <= best ? left + 1 : left;
        }
        if (mid == 0 || mid == n) {
            if (t.error.empty()) t.error = "SAH split left one side empty (the reference asserts, bvh.rs:327-328)";
            return leaf(t, items, n, bounds);
        }
        const uint32_t idx = (uint32_t)t.nodes.size();
        t.nodes.push_back({bounds, 0u, 0u, (uint32_t)axis, 0u});
        const uint32_t l = split(t, items, mid);
        const uint32_t right_first = (uint32_t)t.order.size();
        const uint32_t r = split(t, items + mid, n - mid);
        t.nodes[idx].a = l;
        t.nodes[idx].b = r;
        t.nodes[idx].right_first = right_first;
        return idx;
    }
};

// Splice `src` (local indices / ranks) into the final pre-order arrays: one depth-first pass over the (small) top of the tree fixes
// where every node of it and every deferred sub-tree lands, then the sub-trees are copied in by all threads.
void splice(const SubTree& src, const std::vector<Task>& tasks, RefBvh& out, unsigned threads) {
    struct Frame { uint32_t local; uint32_t parent_final; int side; };
    struct Placement { uint32_t task, node_off, rank_off; };
    std::vector<Placement> placements;
    size_t total_nodes = src.nodes.size(), total_ranks = src.order.size();
    for (const BinNode& n : src.nodes)
        if (n.axis == kDeferred) { total_nodes += tasks[n.a].tree.nodes.size() - 1; total_ranks += tasks[n.a].tree.order.size(); }
    out.nodes.resize(total_nodes);
    out.prim_order.resize(total_ranks);
    uint32_t node_off = 0, rank_off = 0;
    std::vector<Frame> stack;
    stack.push_back({0, 0xFFFFFFFFu, 0});
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        const BinNode& n = src.nodes[f.local];
        const uint32_t me = node_off;
        if (f.parent_final != 0xFFFFFFFFu) {
            if (f.side == 0) out.nodes[f.parent_final].a = me;
            else { out.nodes[f.parent_final].b = me; out.nodes[f.parent_final].right_first = rank_off; }
        }
        if (n.axis == kDeferred) {
            const SubTree& sub = tasks[n.a].tree;
            placements.push_back({n.a, node_off, rank_off});
            node_off += (uint32_t)sub.nodes.size();
            rank_off += (uint32_t)sub.order.size();
            if (out.error.empty() && !sub.error.empty()) out.error = sub.error;
        } else if (n.axis == 3) {
            BinNode c = n;
            c.a = rank_off;
            out.nodes[me] = c;
            for (uint32_t i = 0; i < n.b; ++i) out.prim_order[rank_off + i] = src.order[n.a + i];
            node_off += 1;
            rank_off += n.b;
        } else {
            out.nodes[me] = n;
            node_off += 1;
            // pre-order: left subtree is emitted completely before the right one => push right first
            stack.push_back({n.b, me, 1});
            stack.push_back({n.a, me, 0});
        }
    }
    parallel_chunks(placements.size(), threads, [&](unsigned, size_t b0, size_t e0) {
        for (size_t k = b0; k < e0; ++k) {
            const Placement& pl = placements[k];
            const SubTree& sub = tasks[pl.task].tree;
            for (size_t j = 0; j < sub.nodes.size(); ++j) {
                BinNode c = sub.nodes[j];
                if (c.axis == 3) c.a += pl.rank_off;
                else { c.a += pl.node_off; c.b += pl.node_off; c.right_first += pl.rank_off; }
                out.nodes[pl.node_off + j] = c;
            }
            std::copy(sub.order.begin(), sub.order.end(), out.prim_order.begin() + pl.rank_off);
        }
    });
}

}  // namespace

Box3 primitive_bounds(const cray_scene_desc& d, uint64_t prim) {  // Shape::bounds shape.rs:402-438
    const cray_primitive_desc& p = d.primitives[prim];
    switch (p.shape_kind) {
        case CRAY_SHAPE_SPHERE: {
            const cray_sphere_desc& s = d.spheres[p.shape_index];
            const double r = s.radius;
            const Xform o2w = xf_translate(s.origin[0], s.origin[1], s.origin[2]);
            // Bounds::new takes the component-wise min/max of its corners (bounds.rs:16-21)
            Box3 local{mk(rmin(-r, r), rmin(-r, r), rmin(-r, r)), mk(rmax(-r, r), rmax(-r, r), rmax(-r, r))};
            return transform_box(o2w.fwd, local);
        }
        case CRAY_SHAPE_TRIANGLE: {
            const cray_triangle_desc& t = d.triangles[p.shape_index];
            const V3 v0 = mk(t.v0[0], t.v0[1], t.v0[2]);
            const V3 v1 = v0 + mk(t.e1[0], t.e1[1], t.e1[2]);
            const V3 v2 = v0 + mk(t.e2[0], t.e2[1], t.e2[2]);
            return {mk(rmin(v1.x, rmin(v2.x, v0.x)), rmin(v1.y, rmin(v2.y, v0.y)), rmin(v1.z, rmin(v2.z, v0.z))),
                    mk(rmax(v1.x, rmax(v2.x, v0.x)), rmax(v1.y, rmax(v2.y, v0.y)), rmax(v1.z, rmax(v2.z, v0.z)))};
        }
        default: {
            const cray_disk_desc& k = d.disks[p.shape_index];
            const double r = k.radius;
            const Xform o2w = xf_translate(k.origin[0], k.origin[1], k.origin[2]) * xf_rotate(0, to_radians(k.rotate_x)) * xf_rotate(1, to_radians(k.rotate_y));
            Box3 local{mk(rmin(-r, r), rmin(-r, r), 0.0), mk(rmax(-r, r), rmax(-r, r), 0.0)};
            return transform_box(o2w.fwd, local);
        }
    }
}

namespace {
void build_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads, bool all_axes, int gpu_device);
thread_local bool t_last_build_on_device = false;
}
bool last_reference_build_was_on_device() { return t_last_build_on_device; }
void build_reference_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads, int gpu_device) { build_bvh(d, out, threads, false, gpu_device); }
void build_quality_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads) { build_bvh(d, out, threads, true, -1); }
namespace {
void build_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads, bool all_axes, int gpu_device) {
    PhaseTimer timer;
    const size_t n = (size_t)d.n_primitives;
    out = RefBvh{};
    if (!all_axes) t_last_build_on_device = false;
    if (n == 0) { out.error = "no primitives"; return; }
    if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<Item> items(n);
    {
        // primitive bounds in parallel (pure function of the description)
        std::vector<std::thread> pool;
        const size_t chunk = (n + threads - 1) / threads;
        auto work = [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) {
                Box3 bx = primitive_bounds(d, i);
                items[i] = {{bx.lo.x, bx.lo.y, bx.lo.z}, {bx.hi.x, bx.hi.y, bx.hi.z}, (uint32_t)i, 0u};
            }
        };
        for (unsigned t = 1; t < threads; ++t) {
            size_t b = t * chunk, e = std::min(n, b + chunk);
            if (b < e) pool.emplace_back(work, b, e);
        }
        work(0, std::min(n, chunk));
        for (auto& th : pool) th.join();
    }
    timer.mark("primitive bounds");
    Accum all;
    {   // (a min / max over all items: chunked over the threads, no bit depends on the chunking)
        std::vector<Accum> part(threads);
        parallel_chunks(n, threads, [&](unsigned c, size_t b0, size_t e0) {
            Accum local;
            for (size_t i = b0; i < e0; ++i) local.add(items[i].lo, items[i].hi);
            part[c] = local;
        });
        for (unsigned c = 0; c < threads; ++c)
            if (part[c].some) all.add(part[c].lo, part[c].hi);
    }
    out.bounds = all.box();  // Bvh::bounds bvh.rs:53
    timer.mark("scene bounds");
    if (!all_axes && gpu_device >= 0 && n > (1u << 16)) {
        const char* e = std::getenv("CRAY_GPU_BUILD");
        if (!e || std::atoi(e) != 0) {
            std::string why;
            if (build_reference_bvh_gpu(items.data(), n, gpu_device, out, why)) {
                timer.mark("tree on the device");
                t_last_build_on_device = true;
                return;
            }
            if (timer.on) std::fprintf(stderr, "[cray build] device build not used: %s\n", why.c_str());
            out.nodes.clear(); out.prim_order.clear(); out.error.clear();
        }
    }

    std::vector<Task> tasks;
    SubTree top;
    Builder top_builder{(threads > 1 && n > (1u << 16)) ? std::max<size_t>(1u << 15, n / (threads * 8)) : 0, &tasks, threads, all_axes};
    top_builder.split(top, items.data(), n);
    timer.mark("top of the tree");
    if (!tasks.empty()) {
        std::atomic<size_t> next{0};
        auto run = [&]() {
            Builder b{0, nullptr, 1, all_axes};
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= tasks.size()) break;
                b.split(tasks[i].tree, tasks[i].items, tasks[i].n);
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(run);
        run();
        for (auto& th : pool) th.join();
    }
    timer.mark("sub-trees");
    out.error = top.error;
    splice(top, tasks, out, threads);
    timer.mark("splice");
}
}  // namespace

// ---- planar contact analysis (see bvh_build.hpp) ------------------------------------------------------------------------

void find_contacts(const cray_scene_desc& d, const RefBvh& ref, ContactInfo& out) {
    PhaseTimer timer;
    const size_t np = (size_t)d.n_primitives, nn = ref.nodes.size();
    out = ContactInfo{};
    out.node_flags.assign(nn, 0);
    out.prim_flag.assign(np, 0);
    if (nn == 0) return;
    // a hit location o + d * t carries a few ulps of the largest coordinate along the path
    double big = 1.0;
    for (int a = 0; a < 3; ++a) {
        big = std::max(big, std::max(std::fabs(ref.bounds.lo[a]), std::fabs(ref.bounds.hi[a])));
        big = std::max(big, std::fabs(d.camera.origin[a]));
    }
    const double noise = 16.0 * 2.220446049250313e-16 * big;
    out.noise = noise;
    const double grow = kContactTol + noise;
    std::unique_ptr<std::atomic<uint8_t>[]> marked(new std::atomic<uint8_t>[nn]);
    for (size_t i = 0; i < nn; ++i) marked[i].store(0, std::memory_order_relaxed);
    const unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<size_t> next{0};
    constexpr size_t kBatch = 4096;
    auto run = [&]() {
        std::vector<uint32_t> stack;
        for (;;) {
            const size_t b0 = next.fetch_add(kBatch);
            if (b0 >= np) break;
            for (size_t p = b0; p < std::min(np, b0 + kBatch); ++p) {
                const Box3 pb = primitive_bounds(d, p);
                bool planar[3], any = false;
                for (int a = 0; a < 3; ++a) { planar[a] = pb.hi[a] - pb.lo[a] <= kPlanarThickness; any |= planar[a]; }
                if (!any) continue;
                stack.clear();
                stack.push_back(0u);
                while (!stack.empty()) {
                    const uint32_t ni = stack.back();
                    stack.pop_back();
                    const BinNode& n = ref.nodes[ni];
                    bool overlap = true;
                    for (int a = 0; a < 3; ++a) overlap &= n.box.lo[a] - grow <= pb.hi[a] && n.box.hi[a] + grow >= pb.lo[a];
                    if (!overlap) continue;
                    for (int a = 0; a < 3; ++a) {
                        if (!planar[a] || !(n.box.hi[a] - n.box.lo[a] > kThinNode)) continue;
                        // some point of the primitive (its plane, give or take the rounding of a hit location) lies outside this
                        // face by no more than the reference's epsilon
                        const bool lo_face = n.box.lo[a] > pb.lo[a] - noise && n.box.lo[a] <= pb.hi[a] + grow;
                        const bool hi_face = n.box.hi[a] < pb.hi[a] + noise && n.box.hi[a] >= pb.lo[a] - grow;
                        if (lo_face || hi_face) {
                            marked[ni].store(1, std::memory_order_relaxed);
                            // (no path ray ever leaves an area light: its black matte ends the path and cancels the light sample, primitive.rs:43-46)
                            if (d.primitives[p].area_light < 0) out.prim_flag[p] = 1;
                        }
                    }
                    if (n.axis != 3) { stack.push_back(n.b); stack.push_back(n.a); }
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads && (size_t)t * kBatch < np; ++t) pool.emplace_back(run);
    run();
    for (auto& th : pool) th.join();
    // children follow their parents in the pre-order array: a reverse sweep sees them first
    for (size_t i = nn; i-- > 0;) {
        const BinNode& n = ref.nodes[i];
        uint8_t f = marked[i].load(std::memory_order_relaxed) ? (CONTACT_NODE | CONTACT_BELOW) : 0;
        if (n.axis != 3) f |= (out.node_flags[n.a] | out.node_flags[n.b]) & CONTACT_BELOW;
        out.node_flags[i] = f;
        out.n_nodes += f & CONTACT_NODE;
    }
    for (uint8_t f : out.prim_flag) out.n_prims += f;
    timer.mark("planar contact analysis");
}

// ---- 8-wide collapse ----------------------------------------------------------------------------------

namespace {

float float_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}

// ---- cost-optimal collapse (dynamic programme over the binary tree) ----------------------------------------------------
//
// Every primitive ends up alone in some leaf slot with its own box, so the SAH cost of the primitive tests is the same for
// every collapse; what varies is the sum over wide nodes of their surface area (the chance that a ray has to test the
// node's 8 child boxes).  With C(n, i) = the least such sum when the subtree of binary node n hangs off its parent as at
// most i children (Ylitie, Karras, Laine 2017, sec. 3.2):
//   D(n, j) = min over 0 < k < j of  C(left, k) + C(right, j - k)          j = 2..8
//   C(n, 1) = area(n) + D(n, 8)                                            n becomes a wide node
//   C(n, i) = min(D(n, i), C(n, i - 1))                                    i = 2..7
//   C(primitive, i) = 0
// Leaves of the reference tree with 2..4 primitives take part as small binary subtrees over their primitives.
constexpr uint32_t kPrimRef = 0x80000000u;

struct DpNode {
    uint32_t left, right;   // index into the DP array, or kPrimRef | primitive
    float area;
    float cost[7];          // C(n, 1..7)
    uint8_t split[7];       // for j = 2..8: k of D(n, j) in the low 3 bits; bit 7 (j <= 7): C(n, j) falls back to C(n, j - 1)
    uint8_t _pad;
};

struct Kid {
    uint32_t ref;    // kPrimRef | primitive, or DP node index
    Box3 box;
};

struct Collapser {
    const cray_scene_desc& desc;
    const RefBvh& ref;
    WideBvh& out;
    std::vector<DpNode> dp;          // [0, ref.nodes.size()): the reference nodes; then the virtual nodes inside multi-primitive leaves
    std::vector<Box3> virtual_box;   // boxes of the virtual nodes

    Box3 box_of(uint32_t r) const {
        if (r & kPrimRef) return primitive_bounds(desc, r & ~kPrimRef);
        return r < ref.nodes.size() ? ref.nodes[r].box : virtual_box[r - ref.nodes.size()];
    }
    float cost_of(uint32_t r, int i) const { return (r & kPrimRef) ? 0.0f : dp[r].cost[i - 1]; }

    void solve(DpNode& n) {
        float d[9];
        uint8_t dk[9];
        for (int j = 2; j <= 8; ++j) {
            float best = HUGE_VALF;
            int bk = 1;
            for (int k = 1; k < j; ++k) {
                if (k > 7 || j - k > 7) continue;
                const float c = cost_of(n.left, k) + cost_of(n.right, j - k);
                if (c < best) { best = c; bk = k; }
            }
            d[j] = best;
            dk[j] = (uint8_t)bk;
        }
        n.cost[0] = n.area + d[8];
        n.split[6] = dk[8];
        for (int i = 2; i <= 7; ++i) {
            if (d[i] <= n.cost[i - 2]) { n.cost[i - 1] = d[i]; n.split[i - 2] = dk[i]; }
            else { n.cost[i - 1] = n.cost[i - 2]; n.split[i - 2] = (uint8_t)(dk[i] | 0x80u); }
        }
    }

    uint32_t ref_of_child(uint32_t bin) const {
        const BinNode& c = ref.nodes[bin];
        return (c.axis == 3 && c.b == 1) ? (kPrimRef | ref.prim_order[c.a]) : bin;
    }
    uint32_t add_virtual(uint32_t l, uint32_t r) {
        DpNode v{};
        v.left = l; v.right = r;
        const Box3 b = box_union(box_of(l), box_of(r));
        v.area = (float)box_surface_area(b);
        virtual_box.push_back(b);
        solve(v);
        dp.push_back(v);
        return (uint32_t)dp.size() - 1;
    }

    void run_dp() {
        const size_t nr = ref.nodes.size();
        dp.assign(nr, DpNode{});
        // children follow their parents in the pre-order array: a reverse sweep sees them first
        for (size_t i = nr; i-- > 0;) {
            const BinNode& b = ref.nodes[i];
            DpNode n{};
            n.area = (float)box_surface_area(b.box);
            if (b.axis == 3) {
                if (b.b == 1) continue;  // referenced as a primitive, never as a node
                uint32_t p[4];
                for (uint32_t k = 0; k < b.b; ++k) p[k] = kPrimRef | ref.prim_order[b.a + k];
                if (b.b == 2) { n.left = p[0]; n.right = p[1]; }
                else if (b.b == 3) { n.left = add_virtual(p[0], p[1]); n.right = p[2]; }
                else { n.left = add_virtual(p[0], p[1]); n.right = add_virtual(p[2], p[3]); }
            } else {
                n.left = ref_of_child(b.a);
                n.right = ref_of_child(b.b);
            }
            solve(n);
            dp[i] = n;
        }
    }

    // The children subtree r contributes when it may use at most i slots of its parent.
    void collect(uint32_t r, int i, std::vector<Kid>& kids) const {
        if (r & kPrimRef) { kids.push_back({r, box_of(r)}); return; }
        const DpNode& n = dp[r];
        while (i >= 2 && (n.split[i - 2] & 0x80u)) i -= 1;
        if (i == 1) { kids.push_back({r, box_of(r)}); return; }
        const int k = n.split[i - 2] & 7;
        collect(n.left, k, kids);
        collect(n.right, i - k, kids);
    }

    // Wide nodes (not counting r's own) and primitives the subtree of wide node r occupies: integers from the DP decisions only.
    struct Extent { uint64_t nodes, prims; };
    void count_slots(uint32_t r, int i, Extent& e) const {
        if (r & kPrimRef) { e.prims += 1; return; }
        const DpNode& n = dp[r];
        while (i >= 2 && (n.split[i - 2] & 0x80u)) i -= 1;
        if (i == 1) { e.nodes += 1; count_node(r, e); return; }
        const int k = n.split[i - 2] & 7;
        count_slots(n.left, k, e);
        count_slots(n.right, i - k, e);
    }
    void count_node(uint32_t r, Extent& e) const {
        if (r & kPrimRef) { e.prims += 1; return; }
        const DpNode& n = dp[r];
        const int k = n.split[6] & 7;
        count_slots(n.left, k, e);
        count_slots(n.right, 8 - k, e);
    }

    // Layout = depth-first: a node's primitives and its block of interior children are placed when the node is visited, then
    // each child's subtree follows in order.  A subtree therefore owns one contiguous range of nodes and of primitives, and
    // subtrees below kTaskDepth are written by a thread pool into ranges reserved from their extents -- the result is the
    // array a single thread would have produced.
    struct Cursor { uint32_t node, prim; };
    struct EmitTask { uint32_t wide_idx, ref, depth; Cursor cursor; };
    static constexpr uint32_t kTaskDepth = 3;
    std::vector<EmitTask>* emit_tasks = nullptr;  // non-null while the top of the tree is laid out

    void expand(uint32_t wide_idx, uint32_t r, uint32_t depth, Cursor& cur, uint32_t& deepest) {
        if (emit_tasks && depth > kTaskDepth && !(r & kPrimRef)) {
            Extent e{0, 0};
            count_node(r, e);
            emit_tasks->push_back({wide_idx, r, depth, cur});
            cur.node += (uint32_t)e.nodes;
            cur.prim += (uint32_t)e.prims;
            return;
        }
        std::vector<Kid> kids;
        if (r & kPrimRef) kids.push_back({r, box_of(r)});  // a scene of one primitive
        else {
            const DpNode& n = dp[r];
            const int k = n.split[6] & 7;
            collect(n.left, k, kids);
            collect(n.right, 8 - k, kids);
        }
        emit(wide_idx, kids, box_of(r), depth, cur, deepest);
    }

    void emit(uint32_t wide_idx, const std::vector<Kid>& kids, const Box3& node_box, uint32_t depth, Cursor& cur, uint32_t& deepest) {
        deepest = std::max(deepest, depth);
        const int k = (int)kids.size();
        // slot assignment: slot bits (x=4, y=2, z=1) set = child sits on the + side of the node centre on that axis
        const V3 nc = box_centroid(node_box);
        double cost[8][8];
        for (int c = 0; c < k; ++c) {
            const V3 dc = box_centroid(kids[c].box) - nc;
            for (int s = 0; s < 8; ++s) cost[c][s] = ((s & 4) ? dc.x : -dc.x) + ((s & 2) ? dc.y : -dc.y) + ((s & 1) ? dc.z : -dc.z);
        }
        int slot_of[8], child_in[8];
        for (int i = 0; i < 8; ++i) { slot_of[i] = -1; child_in[i] = -1; }
        for (int round = 0; round < k; ++round) {
            int bc = -1, bs = -1;
            double best = -HUGE_VAL;
            for (int c = 0; c < k; ++c) {
                if (slot_of[c] >= 0) continue;
                for (int s = 0; s < 8; ++s)
                    if (child_in[s] < 0 && cost[c][s] > best) { best = cost[c][s]; bc = c; bs = s; }
            }
            slot_of[bc] = bs;
            child_in[bs] = bc;
        }
        // quantisation frame
        WideNode w{};
        const float p[3] = {float_down(node_box.lo.x), float_down(node_box.lo.y), float_down(node_box.lo.z)};
        w.px = p[0]; w.py = p[1]; w.pz = p[2];
        int e[3];
        double scale[3];
        for (int ax = 0; ax < 3; ++ax) {
            const double ext = node_box.hi[ax] - (double)p[ax];
            int ex = -40;
            if (ext > 0.0) {
                int fe;
                std::frexp(ext / 255.0, &fe);  // ext/255 = m * 2^fe, m in [0.5,1)  =>  2^fe >= ext/255
                ex = fe;
                while (std::ceil(ext / std::ldexp(1.0, ex)) > 255.0) ex += 1;
            }
            ex = std::max(-100, std::min(100, ex));
            e[ax] = ex;
            scale[ax] = std::ldexp(1.0, ex);
        }
        w.ex = (uint8_t)(e[0] + 127); w.ey = (uint8_t)(e[1] + 127); w.ez = (uint8_t)(e[2] + 127);
        w.prim_base = cur.prim;
        std::vector<uint32_t> interior_kids;
        for (int s = 0; s < 8; ++s) {
            const int c = child_in[s];
            if (c < 0) continue;
            const Kid& kid = kids[c];
            for (int ax = 0; ax < 3; ++ax) {
                double ql = std::floor((kid.box.lo[ax] - (double)p[ax]) / scale[ax]);
                double qh = std::ceil((kid.box.hi[ax] - (double)p[ax]) / scale[ax]);
                while (ql > 0.0 && (double)p[ax] + ql * scale[ax] > kid.box.lo[ax]) ql -= 1.0;
                while (qh < 255.0 && (double)p[ax] + qh * scale[ax] < kid.box.hi[ax]) qh += 1.0;
                ql = std::max(0.0, std::min(255.0, ql));
                qh = std::max(0.0, std::min(255.0, qh));
                w.qlo[ax][s] = (uint8_t)ql;
                w.qhi[ax][s] = (uint8_t)qh;
            }
            if (kid.ref & kPrimRef) {
                w.leafmask |= (uint8_t)(1u << s);
                out.prim_order[cur.prim++] = kid.ref & ~kPrimRef;
            } else {
                w.imask |= (uint8_t)(1u << s);
                interior_kids.push_back(kid.ref);
            }
        }
        w.child_base = cur.node;
        out.nodes[wide_idx] = w;
        const uint32_t base = cur.node;
        cur.node += (uint32_t)interior_kids.size();
        for (size_t i = 0; i < interior_kids.size(); ++i) expand(base + (uint32_t)i, interior_kids[i], depth + 1, cur, deepest);
    }
};

}  // namespace

void collapse_to_wide(const cray_scene_desc& d, const RefBvh& ref, WideBvh& out) {
    out = WideBvh{};
    PhaseTimer timer;
    Collapser c{d, ref, out};
    c.run_dp();
    timer.mark("collapse: dynamic programme");
    const uint32_t root = c.ref_of_child(0);
    Collapser::Extent total{1, 0};
    c.count_node(root, total);
    out.nodes.resize(total.nodes);
    out.prim_order.resize(total.prims);
    // the top of the tree on this thread, handing out the ranges of the subtrees below it ...
    std::vector<Collapser::EmitTask> tasks;
    c.emit_tasks = &tasks;
    Collapser::Cursor cur{1, 0};
    uint32_t deepest = 0;
    c.expand(0, root, 1, cur, deepest);
    c.emit_tasks = nullptr;
    // ... and the subtrees on all threads
    const unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<uint32_t> task_depth(tasks.size(), 0);
    std::atomic<size_t> next{0};
    auto run = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= tasks.size()) break;
            Collapser::Cursor local = tasks[i].cursor;
            c.expand(tasks[i].wide_idx, tasks[i].ref, tasks[i].depth, local, task_depth[i]);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads && t < tasks.size(); ++t) pool.emplace_back(run);
    run();
    for (auto& th : pool) th.join();
    for (uint32_t dd : task_depth) deepest = std::max(deepest, dd);
    out.depth = deepest;
    timer.mark("collapse: emit");
}

}  // namespace cray
