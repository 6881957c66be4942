// Host BVH construction.
//  * build_reference_bvh: the reference's binned-SAH binary tree (src/bvh.rs:234-336 with src/util.rs
//    partition_by), reproduced decision-for-decision so that leaf contents, leaf order and split axes
//    are identical to the reference's -- they define the visit order and therefore which primitive wins
//    an exact-t tie.  Sub-trees are built by a thread pool (the tree is deterministic, the order of
//    construction is not observable).
//  * collapse_to_wide: greedy surface-area collapse of that tree into the GPU's 8-wide node format with
//    8-bit quantised child boxes (conservatively rounded outward) and octant-ordered child slots.  Leaves of
//    the binary tree are opened too: a leaf slot of a wide node holds exactly ONE primitive with its own
//    quantised box, so the f64 primitive test only runs on primitives whose own box the ray enters.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/cray_b200.h"
#include "cray_math.cuh"

namespace cray {

struct PhaseTimer {  // CRAY_BUILD_TIMING=1 prints where scene construction spends its time
    bool on = std::getenv("CRAY_BUILD_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cray build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct BinNode {       // 64 B, uploaded as-is for the exact traversal mode
    Box3 box;          // f64 bounds, bit-identical to the reference's node bounds
    uint32_t a, b;     // interior: left, right node index (pre-order); leaf: first, count in leaf order
    uint32_t axis;     // 0..2 split axis, 3 = leaf
    uint32_t right_first;  // interior: leaf-order rank of the first primitive of the right subtree
};
static_assert(sizeof(BinNode) == 64, "BinNode layout");

struct RefBvh {
    std::vector<BinNode> nodes;        // pre-order, root = 0
    std::vector<uint32_t> prim_order;  // leaf order -> primitive index
    Box3 bounds;
    std::string error;                 // non-empty where the reference would panic (bvh.rs:245, :327-328)
};

// Bounds of primitive i exactly as Shape::bounds computes them (src/shape.rs:402-438).
Box3 primitive_bounds(const cray_scene_desc& d, uint64_t prim);
// `gpu_device` >= 0: scenes of more than 2^16 primitives build the tree on that device (bvh_build_gpu.cu: the same tree, node for
// node); the host builder takes over if there is no usable device or CRAY_GPU_BUILD=0.
void build_reference_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads = 0, int gpu_device = -1);

struct BuildItem {  // PrimitiveInfo (bvh.rs:149-154); the centroid is recomputed from the box (same expression, same bits)
    double lo[3], hi[3];
    uint32_t prim;
    uint32_t bucket;
};
static_assert(sizeof(BuildItem) == 56, "BuildItem layout");
// The tree over `items` (their boxes, in primitive order) built on `device`; false (with the reason) if that did not happen.
bool build_reference_bvh_gpu(const BuildItem* items, size_t n, int device, RefBvh& out, std::string& why);
void finish_device_build_release();           // waits for the background release of the last device build's buffers
bool last_reference_build_was_on_device();   // of this thread's last build_reference_bvh call (tests)
// The same builder choosing, at every node, the cheapest of the three axes' best binned splits instead of the reference's
// longest-centroid-axis rule: the tree the 8-wide BVH can be collapsed from (hits do not depend on it: the f64 leaf tests decide,
// and exact-t ties are resolved on the reference tree).
void build_quality_bvh(const cray_scene_desc& d, RefBvh& out, unsigned threads = 0);

#ifndef CRAY_NODE96
#define CRAY_NODE96 0   // 1 (tuning): nodes padded to 96 bytes on 32-byte boundaries, fetched with three 256-bit loads
#endif
struct alignas(CRAY_NODE96 ? 32 : 16) WideNode {  // 80 B: five 16-byte loads
    float px, py, pz;          // quantisation frame origin
    uint8_t ex, ey, ez;        // per-axis scale = 2^(e - 127) (raw f32 exponent field)
    uint8_t imask;             // bit s: slot s holds an interior child
    uint32_t child_base;       // first interior child (children contiguous, ascending slot)
    uint32_t prim_base;        // first primitive of this node in wide leaf order (primitives contiguous, ascending slot)
    uint8_t leafmask;          // bit s: slot s holds exactly one primitive
    uint8_t _pad[7];
    uint8_t qlo[3][8];         // quantised child box minima  [axis][slot]
    uint8_t qhi[3][8];         // quantised child box maxima
};
static_assert(sizeof(WideNode) == (CRAY_NODE96 ? 96 : 80), "WideNode layout");

constexpr int kWideStackLimit = 24;  // traversal stack entries per ray (traverse.cuh): one per level

struct WideBvh {
    std::vector<WideNode> nodes;
    std::vector<uint32_t> prim_order;  // wide leaf order -> primitive index
    uint32_t depth = 0;
};
void collapse_to_wide(const cray_scene_desc& d, const RefBvh& ref, WideBvh& out);

// ---- planar contact (the reference's false box misses, SURVEY A-4b) -------------------------------------------------------
//
// Bvh::intersect / intersects visit a node iff `bounds.intersects(ray) || bounds.contains(origin)` (bvh.rs:70,:117), and
// Bounds::intersects (bounds.rs:62-88) accepts only if the slab entry OR exit distance lies in (1e-9, ray.max_distance).  A ray
// whose origin lies OUTSIDE a box by no more than 1e-9 (entry <= 1e-9, not "contained") and that ends inside the box -- or has
// already found a hit nearer than the box's exit -- culls the whole subtree: the "box shaped hole" of scenes/rounding-error.cry.
// Only origins in that outer shell of some node box can make the reference's result differ from a conservative traversal's, and
// origins land there systematically only where a planar primitive coincides with a face of a node box (a floor at the root
// box's minimum, a ground plane under a resting sphere).  find_contacts marks those boxes and primitives at build time; at
// run time a ray that leaves a marked primitive is tested against the marked boxes (traverse.cuh: origin_in_contact_shell)
// and, if it starts in such a shell, is traced by the reference-order traversal instead of the wide one.
constexpr double kContactTol = 1.001e-9;       // EPSILON (constants.rs:1) for a unit direction, with margin
constexpr double kPlanarThickness = 1e-5;      // a primitive whose box is thinner than this on an axis is planar on it
constexpr double kThinNode = 2.5e-9;           // node boxes thinner than this on the axis hold only geometry in the plane itself
enum : uint8_t { CONTACT_NODE = 1, CONTACT_BELOW = 2 };

struct ContactInfo {
    std::vector<uint8_t> node_flags;   // per binary node: CONTACT_NODE | CONTACT_BELOW (the node or a descendant is marked)
    std::vector<uint8_t> prim_flag;    // per primitive: rays leaving it can start in the outer shell of a marked box
    uint64_t n_nodes = 0, n_prims = 0;
    double noise = 0.0;                // absolute rounding allowance for computed hit locations
};
void find_contacts(const cray_scene_desc& d, const RefBvh& ref, ContactInfo& out);

}  // namespace cray
