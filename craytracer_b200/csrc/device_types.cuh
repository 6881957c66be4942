// Device-resident scene layout (HBM) shared by all kernels.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include "../../include/cray_b200.h"
#include "bvh_build.hpp"
#include "cray_math.cuh"

namespace cray {

enum : uint32_t { PRIM_SPHERE = CRAY_SHAPE_SPHERE, PRIM_TRIANGLE = CRAY_SHAPE_TRIANGLE, PRIM_DISK = CRAY_SHAPE_DISK };

// 80-byte intersection record, stored in leaf order so that the <= 4 primitives of a leaf are contiguous.
//   triangle: d = v0, e1, e2                       (Shape::Triangle, shape.rs:30-33)
//   sphere  : d[0..2] = origin, d[3] = radius      (object_to_world = translate(origin))
//   disk    : aux = index into SceneView::disks
struct alignas(16) LeafPrim {
    double d[9];
    uint32_t prim;   // index in reference primitive order
    uint32_t kind;   // PRIM_* | shade class << 8 (k_shade groups paths by it) | disk: index << 16, triangle: kKindFlatTriangle | kKindContact
};
static_assert(sizeof(LeafPrim) == 80, "LeafPrim layout");

// 48-byte f32 intersection record of the opt-in F32 mode, parallel to SceneView::wide_prims (same leaf slot):
// the three vertices rounded to f32 -- triangles sharing a vertex share its rounded value, so the mesh stays closed.
// Spheres and disks keep `kind` only; their test reads the f64 record of the same slot.
struct alignas(16) Tri32 {
    float v0[3], v1[3], v2[3];
    uint32_t prim, kind;   // as LeafPrim
    uint32_t _pad;
};
static_assert(sizeof(Tri32) == 48, "Tri32 layout");

struct DiskXf {      // Shape::Disk, shape.rs:41-46
    Affine o2w;      // object_to_world.matrix
    Affine w2o;      // object_to_world.inverse == world_to_object.matrix
    double radius, inner_radius;
};

// Shading attributes of a triangle primitive, indexed by PRIMITIVE index (slots of non-triangle primitives are unused).
// The first 32-byte sector is all a flat-shaded, untextured triangle needs: its normal and what Primitive binds to it.
struct alignas(32) TriShade {
    double n0[3];
    int32_t material;      // into SceneView::materials (the black matte for area lights, primitive.rs:43-46)
    int32_t area_light;    // into SceneView::lights, or -1
    double n01[3], n02[3];
    double uv0[2], uv01[2], uv02[2];
};
static_assert(sizeof(TriShade) == 128, "TriShade layout");
constexpr uint32_t kKindFlatTriangle = 1u << 16;  // LeafPrim::kind flag: n01 and n02 are all +0.0 (no vertex normals)
// LeafPrim::kind flag: the primitive is planar and coincides with a face of some BVH node box, so a ray leaving it may start in the
// outer shell of that box, where the reference's box test culls it (bvh_build.hpp "planar contact")
constexpr uint32_t kKindContact = 1u << 31;
CRAY_HD uint32_t disk_index(uint32_t kind) { return (kind >> 16) & 0x7FFFu; }

enum : uint32_t { LOBE_LAMBERTIAN = 0, LOBE_OREN_NAYAR = 1, LOBE_CONDUCTOR = 2, LOBE_SPECULAR_BRDF = 3, LOBE_SPECULAR_BTDF = 4, LOBE_FRESNEL_SPECULAR = 5 };

struct DevLobe {     // BxDF, bxdf.rs:23-50
    uint32_t kind, _pad;
    cray_texture_desc t0, t1, sigma;
    double eta_i, eta_t;
};
struct DevMaterial { // Material, material.rs:13-17
    uint32_t is_bsdf, n_lobes;
    uint32_t needs_uv;         // some texture of some lobe is not a constant: the hit's texture coordinates are read
    uint32_t all_delta;        // every lobe is a perfect-specular one: BxDF::f is black for every direction (bxdf.rs:259-264)
    DevLobe lobes[2];
};
struct DevImage {
    uint32_t width, height;
    uint64_t offset;  // into SceneView::texels (RGB8)
};
struct DevLight {    // Light, light.rs:25-43
    uint32_t kind;
    int32_t prim;
    double v[3];
    double color[3];
    LeafPrim shape;  // AREA: the emitting shape (for Shape::sample / pdf_from)
    double area;     // Shape::area()
};
struct DevCamera {   // Camera, camera.rs:15-23
    double camera_from_raster[4][4];
    Affine world_from_camera;
    double lens_radius, focal_distance;
    uint32_t perspective, width, height, _pad;
};

struct SceneView {
    // exact traversal (reference binary BVH)
    const BinNode* bin_nodes;
    const LeafPrim* bin_prims;       // reference leaf order
    // fast traversal (8-wide BVH)
    const WideNode* wide_nodes;
    const LeafPrim* wide_prims;      // wide leaf order
    const uint32_t* rank_of_prim;    // primitive -> rank in the reference leaf order (exact-t tie breaking)
    const uint32_t* wide_slot_of_prim;  // primitive -> wide leaf slot (hits of the reference-order traversal inside the fast mode)
    const uint8_t* bin_contact;      // per binary node: CONTACT_NODE | CONTACT_BELOW (bvh_build.hpp); null when no node is marked
    const Tri32* wide_tris32;        // F32 mode: f32 triangles in wide leaf order (null unless CRAY_BUILD_F32)
    const DiskXf* disks;
    // shading
    const cray_primitive_desc* prims;
    const TriShade* tri_shade;
    const cray_sphere_desc* spheres;
    const DevMaterial* materials;
    const DevImage* images;
    const uint8_t* texels;
    const double* gamma_lut;         // (c/255)^2.2, c = 0..255 (Color::from_rgb color.rs:39-46)
    const DevLight* lights;
    const double* light_cdf;
    uint32_t n_lights, max_depth;
    DevCamera camera;
    Box3 bounds;
};

}  // namespace cray
