/*
 * cray_b200.h — C ABI of the B200-native path-tracing core for craytracer scenes.
 *
 * The reference (banga/craytracer, pure Rust) has no FFI/plugin layer; its seams
 * for the hot path are plain Rust calls.  Every entry point below replaces one of
 * those seams (cited as file:line relative to the reference tree) with plain
 * pointers and sizes, so that a Rust host can bind it with `extern "C"` (see
 * INTEGRATION.md for the binding a maintainer would add).
 *
 *   S0  Scene::new / parse_scene        src/scene.rs:25, src/scene_parser.rs:1078
 *   S1  render                          src/bin/craytracer.rs:224
 *   S2  path_integrator::estimate_Li    src/path_integrator.rs:41
 *   S3  Scene::intersect / intersects   src/scene.rs:55, :59
 *
 * Conventions: every function returns CRAY_OK (0) or a negative CRAY_E_* code and
 * never throws or aborts across the ABI (reference panics become codes; the text
 * is available from cray_last_error()).  The caller owns every buffer it passes.
 * There is no CPU fallback: if no CUDA device is usable the calls fail.
 */
#ifndef CRAY_B200_H
#define CRAY_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRAY_OK 0
#define CRAY_E_INVALID (-1)   /* bad argument / malformed description            */
#define CRAY_E_CUDA (-2)      /* CUDA runtime error (no device, OOM, launch...)  */
#define CRAY_E_PARSE (-3)     /* scene file error (ParserError, scene_parser.rs:72) */
#define CRAY_E_IO (-4)        /* file not found / unreadable                     */
#define CRAY_E_BVH (-5)       /* the reference's SAH builder would panic (bvh.rs:245,:327) */
#define CRAY_E_UNSUPPORTED (-6)

#define CRAY_NO_HIT 0xFFFFFFFFu

/* ---- flat scene description (what Scene::new is given, src/scene.rs:25) ------- */

/* Shape constructor arguments, reference primitive order is preserved by the
 * cray_primitive_desc array that indexes into these. */
typedef struct cray_sphere_desc {   /* Shape::new_sphere  src/shape.rs:56 */
    double origin[3];
    double radius;
} cray_sphere_desc;

typedef struct cray_triangle_desc { /* Shape::Triangle fields  src/shape.rs:30-40 */
    double v0[3], e1[3], e2[3];
    double n0[3], n01[3], n02[3];
    double uv0[2], uv01[2], uv02[2];
} cray_triangle_desc;

typedef struct cray_disk_desc {     /* Shape::new_disk  src/shape.rs:133 */
    double origin[3];
    double rotate_x, rotate_y;      /* degrees */
    double radius, inner_radius;
} cray_disk_desc;

enum { CRAY_SHAPE_SPHERE = 0, CRAY_SHAPE_TRIANGLE = 1, CRAY_SHAPE_DISK = 2 };

typedef struct cray_primitive_desc { /* Primitive  src/primitive.rs:15-25 */
    uint32_t shape_kind;
    uint32_t shape_index;            /* into spheres / triangles / disks        */
    int32_t material;                /* into materials; ignored for area lights */
    int32_t area_light;              /* into lights, or -1                      */
} cray_primitive_desc;

enum { CRAY_TEX_CONSTANT = 0, CRAY_TEX_CHECKERBOARD = 1, CRAY_TEX_IMAGE = 2 };

typedef struct cray_texture_desc {   /* Texture<T>  src/texture.rs:8-12 */
    uint32_t kind;
    int32_t image;                   /* into images (kind == IMAGE)             */
    double a[3];                     /* constant value / checker "a"; Texture<f64> uses a[0] */
    double b[3];                     /* checker "b"                             */
    double scale;                    /* checker scale                           */
} cray_texture_desc;

typedef struct cray_image_desc {     /* image::RgbImage, row-major RGB8         */
    uint32_t width, height;
    const uint8_t* rgb;
} cray_image_desc;

enum { CRAY_MAT_MATTE = 0, CRAY_MAT_GLASS = 1, CRAY_MAT_PLASTIC = 2, CRAY_MAT_METAL = 3 };

typedef struct cray_material_desc {  /* arguments of Material::new_*  src/material.rs:20-70 */
    uint32_t kind;
    uint32_t _pad;
    cray_texture_desc t0;            /* matte reflectance | glass reflectance   | plastic diffuse   | metal eta */
    cray_texture_desc t1;            /* -                 | glass transmittance | plastic specular  | metal k   */
    cray_texture_desc t2;            /* matte sigma (f64) | -                   | plastic roughness (f64) | -   */
    double eta;                      /* glass eta_t                              */
} cray_material_desc;

enum { CRAY_LIGHT_POINT = 0, CRAY_LIGHT_DISTANT = 1, CRAY_LIGHT_INFINITE = 2, CRAY_LIGHT_AREA = 3 };

typedef struct cray_light_desc {     /* Light  src/light.rs:25-43 */
    uint32_t kind;
    int32_t primitive;               /* AREA: the primitive whose shape emits   */
    double v[3];                     /* POINT origin | DISTANT direction (already normalised, scene_parser.rs:888) */
    double color[3];                 /* intensity | emittance                   */
} cray_light_desc;

enum { CRAY_CAMERA_PERSPECTIVE = 0, CRAY_CAMERA_ORTHOGRAPHIC = 1 };

typedef struct cray_camera_desc {    /* Camera::perspective / orthographic args  src/camera.rs:76,:105 */
    uint32_t kind;
    uint32_t width, height;          /* Film  src/film.rs:2 */
    uint32_t _pad;
    double origin[3], target[3], up[3];
    double fov;                      /* degrees; perspective only */
    double lens_radius, focal_distance;
} cray_camera_desc;

typedef struct cray_scene_desc {
    uint32_t max_depth;              /* scene_parser.rs:796 default 8 */
    uint32_t num_samples;            /* scene_parser.rs:797 default 4 */
    cray_camera_desc camera;
    uint64_t n_spheres, n_triangles, n_disks, n_primitives, n_materials, n_lights, n_images;
    const cray_sphere_desc* spheres;
    const cray_triangle_desc* triangles;
    const cray_disk_desc* disks;
    const cray_primitive_desc* primitives;   /* reference primitive order */
    const cray_material_desc* materials;
    const cray_light_desc* lights;            /* explicit lights, then area lights in primitive order (scene_parser.rs:1088-1101) */
    const cray_image_desc* images;
} cray_scene_desc;

/* ---- host-side scene ingest (S0; src/scene_parser.rs:1078, src/obj.rs:26) ----- */

typedef struct cray_host_scene cray_host_scene;   /* owns the arrays a cray_scene_desc points into */

/* parse_scene on the text of a .cry file.  Mesh paths are resolved against
 * `base_dir` (the reference resolves them against the process CWD). */
int cray_host_scene_parse(const char* cry_text, const char* base_dir, cray_host_scene** out);
int cray_host_scene_load(const char* cry_path, const char* base_dir, cray_host_scene** out);
const cray_scene_desc* cray_host_scene_desc(const cray_host_scene*);
void cray_host_scene_destroy(cray_host_scene*);
/* line/column of the last CRAY_E_PARSE on this thread (Location, scene_parser.rs:1-11); 0:0 if none */
void cray_last_error_location(uint32_t* line, uint32_t* column);

/* Warnings gathered while reading a scene (unused keys, MTL problems, stand-in meshes). */
uint64_t cray_host_scene_num_warnings(const cray_host_scene*);
const char* cray_host_scene_warning(const cray_host_scene*, uint64_t i);
/* The reference's repository does not ship two of its meshes (objs/xyzrgb_dragon.obj, objs/staircase/staircase.obj,
 * .MISSING_LARGE_BLOBS).  A scene that names `file_name` and does not find it on disk gets a seeded procedural mesh instead:
 * kind 0 = closed displaced tube of `triangles` triangles standing on y = -40 (the dragon's extent), kind 1 = interior with
 * stairs using the materials of the staircase MTL.  Process-wide; without a registration a missing mesh is CRAY_E_IO. */
void cray_register_standin_mesh(const char* file_name, int kind, uint64_t triangles, uint64_t seed);

/* ---- device scene (S0) --------------------------------------------------------- */

typedef struct cray_scene cray_scene;   /* opaque; owns all device memory on its GPU */

/* Traversal structures built at create time. */
#define CRAY_BUILD_EXACT 1u   /* reference binary SAH BVH, f64 boxes, reference visit order (bvh.rs:58-147) */
#define CRAY_BUILD_FAST 2u    /* 8-wide quantised BVH collapsed from the same tree */
#define CRAY_BUILD_F32 4u     /* + 48-byte f32 triangle records beside the wide BVH (implies CRAY_BUILD_FAST) */

int cray_scene_create(const cray_scene_desc* desc, int device, uint32_t build_flags, cray_scene** out);
/* The same scene on n devices of this process (for cray_render_multi): the BVH build and the record arrays are made once
 * on the host and uploaded to every device by its own thread.  out[0..n) are all set, or all null on failure. */
int cray_scene_create_multi(const cray_scene_desc* desc, const int* devices, int n, uint32_t build_flags, cray_scene** out);
void cray_scene_destroy(cray_scene*);

/* ---- S3: fixed ray batches (Scene::intersect / intersects, src/scene.rs:55,:59) */

typedef struct cray_ray {          /* Ray  src/ray.rs:7-11 */
    double origin[3];
    double direction[3];
    double max_distance;           /* +inf for Ray::new */
} cray_ray;

typedef struct cray_hit {
    uint32_t prim;                 /* index into desc.primitives, CRAY_NO_HIT on miss */
    uint32_t _pad;
    double t;                      /* PrimitiveIntersection::distance */
    double u, v;                   /* triangle: Moeller-Trumbore barycentrics (shape.rs:236-243); sphere/disk: surface uv */
} cray_hit;

typedef struct cray_surface {      /* the rest of PrimitiveIntersection  src/intersection.rs:17-24 */
    double location[3];
    double normal[3];
    double uv[2];
} cray_surface;

/* EXACT and FAST return the primitive and distance of the reference's algorithm (as restated by the CPU oracle) bit for bit.
 * EXACT walks the reference's binary tree in the reference's order for every ray.  FAST walks the 8-wide tree, except for rays
 * whose origin lies in the 1e-9 outer shell of a node box that a planar primitive touches ("planar contact": only there can the
 * reference's box test, bounds.rs:62-88, cull a subtree a conservative traversal would enter -- the documented bug of
 * scenes/rounding-error.cry); those rays are traced in reference order too, so the false misses are reproduced.
 * F32 is the opt-in fast mode of SURVEY 8f n4: the same wide BVH, triangles tested in
 * f32 with the watertight test of Woop, Benthin and Wald (2013) on ray-relative vertices, a ray never re-hits the triangle it
 * leaves, and the hit that was found is re-evaluated in f64 -- so t, u, v equal the parity modes' whenever the same primitive
 * is found, which f32 cannot guarantee for rays grazing an edge.  Spheres and disks stay in f64. */
enum { CRAY_TRAVERSE_EXACT = 0, CRAY_TRAVERSE_FAST = 1, CRAY_TRAVERSE_F32 = 2 };

/* Host buffers in, host buffers out (copies are part of the call). `surf` may be NULL. */
int cray_trace_closest(cray_scene*, int mode, const cray_ray* rays, uint64_t n, cray_hit* hits, cray_surface* surf);
int cray_trace_any(cray_scene*, int mode, const cray_ray* rays, uint64_t n, uint8_t* occluded);
/* Same, on buffers already resident on the scene's device; runs on `stream` (a cudaStream_t) and does not synchronise. */
int cray_trace_closest_device(cray_scene*, int mode, const cray_ray* d_rays, uint64_t n, cray_hit* d_hits, cray_surface* d_surf, void* stream);
int cray_trace_any_device(cray_scene*, int mode, const cray_ray* d_rays, uint64_t n, uint8_t* d_occluded, void* stream);

/* ---- S2: radiance samples (render_pixel + estimate_Li, craytracer.rs:148, path_integrator.rs:41) */

/* One radiance sample per (x[i], y[i], sample_index[i]) with SobolSampler::new(seed, ..). rgb: n*3 f64. */
int cray_estimate_li(cray_scene*, int mode, uint64_t seed, const uint32_t* x, const uint32_t* y,
                     const uint32_t* sample_index, uint64_t n, double* rgb);

/* ---- S1: whole frame (render, src/bin/craytracer.rs:224) ----------------------- */

typedef struct cray_render_stats {
    uint64_t samples;          /* W*H*(sample_end-sample_begin) */
    uint64_t closest_rays;     /* Scene::intersect calls  */
    uint64_t shadow_rays;      /* Scene::intersects calls the reference makes: one per path vertex that hits (path_integrator.rs:141) */
    uint64_t nan_samples;      /* samples dropped where the reference would assert (path_integrator.rs:208) */
    uint64_t iterations;       /* wavefront iterations */
    uint64_t kernel_launches;  /* kernels launched by this call */
    double render_ms;          /* device time of the render region (CUDA events) */
    double trace_ms;           /* device time inside the closest-hit (extend) traversal launches */
    double shadow_ms;          /* ... inside the any-hit (shadow) traversal launches */
    double shade_ms;           /* ... inside the shading launches */
    double generate_ms;        /* ... inside the flush / camera-ray launches */
    uint64_t shadow_rays_traced; /* shadow rays this implementation traced: a vertex whose light sample cannot contribute (black
                                    contribution, all-specular material) makes the reference's call above but needs no ray here */
    uint64_t contact_rays;     /* closest + shadow rays of the fast mode that started in a contact shell and took the reference-order traversal */
} cray_render_stats;

/* Samples [sample_begin, sample_end) of every pixel; film is the SUM over those samples
 * (W*H*3 f32, row-major RGB, offset = x + y*W as craytracer.rs:183).  Host film. */
int cray_render(cray_scene*, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                float* rgb_sum, cray_render_stats* stats);
/* Device film (f32, W*H*3, overwritten); runs on `stream`, synchronises it before returning stats. */
int cray_render_device(cray_scene*, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                       float* d_rgb_sum, void* stream, cray_render_stats* stats);

/* One frame on n GPUs of this process.  scenes[k] was created on its own device from the same description; GPU k
 * renders the k-th contiguous slice of [sample_begin, sample_end) for every pixel, the f32 sum films are combined by ONE
 * ncclReduce(sum) onto scenes[0]'s device and copied to `rgb_sum` (host, W*H*3).  NCCL is loaded at run time (libnccl.so.2);
 * CRAY_E_UNSUPPORTED if it is not installed.  Counters in `stats` are sums over GPUs, times are the slowest GPU's. */
int cray_render_multi(cray_scene* const* scenes, int n, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                      float* rgb_sum, cray_render_stats* stats);

/* ---- output (the EXR save of src/bin/craytracer.rs:367-369) ----------------------- */

/* Writes `rgb` (W*H*3 f32, row-major, linear RGB -- what render() hands to on_render_finish) as an OpenEXR file with three
 * 32-bit float channels (the reference's `write_rgb_file`, exr crate): single-part scan-line image, no compression. */
int cray_write_exr(const char* path, uint32_t width, uint32_t height, const float* rgb);

/* ---- introspection ------------------------------------------------------------- */

typedef struct cray_scene_info {
    uint64_t n_primitives, n_lights;
    uint64_t exact_nodes, exact_bytes;      /* binary BVH */
    uint64_t wide_nodes, wide_bytes;        /* 8-wide BVH */
    uint64_t leaf_prim_bytes;               /* leaf-ordered primitive records */
    uint64_t wide_depth;
    uint32_t width, height, max_depth, num_samples;
    double bvh_build_ms, upload_ms;
    uint64_t contact_nodes;                 /* binary-BVH node boxes a planar primitive touches (see the traversal modes above) */
    uint64_t contact_primitives;            /* primitives whose outgoing rays are checked against those boxes */
} cray_scene_info;
int cray_scene_get_info(const cray_scene*, cray_scene_info* out);

/* Host-side BVH dump for parity tests against the oracle's restatement of bvh.rs:234-336:
 * pre-order list of nodes; leaf: axis = 3, a = first index into prim_order, b = count;
 * interior: axis = split axis, a = left child index, b = right child index. */
typedef struct cray_bvh_node_dump { double min[3], max[3]; uint32_t axis, a, b, _pad; } cray_bvh_node_dump;
int cray_build_reference_bvh(const cray_scene_desc* desc, cray_bvh_node_dump** nodes, uint64_t* n_nodes,
                             uint32_t** prim_order, uint64_t* n_prims);
void cray_free(void*);

/* sizeof of the structs above in declaration order (sphere, triangle, disk, primitive, texture, image, material, light, camera,
 * scene desc, ray, hit, surface, render stats, scene info, bvh node dump): a binding checks its layout against it. */
int cray_abi_struct_sizes(uint32_t* out, int capacity);

const char* cray_last_error(void);   /* thread-local */
const char* cray_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CRAY_B200_H */
