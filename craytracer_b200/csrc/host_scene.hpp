// Host-side scene ingest: typed conversion of a parsed .cry file into the flat cray_scene_desc,
// OBJ/MTL mesh loading and the procedural stand-in meshes for assets that are not redistributable.
// Mirrors src/scene_parser.rs:775-1117 and src/obj.rs:26-220 of the reference (same primitive
// order, RH->LH flips, MTL -> material mapping, Ke area lights).
#pragma once
#include <string>
#include <vector>
#include "../../include/cray_b200.h"
#include "cry_parser.hpp"

namespace cray {

struct HostScene {
    std::vector<cray_sphere_desc> spheres;
    std::vector<cray_triangle_desc> triangles;
    std::vector<cray_disk_desc> disks;
    std::vector<cray_primitive_desc> primitives;
    std::vector<cray_material_desc> materials;
    std::vector<cray_light_desc> lights;
    std::vector<cray_image_desc> images;
    std::vector<std::vector<uint8_t>> image_data;
    std::vector<std::string> warnings;
    cray_scene_desc desc{};
    void finalize();  // point desc at the vectors
};

struct IoError {
    std::string message;
};
struct UnsupportedError {
    std::string message;
};

// parse_scene scene_parser.rs:1078-1117.  Throws ParserError / IoError / UnsupportedError.
HostScene* build_host_scene(const std::string& cry_text, const std::string& base_dir);

// Indexed triangle mesh in OBJ conventions (right-handed, v up-flipped later by the loader).
struct MeshModel {
    std::string name;
    int material_id = -1;             // into the MTL material list, -1 = fallback
    std::vector<double> positions;    // 3 per vertex (single-index)
    std::vector<double> normals;      // 3 per vertex or empty
    std::vector<double> texcoords;    // 2 per vertex or empty
    std::vector<uint32_t> indices;    // 3 per triangle
};

// Image decoder hook for texture formats other than binary PPM (set from the Python host, which has PIL).
// Must return 0 and a malloc()ed RGB8 buffer on success.
typedef int (*image_decoder_fn)(const char* path, uint32_t* width, uint32_t* height, uint8_t** rgb);
void set_image_decoder(image_decoder_fn fn);

// Procedural stand-in for a mesh file that is missing from the tree (see DESIGN.md "assets").
// kind 0: "dragon" -- closed displaced tube, exactly `triangles` triangles, no vertex normals / uvs.
// kind 1: "interior" -- stairs/rooms with uvs cycling through the MTL materials next to `file_name`.
void register_standin_mesh(const std::string& file_name, int kind, uint64_t triangles, uint64_t seed);
void clear_standin_meshes();

}  // namespace cray
