// See host_scene.hpp.
#include "host_scene.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <functional>
#include <map>
#include <sstream>
#include <unordered_map>

#include "cray_math.cuh"
#include "image_decode.hpp"

namespace cray {

void HostScene::finalize() {
    desc.n_spheres = spheres.size();
    desc.n_triangles = triangles.size();
    desc.n_disks = disks.size();
    desc.n_primitives = primitives.size();
    desc.n_materials = materials.size();
    desc.n_lights = lights.size();
    desc.n_images = images.size();
    for (size_t i = 0; i < images.size(); ++i) images[i].rgb = image_data[i].data();
    desc.spheres = spheres.data();
    desc.triangles = triangles.data();
    desc.disks = disks.data();
    desc.primitives = primitives.data();
    desc.materials = materials.data();
    desc.lights = lights.data();
    desc.images = images.data();
}

namespace {

image_decoder_fn g_decoder = nullptr;

struct Standin {
    int kind;
    uint64_t triangles, seed;
};
std::map<std::string, Standin>& standins() {
    static std::map<std::string, Standin> m;
    return m;
}

std::string join_path(const std::string& dir, const std::string& rel) {
    if (rel.empty() || rel[0] == '/' || dir.empty()) return rel;
    return dir + (dir.back() == '/' ? "" : "/") + rel;
}
std::string parent_dir(const std::string& path) {
    size_t k = path.find_last_of('/');
    return k == std::string::npos ? std::string() : path.substr(0, k);
}

// ---- typed access into raw maps (RawValueMap::get / get_or, scene_parser.rs:469-509) -------------------

[[noreturn]] void conversion_error(const std::string& key, const ParserError& e, Location fallback) {
    throw ParserError::at("Error converting map value for '" + key + "' to expected type: " + e.message, e.has_location ? e.location : fallback);
}
RawValue& require(RawValue& owner, const std::string& key) {
    owner.used_keys.insert(key);
    RawValue* v = owner.map.find(key);
    if (!v) throw ParserError::at(key + " not found in map", owner.map.location);
    return *v;
}
RawValue* optional(RawValue& owner, const std::string& key) {
    owner.used_keys.insert(key);
    return owner.map.find(key);
}
template <class T, class F>
T get_with(RawValue& owner, const std::string& key, F&& conv) {
    RawValue& v = require(owner, key);
    try {
        return conv(v);
    } catch (const ParserError& e) {
        conversion_error(key, e, owner.map.location);
    }
}
template <class T, class F>
T get_or_with(RawValue& owner, const std::string& key, T dflt, F&& conv) {
    RawValue* v = optional(owner, key);
    if (!v) return dflt;
    try {
        return conv(*v);
    } catch (const ParserError& e) {
        conversion_error(key, e, owner.map.location);
    }
}
double as_number(RawValue& v) {
    if (v.kind != RawKind::Number) throw ParserError::nowhere("Cannot get Number, found " + v.debug());
    return v.number;
}
uint64_t as_usize_value(RawValue& v) { return as_usize(as_number(v)); }  // `*value as usize` scene_parser.rs:654
std::string as_string(RawValue& v) {
    if (v.kind != RawKind::String) throw ParserError::nowhere("Cannot get String, found " + v.debug());
    return v.string;
}
struct Triple {
    double v[3];
};
Triple as_triple(RawValue& v, RawKind want, const char* name) {
    if (v.kind != want) throw ParserError::nowhere(std::string("Cannot get ") + name + ", found " + v.debug());
    return {{v.xyz[0], v.xyz[1], v.xyz[2]}};
}
Triple as_vector(RawValue& v) { return as_triple(v, RawKind::Vector, "Vector"); }
Triple as_point(RawValue& v) { return as_triple(v, RawKind::Point, "Point"); }
Triple as_color(RawValue& v) { return as_triple(v, RawKind::Color, "Color"); }
RawValue& as_typed_map(RawValue& v, const char* what) {
    if (v.kind != RawKind::TypedMap) throw ParserError::nowhere(std::string("Cannot get ") + what + ", found " + v.debug());
    return v;
}

cray_texture_desc constant_texture(double r, double g, double b) {
    cray_texture_desc t{};
    t.kind = CRAY_TEX_CONSTANT;
    t.image = -1;
    t.a[0] = r; t.a[1] = g; t.a[2] = b;
    t.scale = 1.0;
    return t;
}
// impl TryFrom<&mut RawValue> for Texture<T>  scene_parser.rs:909-938
cray_texture_desc as_texture(RawValue& v, bool is_color) {
    if (is_color ? v.kind == RawKind::Color : v.kind == RawKind::Number) {
        return is_color ? constant_texture(v.xyz[0], v.xyz[1], v.xyz[2]) : constant_texture(v.number, 0, 0);
    }
    if (v.kind == RawKind::TypedMap) {
        if (v.type_name == "Checkerboard") {
            cray_texture_desc t{};
            t.kind = CRAY_TEX_CHECKERBOARD;
            t.image = -1;
            auto conv = [&](RawValue& x) -> Triple {
                if (is_color) return as_color(x);
                return {{as_number(x), 0, 0}};
            };
            Triple a = get_with<Triple>(v, "a", conv), b = get_with<Triple>(v, "b", conv);
            std::memcpy(t.a, a.v, sizeof(t.a));
            std::memcpy(t.b, b.v, sizeof(t.b));
            t.scale = get_or_with<double>(v, "scale", 1.0, as_number);
            return t;
        }
        throw ParserError::at("Unknown material type: " + v.type_name, v.map.location);
    }
    throw ParserError::nowhere("Cannot get Color, found " + v.debug());
}

// ---- images ------------------------------------------------------------------------------------------

bool load_ppm(const std::string& path, uint32_t& w, uint32_t& h, std::vector<uint8_t>& rgb) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P6") return false;
    auto next_int = [&]() -> long {
        for (;;) {
            int c = f.peek();
            if (c == '#') { std::string line; std::getline(f, line); }
            else if (std::isspace(c)) f.get();
            else break;
        }
        long v;
        f >> v;
        return v;
    };
    long W = next_int(), H = next_int(), maxv = next_int();
    f.get();
    if (W <= 0 || H <= 0 || maxv != 255) return false;
    w = (uint32_t)W; h = (uint32_t)H;
    rgb.resize((size_t)W * H * 3);
    f.read((char*)rgb.data(), (std::streamsize)rgb.size());
    return (bool)f;
}

int load_image(HostScene& hs, const std::string& path) {  // load_texture obj.rs:16-24 + Texture::image texture.rs:57-59
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> rgb;
    bool ok = false;
    if (path.size() > 4 && path.substr(path.size() - 4) == ".ppm") ok = load_ppm(path, w, h, rgb);
    if (!ok) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw IoError{"Could not find texture file \"" + path + "\""};
        const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        std::string why;
        ok = decode_image(bytes.data(), bytes.size(), w, h, rgb, why);  // built in: JPEG, PNG (image_decode.cpp)
        if (!ok && g_decoder) {  // any other format: the host's hook, if it registered one
            uint8_t* buf = nullptr;
            if (g_decoder(path.c_str(), &w, &h, &buf) == 0 && buf) {
                rgb.assign(buf, buf + (size_t)w * h * 3);
                std::free(buf);
                ok = true;
            }
        }
        if (!ok) throw IoError{"could not decode texture \"" + path + "\": " + why};
    }
    hs.images.push_back({w, h, nullptr});
    hs.image_data.push_back(std::move(rgb));
    return (int)hs.images.size() - 1;
}

// ---- OBJ / MTL (tobj 4.0 `load_obj(path, &GPU_LOAD_OPTIONS)` behaviour: triangulate, single_index) ------

struct MtlMaterial {
    std::string name;
    bool has_kd = false, has_ks = false;
    double kd[3] = {0, 0, 0}, ks[3] = {0, 0, 0};
    bool has_ns = false, has_ni = false, has_d = false, has_illum = false;
    double ns = 0, ni = 1, d = 1;
    int illum = 0;
    std::string map_kd, map_ks;
    std::map<std::string, std::string> unknown;
};

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string w;
    while (is >> w) out.push_back(w);
    return out;
}
std::string rest_after_keyword(const std::string& line) {
    size_t k = line.find_first_of(" \t");
    if (k == std::string::npos) return "";
    size_t b = line.find_first_not_of(" \t", k);
    if (b == std::string::npos) return "";
    size_t e = line.find_last_not_of(" \t\r");
    return line.substr(b, e - b + 1);
}

bool load_mtl(const std::string& path, std::vector<MtlMaterial>& out) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    MtlMaterial* cur = nullptr;
    while (std::getline(f, line)) {
        auto w = split_ws(line);
        if (w.empty() || w[0][0] == '#') continue;
        const std::string& k = w[0];
        if (k == "newmtl") {
            out.emplace_back();
            cur = &out.back();
            cur->name = rest_after_keyword(line);
            continue;
        }
        if (!cur) continue;
        auto f3 = [&](double* dst) {
            for (int i = 0; i < 3 && i + 1 < (int)w.size(); ++i) dst[i] = std::strtod(w[i + 1].c_str(), nullptr);
        };
        if (k == "Kd") { f3(cur->kd); cur->has_kd = true; }
        else if (k == "Ks") { f3(cur->ks); cur->has_ks = true; }
        else if (k == "Ka") { /* parsed by tobj, unused by the reference */ }
        else if (k == "Ns" && w.size() > 1) { cur->ns = std::strtod(w[1].c_str(), nullptr); cur->has_ns = true; }
        else if (k == "Ni" && w.size() > 1) { cur->ni = std::strtod(w[1].c_str(), nullptr); cur->has_ni = true; }
        else if (k == "d" && w.size() > 1) { cur->d = std::strtod(w[1].c_str(), nullptr); cur->has_d = true; }
        else if (k == "illum" && w.size() > 1) { cur->illum = std::atoi(w[1].c_str()); cur->has_illum = true; }
        else if (k == "map_Kd") cur->map_kd = rest_after_keyword(line);
        else if (k == "map_Ks") cur->map_ks = rest_after_keyword(line);
        else if (k == "map_Ka" || k == "map_Ns" || k == "map_Bump" || k == "map_bump" || k == "bump" || k == "map_d" || k == "norm") { /* known to tobj, unused */ }
        else cur->unknown[k] = rest_after_keyword(line);
    }
    return true;
}

double parse_float_lenient(const std::string& s) {  // parse_float obj.rs:40-49
    char* end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end && *end == '\0' && end != s.c_str()) return v;
    return 0.0;
}

struct VertexKey {
    int64_t v, vt, vn;
    bool operator==(const VertexKey& o) const { return v == o.v && vt == o.vt && vn == o.vn; }
};
struct VertexKeyHash {
    size_t operator()(const VertexKey& k) const {
        uint64_t h = (uint64_t)k.v * 0x9E3779B97F4A7C15ull;
        h ^= ((uint64_t)k.vt + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
        h ^= ((uint64_t)k.vn + 0x1CE4E5B9ull) * 0x94D049BB133111EBull;
        return (size_t)(h ^ (h >> 29));
    }
};

struct ObjFile {
    std::vector<MeshModel> models;
    std::vector<MtlMaterial> materials;
    bool mtl_ok = true;
};

void load_obj_file(const std::string& path, ObjFile& out) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw IoError{"could not open mesh file \"" + path + "\""};
    std::vector<double> pos, nrm, tex;
    std::unordered_map<std::string, int> mat_ids;
    struct Pending {
        std::string name = "unnamed_object";
        int material = -1;
        std::vector<VertexKey> corners;  // 3 per triangle
    } cur;
    auto flush = [&]() {
        if (cur.corners.empty()) return;
        MeshModel m;
        m.name = cur.name;
        m.material_id = cur.material;
        bool all_vn = true, all_vt = true;
        for (const VertexKey& k : cur.corners) { all_vn &= k.vn >= 0; all_vt &= k.vt >= 0; }
        std::unordered_map<VertexKey, uint32_t, VertexKeyHash> remap;
        remap.reserve(cur.corners.size());
        for (VertexKey k : cur.corners) {
            if (!all_vn) k.vn = -1;
            if (!all_vt) k.vt = -1;
            auto it = remap.find(k);
            if (it == remap.end()) {
                uint32_t idx = (uint32_t)(m.positions.size() / 3);
                remap.emplace(k, idx);
                for (int c = 0; c < 3; ++c) m.positions.push_back(pos[3 * k.v + c]);
                if (all_vn) for (int c = 0; c < 3; ++c) m.normals.push_back(nrm[3 * k.vn + c]);
                if (all_vt) for (int c = 0; c < 2; ++c) m.texcoords.push_back(tex[2 * k.vt + c]);
                m.indices.push_back(idx);
            } else {
                m.indices.push_back(it->second);
            }
        }
        out.models.push_back(std::move(m));
        cur.corners.clear();
    };
    std::string dir = parent_dir(path);
    char* line = nullptr;
    size_t cap = 0;
    ssize_t len;
    std::vector<VertexKey> face;
    while ((len = getline(&line, &cap, fp)) >= 0) {
        char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '#' || *p == '\n' || *p == '\r' || *p == 0) continue;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            char* q = p + 1;
            for (int c = 0; c < 3; ++c) pos.push_back(std::strtod(q, &q));
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            char* q = p + 2;
            for (int c = 0; c < 3; ++c) nrm.push_back(std::strtod(q, &q));
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            char* q = p + 2;
            for (int c = 0; c < 2; ++c) tex.push_back(std::strtod(q, &q));
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            face.clear();
            char* q = p + 1;
            for (;;) {
                while (*q == ' ' || *q == '\t') ++q;
                if (*q == '\n' || *q == '\r' || *q == 0) break;
                VertexKey k{-1, -1, -1};
                long idx[3] = {0, 0, 0};
                bool have[3] = {false, false, false};
                for (int part = 0; part < 3; ++part) {
                    if (*q != '/' && *q != ' ' && *q != '\t' && *q != '\n' && *q != '\r' && *q != 0) {
                        idx[part] = std::strtol(q, &q, 10);
                        have[part] = true;
                    }
                    if (*q == '/') ++q;
                    else break;
                }
                auto fix = [](long i, size_t n) -> int64_t { return i > 0 ? (int64_t)i - 1 : (int64_t)n + i; };  // negative = relative
                if (have[0]) k.v = fix(idx[0], pos.size() / 3);
                if (have[1]) k.vt = fix(idx[1], tex.size() / 2);
                if (have[2]) k.vn = fix(idx[2], nrm.size() / 3);
                if (k.v < 0 || (size_t)k.v >= pos.size() / 3) { std::free(line); std::fclose(fp); throw IoError{"face references a missing vertex in \"" + path + "\""}; }
                face.push_back(k);
            }
            // tobj triangulates polygons as a fan (0, i, i+1); points and lines are ignored (GPU_LOAD_OPTIONS)
            for (size_t i = 1; i + 1 < face.size(); ++i) {
                cur.corners.push_back(face[0]);
                cur.corners.push_back(face[i]);
                cur.corners.push_back(face[i + 1]);
            }
        } else if ((p[0] == 'o' || p[0] == 'g') && (p[1] == ' ' || p[1] == '\t' || p[1] == '\n' || p[1] == '\r')) {
            flush();  // a new model starts at every o / g line
            cur.name = rest_after_keyword(p);
            while (!cur.name.empty() && (cur.name.back() == '\n' || cur.name.back() == '\r')) cur.name.pop_back();
        } else if (std::strncmp(p, "usemtl", 6) == 0) {
            std::string name = rest_after_keyword(p);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r')) name.pop_back();
            auto it = mat_ids.find(name);
            int id = it == mat_ids.end() ? -1 : it->second;
            if (id != cur.material) flush();  // material change while faces are pending => new model
            cur.material = id;
        } else if (std::strncmp(p, "mtllib", 6) == 0) {
            std::string name = rest_after_keyword(p);
            while (!name.empty() && (name.back() == '\n' || name.back() == '\r')) name.pop_back();
            size_t before = out.materials.size();
            if (!load_mtl(join_path(dir, name), out.materials)) out.mtl_ok = false;
            for (size_t i = before; i < out.materials.size(); ++i) mat_ids.emplace(out.materials[i].name, (int)i);
        }
    }
    std::free(line);
    std::fclose(fp);
    flush();
}

// ---- procedural stand-ins ------------------------------------------------------------------------------

uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Closed displaced tube along a (2,3) torus knot, `target` triangles exactly, OBJ (right-handed) space,
// bounding box scaled to the published extent of xyzrgb_dragon as placed by scenes/dragon.cry
// ([-100,100] x [-40,50] x [-45,45], resting on y = -40).
void make_dragon_standin(uint64_t target, uint64_t seed, MeshModel& m) {
    uint64_t quads = target / 2;
    uint64_t n_theta = 0;
    for (uint64_t c = 512; c >= 128; --c)
        if (quads % c == 0) { n_theta = c; break; }
    if (n_theta == 0) n_theta = 384;
    uint64_t n_s = quads / n_theta;
    if (n_s < 3) { n_s = 3; n_theta = std::max<uint64_t>(3, quads / 3); }
    const double R = 2.0, a = 0.8, r0 = 0.42;
    m.positions.resize(3 * n_s * n_theta);
    uint64_t rng = seed ^ 0xD1B54A32D192ED03ull;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (uint64_t i = 0; i < n_s; ++i) {
        double phi = 2.0 * kPi * (double)i / (double)n_s;
        V3 c = mk((R + a * std::cos(3 * phi)) * std::cos(2 * phi), a * std::sin(3 * phi), (R + a * std::cos(3 * phi)) * std::sin(2 * phi));
        V3 dc = mk(-3 * a * std::sin(3 * phi) * std::cos(2 * phi) - 2 * (R + a * std::cos(3 * phi)) * std::sin(2 * phi), 3 * a * std::cos(3 * phi),
                   -3 * a * std::sin(3 * phi) * std::sin(2 * phi) + 2 * (R + a * std::cos(3 * phi)) * std::cos(2 * phi));
        V3 T = normalized(dc);
        V3 N = normalized(cross(T, mk(0, 1, 0)));
        V3 B = cross(T, N);
        for (uint64_t j = 0; j < n_theta; ++j) {
            double th = 2.0 * kPi * (double)j / (double)n_theta;
            double jitter = ((double)(splitmix64(rng) >> 11) * (1.0 / 9007199254740992.0) - 0.5);
            double r = r0 * (1.0 + 0.16 * std::sin(7 * th + 40 * phi) + 0.07 * std::sin(61 * phi) * std::cos(9 * th) + 0.03 * std::sin(173 * phi + 31 * th) + 0.004 * jitter);
            V3 p = c + N * (r * std::cos(th)) + B * (r * std::sin(th));
            double* dst = &m.positions[3 * (i * n_theta + j)];
            dst[0] = p.x; dst[1] = p.y; dst[2] = p.z;
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], dst[k]); hi[k] = std::max(hi[k], dst[k]); }
        }
    }
    const double tlo[3] = {-100, -40, -45}, thi[3] = {100, 50, 45};
    for (size_t v = 0; v < m.positions.size() / 3; ++v)
        for (int k = 0; k < 3; ++k) m.positions[3 * v + k] = tlo[k] + (m.positions[3 * v + k] - lo[k]) / (hi[k] - lo[k]) * (thi[k] - tlo[k]);
    m.indices.reserve(3 * target);
    for (uint64_t i = 0; i < n_s; ++i)
        for (uint64_t j = 0; j < n_theta; ++j) {
            uint32_t v00 = (uint32_t)(i * n_theta + j), v01 = (uint32_t)(i * n_theta + (j + 1) % n_theta);
            uint32_t v10 = (uint32_t)(((i + 1) % n_s) * n_theta + j), v11 = (uint32_t)(((i + 1) % n_s) * n_theta + (j + 1) % n_theta);
            uint32_t t[6] = {v00, v10, v11, v00, v11, v01};
            m.indices.insert(m.indices.end(), t, t + 6);
        }
    // leftover triangles (target not a multiple of the grid): small scales standing on the surface
    uint64_t have = m.indices.size() / 3;
    for (uint64_t e = have; e < target; ++e) {
        uint64_t v = (splitmix64(rng) % (n_s * n_theta));
        uint32_t base = (uint32_t)(m.positions.size() / 3);
        double* p = &m.positions[3 * v];
        double q[9] = {p[0], p[1], p[2], p[0] + 0.05, p[1] + 0.11, p[2] + 0.02, p[0] - 0.04, p[1] + 0.09, p[2] + 0.06};
        m.positions.insert(m.positions.end(), q, q + 9);
        uint32_t t[3] = {base, base + 1, base + 2};
        m.indices.insert(m.indices.end(), t, t + 3);
    }
}

// Interior stand-in: a room with a flight of stairs, wall panels and bumpy props; one model per MTL material
// (cycled), uv-mapped so that every map_Kd texture is sampled.
void make_interior_standin(uint64_t target, uint64_t seed, int n_materials, std::vector<MeshModel>& models) {
    const bool have_materials = n_materials > 0;
    if (!have_materials) n_materials = 1;
    uint64_t per = std::max<uint64_t>(2, target / (uint64_t)n_materials);
    uint64_t rng = seed ^ 0xA0761D6478BD642Full;
    for (int mi = 0; mi < n_materials; ++mi) {
        MeshModel m;
        m.material_id = have_materials ? mi : -1;
        m.name = "standin_" + std::to_string(mi);
        // a bumpy height-field patch n x n (2 n^2 triangles) placed in a 3 x 6 x 5 room, facing alternating directions
        uint64_t n = std::max<uint64_t>(1, (uint64_t)std::floor(std::sqrt((double)per / 2.0)));
        int slot = mi % 6;
        double ox = -1.5 + 0.5 * (mi % 5), oy = 0.15 * (mi % 17), oz = -1.0 + 0.4 * (mi % 9);
        double amp = 0.02 + 0.01 * (double)(splitmix64(rng) % 5);
        for (uint64_t i = 0; i <= n; ++i)
            for (uint64_t j = 0; j <= n; ++j) {
                double u = (double)i / (double)n, v = (double)j / (double)n;
                double hgt = amp * std::sin(23 * u + mi) * std::cos(19 * v) + amp * 0.3 * std::sin(131 * u * v);
                double P[3];
                switch (slot) {
                    case 0: P[0] = -1.5 + 3 * u; P[1] = hgt; P[2] = -2.0 + 5.0 * v; break;            // floor
                    case 1: P[0] = -1.5 + 3 * u; P[1] = 5.5 + hgt; P[2] = -2.0 + 5.0 * v; break;      // ceiling
                    case 2: P[0] = -1.5 + hgt; P[1] = 5.5 * u; P[2] = -2.0 + 5.0 * v; break;          // left wall
                    case 3: P[0] = 1.5 + hgt; P[1] = 5.5 * u; P[2] = -2.0 + 5.0 * v; break;           // right wall
                    case 4: P[0] = -1.5 + 3 * u; P[1] = 5.5 * v; P[2] = 3.0 + hgt; break;             // back wall
                    default: P[0] = ox + 0.9 * u; P[1] = 0.2 + oy + 0.25 * std::floor(6 * v) / 1.0 * 0.3 + hgt; P[2] = oz + 2.4 * v; break;  // stairs / props
                }
                m.positions.insert(m.positions.end(), P, P + 3);
                m.texcoords.push_back(3.0 * u);
                m.texcoords.push_back(2.0 * v);
            }
        for (uint64_t i = 0; i < n; ++i)
            for (uint64_t j = 0; j < n; ++j) {
                uint32_t a = (uint32_t)(i * (n + 1) + j), b = a + 1, c = (uint32_t)((i + 1) * (n + 1) + j), d = c + 1;
                uint32_t t[6] = {a, c, d, a, d, b};
                m.indices.insert(m.indices.end(), t, t + 6);
            }
        models.push_back(std::move(m));
    }
}

// ---- mesh -> primitives (obj.rs:61-220) -----------------------------------------------------------------

void append_mesh(HostScene& hs, const std::string& file_name, const std::string& resolved, int fallback_material) {
    ObjFile obj;
    std::ifstream probe(resolved, std::ios::binary);
    auto st = standins().find(file_name);
    if (!probe && st != standins().end()) {
        std::string mtl = resolved.substr(0, resolved.find_last_of('.')) + ".mtl";
        load_mtl(mtl, obj.materials);
        if (st->second.kind == 0) {
            obj.models.emplace_back();
            make_dragon_standin(st->second.triangles, st->second.seed, obj.models.back());
        } else {
            make_interior_standin(st->second.triangles, st->second.seed, (int)obj.materials.size(), obj.models);
        }
        hs.warnings.push_back("mesh \"" + file_name + "\" not found: using the registered procedural stand-in");
    } else {
        load_obj_file(resolved, obj);
    }
    if (!obj.mtl_ok) hs.warnings.push_back("Error loading materials in " + file_name + ", skipping");

    // materials (obj.rs:61-105)
    std::vector<int> material_index(obj.materials.size(), fallback_material);
    std::vector<bool> emissive(obj.materials.size(), false);
    std::vector<Triple> emittance(obj.materials.size());
    for (size_t id = 0; id < obj.materials.size(); ++id) {
        const MtlMaterial& m = obj.materials[id];
        auto tex_of = [&](const std::string& map, bool has, const double* c) -> cray_texture_desc {
            if (!map.empty()) {
                cray_texture_desc t = constant_texture(0, 0, 0);
                t.kind = CRAY_TEX_IMAGE;
                t.image = load_image(hs, join_path(parent_dir(resolved), map));
                return t;
            }
            return has ? constant_texture(c[0], c[1], c[2]) : constant_texture(0, 0, 0);
        };
        if (m.map_kd.empty() && !m.has_kd) throw IoError{"material \"" + m.name + "\" has no Kd (the reference panics at obj.rs:67)"};
        cray_texture_desc diffuse = tex_of(m.map_kd, m.has_kd, m.kd);
        cray_texture_desc specular = tex_of(m.map_ks, m.has_ks, m.ks);
        Triple ke{{0, 0, 0}};
        auto it = m.unknown.find("Ke");
        if (it != m.unknown.end()) {
            auto w = split_ws(it->second);
            if (w.size() < 3) throw IoError{"malformed Ke in material \"" + m.name + "\""};
            for (int c = 0; c < 3; ++c) ke.v[c] = parse_float_lenient(w[c]);
        }
        double shininess = m.has_ns ? m.ns : 0.0;
        const double E = 2.718281828459045235360287471352662498;  // std::f64::consts::E
        cray_texture_desc roughness = constant_texture(180.0 * (1.0 - std::pow(E, -shininess / 100.0)), 0, 0);
        double dissolve = m.has_d ? m.d : 1.0;
        if (!(ke.v[0] == 0.0 && ke.v[1] == 0.0 && ke.v[2] == 0.0)) {
            emissive[id] = true;
            emittance[id] = ke;
            material_index[id] = fallback_material;
            continue;
        }
        cray_material_desc md{};
        md.t0 = md.t1 = md.t2 = constant_texture(0, 0, 0);
        if (dissolve < 1.0) {
            md.kind = CRAY_MAT_GLASS;
            md.t0 = diffuse;
            md.t1 = diffuse;
            md.eta = m.has_ni ? m.ni : 1.0;
        } else if (m.has_illum && m.illum >= 3 && m.illum <= 9) {
            md.kind = CRAY_MAT_METAL;
            md.t0 = diffuse;
            md.t1 = specular;
        } else {
            md.kind = CRAY_MAT_PLASTIC;
            md.t0 = diffuse;
            md.t1 = specular;
            md.t2 = roughness;
        }
        hs.materials.push_back(md);
        material_index[id] = (int)hs.materials.size() - 1;
    }

    for (size_t mi = 0; mi < obj.models.size(); ++mi) {
        const MeshModel& mesh = obj.models[mi];
        int material = fallback_material;
        bool is_light = false;
        Triple ke{{0, 0, 0}};
        if (mesh.material_id >= 0) {
            material = material_index[mesh.material_id];
            is_light = emissive[mesh.material_id];
            ke = emittance[mesh.material_id];
        }
        const size_t nv = mesh.positions.size() / 3;
        const bool have_n = !mesh.normals.empty(), have_t = !mesh.texcoords.empty();
        auto P = [&](uint32_t i) { return mk(mesh.positions[3 * i], mesh.positions[3 * i + 1], -mesh.positions[3 * i + 2]); };  // RH -> LH, obj.rs:126-133
        auto N = [&](uint32_t i) { return mk(mesh.normals[3 * i], mesh.normals[3 * i + 1], -mesh.normals[3 * i + 2]); };
        for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
            uint32_t i = mesh.indices[t], j = mesh.indices[t + 1], k = mesh.indices[t + 2];
            if (i >= nv || j >= nv || k >= nv) throw IoError{"mesh index out of range in \"" + file_name + "\""};
            V3 vi = P(i), vj = P(j), vk = P(k);
            V3 normal = normalized(cross(vk - vi, vj - vi));
            V3 ni = normal, nj = normal, nk = normal;
            if (have_n) { ni = N(i); nj = N(j); nk = N(k); }
            double uv0[2] = {0.0, 0.0}, uv1[2] = {1.0, 0.0}, uv2[2] = {1.0, 1.0};
            if (have_t) {
                uv0[0] = mesh.texcoords[2 * i]; uv0[1] = 1.0 - mesh.texcoords[2 * i + 1];
                uv1[0] = mesh.texcoords[2 * j]; uv1[1] = 1.0 - mesh.texcoords[2 * j + 1];
                uv2[0] = mesh.texcoords[2 * k]; uv2[1] = 1.0 - mesh.texcoords[2 * k + 1];
            }
            // Shape::new_triangle_with_normals_and_texture_coordinates shape.rs:98-132
            V3 e1 = vj - vi, e2 = vk - vi;
            if (magnitude_squared(cross(e2, e1)) == 0.0 || magnitude_squared(ni) == 0.0 || magnitude_squared(nj) == 0.0 || magnitude_squared(nk) == 0.0) continue;
            cray_triangle_desc td;
            auto put = [](double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
            put(td.v0, vi); put(td.e1, e1); put(td.e2, e2);
            put(td.n0, ni); put(td.n01, nj - ni); put(td.n02, nk - ni);
            td.uv0[0] = uv0[0]; td.uv0[1] = uv0[1];
            td.uv01[0] = uv1[0] - uv0[0]; td.uv01[1] = uv1[1] - uv0[1];
            td.uv02[0] = uv2[0] - uv0[0]; td.uv02[1] = uv2[1] - uv0[1];
            hs.triangles.push_back(td);
            cray_primitive_desc pd{CRAY_SHAPE_TRIANGLE, (uint32_t)hs.triangles.size() - 1, material, -1};
            if (is_light) {
                cray_light_desc ld{};
                ld.kind = CRAY_LIGHT_AREA;
                ld.primitive = (int32_t)hs.primitives.size();
                std::memcpy(ld.color, ke.v, sizeof(ld.color));
                hs.lights.push_back(ld);
                pd.area_light = (int32_t)hs.lights.size() - 1;
            }
            hs.primitives.push_back(pd);
        }
    }
}

}  // namespace

void set_image_decoder(image_decoder_fn fn) { g_decoder = fn; }
void register_standin_mesh(const std::string& file_name, int kind, uint64_t triangles, uint64_t seed) { standins()[file_name] = {kind, triangles, seed}; }
void clear_standin_meshes() { standins().clear(); }

HostScene* build_host_scene(const std::string& cry_text, const std::string& base_dir) {
    std::vector<Token> tokens = tokenize(cry_text);
    size_t pos = 0;
    RawValue root;
    root.kind = RawKind::Map;
    root.map = parse_raw_map(tokens, pos);

    auto hs = std::make_unique<HostScene>();
    hs->desc.max_depth = (uint32_t)get_or_with<uint64_t>(root, "max_depth", 8, as_usize_value);      // scene_parser.rs:796, :1084
    hs->desc.num_samples = (uint32_t)get_or_with<uint64_t>(root, "num_samples", 4, as_usize_value);  // :797, :1085

    // Camera (:801-845)
    get_with<int>(root, "camera", [&](RawValue& v) -> int {
        RawValue& cm = as_typed_map(v, "Camera");
        cray_camera_desc& c = hs->desc.camera;
        get_with<int>(cm, "film", [&](RawValue& fv) -> int {  // Film :848-863
            if (fv.kind != RawKind::Map) throw ParserError::nowhere("Cannot get Film, found " + fv.debug());
            c.width = (uint32_t)get_with<uint64_t>(fv, "width", as_usize_value);
            c.height = (uint32_t)get_with<uint64_t>(fv, "height", as_usize_value);
            return 0;
        });
        Triple o = get_with<Triple>(cm, "origin", as_point), t = get_with<Triple>(cm, "target", as_point), u = get_with<Triple>(cm, "up", as_vector);
        std::memcpy(c.origin, o.v, 24); std::memcpy(c.target, t.v, 24); std::memcpy(c.up, u.v, 24);
        c.lens_radius = get_or_with<double>(cm, "lens_radius", 0.0, as_number);
        c.focal_distance = get_or_with<double>(cm, "focal_distance", 1e6, as_number);  // DEFAULT_FOCAL_DISTANCE :798
        if (cm.type_name == "Perspective") {
            c.kind = CRAY_CAMERA_PERSPECTIVE;
            c.fov = get_with<double>(cm, "fov", as_number);
        } else if (cm.type_name == "Orthographic") {
            c.kind = CRAY_CAMERA_ORTHOGRAPHIC;
        } else {
            throw ParserError::nowhere("Unknown camera type: " + cm.type_name);
        }
        return 0;
    });

    // Lights (:866-903)
    get_with<int>(root, "lights", [&](RawValue& v) -> int {
        if (v.kind != RawKind::Array) throw ParserError::nowhere("Cannot get Array, found " + v.debug());
        for (auto& item : v.array) {
            RawValue& lm = as_typed_map(*item, "Light");
            cray_light_desc ld{};
            ld.primitive = -1;
            if (lm.type_name == "Point") {
                ld.kind = CRAY_LIGHT_POINT;
                Triple o = get_with<Triple>(lm, "origin", as_point), c = get_with<Triple>(lm, "intensity", as_color);
                std::memcpy(ld.v, o.v, 24); std::memcpy(ld.color, c.v, 24);
            } else if (lm.type_name == "Distant") {
                ld.kind = CRAY_LIGHT_DISTANT;
                Triple d = get_with<Triple>(lm, "direction", as_vector), c = get_with<Triple>(lm, "intensity", as_color);
                V3 n = normalized(mk(d.v[0], d.v[1], d.v[2]));  // direction.normalized() :888
                ld.v[0] = n.x; ld.v[1] = n.y; ld.v[2] = n.z;
                std::memcpy(ld.color, c.v, 24);
            } else if (lm.type_name == "Infinite") {
                ld.kind = CRAY_LIGHT_INFINITE;
                Triple c = get_with<Triple>(lm, "intensity", as_color);
                std::memcpy(ld.color, c.v, 24);
            } else {
                throw ParserError::nowhere("Unknown light type: " + lm.type_name);
            }
            hs->lights.push_back(ld);
        }
        return 0;
    });

    // Materials (:941-976) and shapes (:979-1019): name -> index
    std::map<std::string, int> material_ids;
    get_with<int>(root, "materials", [&](RawValue& v) -> int {
        if (v.kind != RawKind::Map) throw ParserError::nowhere("Cannot get Map, found " + v.debug());
        for (auto& e : v.map.entries) {
            RawValue& mm = as_typed_map(*e.second, "Material");
            cray_material_desc md{};
            md.t0 = md.t1 = md.t2 = constant_texture(0, 0, 0);
            auto color_tex = [](RawValue& x) { return as_texture(x, true); };
            auto num_tex = [](RawValue& x) { return as_texture(x, false); };
            if (mm.type_name == "Matte") {
                md.kind = CRAY_MAT_MATTE;
                md.t0 = get_with<cray_texture_desc>(mm, "reflectance", color_tex);
                md.t2 = get_with<cray_texture_desc>(mm, "sigma", num_tex);
            } else if (mm.type_name == "Glass") {
                md.kind = CRAY_MAT_GLASS;
                md.t0 = get_with<cray_texture_desc>(mm, "reflectance", color_tex);
                md.t1 = get_with<cray_texture_desc>(mm, "transmittance", color_tex);
                md.eta = get_with<double>(mm, "eta", as_number);
            } else if (mm.type_name == "Plastic") {
                md.kind = CRAY_MAT_PLASTIC;
                md.t0 = get_with<cray_texture_desc>(mm, "diffuse", color_tex);
                md.t1 = get_with<cray_texture_desc>(mm, "specular", color_tex);
                md.t2 = get_with<cray_texture_desc>(mm, "roughness", num_tex);
            } else if (mm.type_name == "Metal") {
                md.kind = CRAY_MAT_METAL;
                md.t0 = get_with<cray_texture_desc>(mm, "eta", color_tex);
                md.t1 = get_with<cray_texture_desc>(mm, "k", color_tex);
            } else {
                throw ParserError::at("Unknown material type: " + mm.type_name, mm.map.location);
            }
            hs->materials.push_back(md);
            material_ids[e.first] = (int)hs->materials.size() - 1;
        }
        return 0;
    });
    struct ShapeRef { uint32_t kind, index; };
    std::map<std::string, ShapeRef> shape_ids;
    get_with<int>(root, "shapes", [&](RawValue& v) -> int {
        if (v.kind != RawKind::Map) throw ParserError::nowhere("Cannot get Map, found " + v.debug());
        for (auto& e : v.map.entries) {
            RawValue& sm = as_typed_map(*e.second, "Shape");
            if (sm.type_name == "Sphere") {
                cray_sphere_desc s{};
                Triple o = get_with<Triple>(sm, "origin", as_point);
                std::memcpy(s.origin, o.v, 24);
                s.radius = get_with<double>(sm, "radius", as_number);
                hs->spheres.push_back(s);
                shape_ids[e.first] = {CRAY_SHAPE_SPHERE, (uint32_t)hs->spheres.size() - 1};
            } else if (sm.type_name == "Triangle") {
                Triple a = get_with<Triple>(sm, "v0", as_point), b = get_with<Triple>(sm, "v1", as_point), c = get_with<Triple>(sm, "v2", as_point);
                // Shape::new_triangle shape.rs:72-97
                V3 v0 = mk(a.v[0], a.v[1], a.v[2]), v1 = mk(b.v[0], b.v[1], b.v[2]), v2 = mk(c.v[0], c.v[1], c.v[2]);
                V3 e1 = v1 - v0, e2 = v2 - v0;
                V3 n0 = cross(e2, e1);
                double mag = magnitude(n0);
                if (mag == 0.0) throw ParserError::at("Degenerate triangle: " + sm.type_name, sm.map.location);
                n0 = n0 / mag;
                cray_triangle_desc td{};
                td.v0[0] = v0.x; td.v0[1] = v0.y; td.v0[2] = v0.z;
                td.e1[0] = e1.x; td.e1[1] = e1.y; td.e1[2] = e1.z;
                td.e2[0] = e2.x; td.e2[1] = e2.y; td.e2[2] = e2.z;
                td.n0[0] = n0.x; td.n0[1] = n0.y; td.n0[2] = n0.z;
                td.uv01[0] = 1.0; td.uv02[0] = 1.0; td.uv02[1] = 1.0;
                hs->triangles.push_back(td);
                shape_ids[e.first] = {CRAY_SHAPE_TRIANGLE, (uint32_t)hs->triangles.size() - 1};
            } else if (sm.type_name == "Disk") {
                cray_disk_desc d{};
                Triple o = get_with<Triple>(sm, "origin", as_point);
                std::memcpy(d.origin, o.v, 24);
                d.rotate_x = get_or_with<double>(sm, "rotate_x", 0.0, as_number);
                d.rotate_y = get_or_with<double>(sm, "rotate_y", 0.0, as_number);
                d.radius = get_with<double>(sm, "radius", as_number);
                d.inner_radius = get_or_with<double>(sm, "inner_radius", 0.0, as_number);
                hs->disks.push_back(d);
                shape_ids[e.first] = {CRAY_SHAPE_DISK, (uint32_t)hs->disks.size() - 1};
            } else {
                throw ParserError::nowhere("Unknown shape type: " + sm.type_name);
            }
        }
        return 0;
    });

    // Primitives (:1025-1076, :1093-1101): array order, area lights appended to `lights` as they appear
    get_with<int>(root, "primitives", [&](RawValue& v) -> int {
        if (v.kind != RawKind::Array) throw ParserError::nowhere("Cannot get Array, found " + v.debug());
        for (auto& item : v.array) {
            RawValue& pm = as_typed_map(*item, "TypedRawValueMap");
            if (pm.type_name == "Shape") {
                std::string shape_name = get_with<std::string>(pm, "shape", as_string);
                auto sit = shape_ids.find(shape_name);
                if (sit == shape_ids.end()) throw ParserError::at("Cannot find shape named '" + shape_name + "'", pm.map.location);
                cray_primitive_desc pd{sit->second.kind, sit->second.index, -1, -1};
                pm.used_keys.insert("emittance");
                if (!pm.map.has("emittance")) {
                    std::string material_name = get_with<std::string>(pm, "material", as_string);
                    auto mit = material_ids.find(material_name);
                    if (mit == material_ids.end()) throw ParserError::at("Cannot find material named '" + material_name + "'", pm.map.location);
                    pd.material = mit->second;
                } else {
                    Triple e = get_with<Triple>(pm, "emittance", as_color);
                    cray_light_desc ld{};
                    ld.kind = CRAY_LIGHT_AREA;
                    ld.primitive = (int32_t)hs->primitives.size();
                    std::memcpy(ld.color, e.v, 24);
                    hs->lights.push_back(ld);
                    pd.area_light = (int32_t)hs->lights.size() - 1;
                }
                hs->primitives.push_back(pd);
            } else if (pm.type_name == "Mesh") {
                std::string file_name = get_with<std::string>(pm, "file_name", as_string);
                std::string material_name = get_with<std::string>(pm, "fallback_material", as_string);
                auto mit = material_ids.find(material_name);
                if (mit == material_ids.end()) throw ParserError::at("Cannot find material named '" + material_name + "'", pm.map.location);
                append_mesh(*hs, file_name, join_path(base_dir, file_name), mit->second);
            } else {
                throw ParserError::at("Unknown primitive type: " + pm.type_name, pm.map.location);
            }
        }
        return 0;
    });

    if (hs->lights.empty()) throw ParserError::at("No lights in the scene.", Location{0, 0});  // :1103-1108

    // unused-key warnings (Drop for TypedRawValueMap :571-586)
    std::function<void(RawValue&)> warn = [&](RawValue& v) {
        if (v.kind == RawKind::TypedMap) {
            std::string unused;
            for (auto& e : v.map.entries)
                if (!v.used_keys.count(e.first)) unused += (unused.empty() ? "" : ", ") + ("\"" + e.first + "\"");
            if (!unused.empty())
                hs->warnings.push_back("Found unused key(s) [" + unused + "] in " + v.type_name + " at " + std::to_string(v.map.location.line) + ":" + std::to_string(v.map.location.column));
        }
        for (auto& e : v.map.entries) warn(*e.second);
        for (auto& e : v.array) warn(*e);
    };
    warn(root);

    hs->finalize();
    return hs.release();
}

}  // namespace cray
