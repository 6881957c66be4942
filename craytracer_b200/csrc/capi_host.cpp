// Host-only entry points of include/cray_b200.h: scene ingest (S0) and the parser introspection used by the
// host-logic tests that restate the reference's tests/test_parser.rs.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/cray_b200.h"
#include "cry_parser.hpp"
#include "host_scene.hpp"

namespace cray {
void set_error(const std::string& msg);
}

namespace {
thread_local uint32_t g_line = 0, g_column = 0;

int fail_parse(const cray::ParserError& e) {
    cray::set_error(e.message);
    g_line = e.has_location ? e.location.line : 0;
    g_column = e.has_location ? e.location.column : 0;
    return CRAY_E_PARSE;
}
char* dup_string(const std::string& s) {
    char* p = (char*)std::malloc(s.size() + 1);
    std::memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
}  // namespace

struct cray_host_scene {
    cray::HostScene* scene;
};

extern "C" {

void cray_last_error_location(uint32_t* line, uint32_t* column) {
    if (line) *line = g_line;
    if (column) *column = g_column;
}

int cray_host_scene_parse(const char* cry_text, const char* base_dir, cray_host_scene** out) {
    if (!cry_text || !out) { cray::set_error("bad arguments"); return CRAY_E_INVALID; }
    *out = nullptr;
    g_line = g_column = 0;
    try {
        cray::HostScene* hs = cray::build_host_scene(cry_text, base_dir ? base_dir : "");
        *out = new cray_host_scene{hs};
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    } catch (const cray::IoError& e) {
        cray::set_error(e.message);
        return CRAY_E_IO;
    } catch (const cray::UnsupportedError& e) {
        cray::set_error(e.message);
        return CRAY_E_UNSUPPORTED;
    } catch (const std::exception& e) {
        cray::set_error(e.what());
        return CRAY_E_INVALID;
    }
}

int cray_host_scene_load(const char* cry_path, const char* base_dir, cray_host_scene** out) {
    if (!cry_path || !out) { cray::set_error("bad arguments"); return CRAY_E_INVALID; }
    std::ifstream f(cry_path, std::ios::binary);
    if (!f) { cray::set_error(std::string("Error reading scene file ") + cry_path); return CRAY_E_IO; }
    std::stringstream ss;
    ss << f.rdbuf();
    return cray_host_scene_parse(ss.str().c_str(), base_dir, out);
}

const cray_scene_desc* cray_host_scene_desc(const cray_host_scene* hs) { return hs ? &hs->scene->desc : nullptr; }

void cray_host_scene_destroy(cray_host_scene* hs) {
    if (!hs) return;
    delete hs->scene;
    delete hs;
}

// number of warnings / i-th warning (unused keys, stand-in meshes, MTL problems)
uint64_t cray_host_scene_num_warnings(const cray_host_scene* hs) { return hs ? hs->scene->warnings.size() : 0; }
const char* cray_host_scene_warning(const cray_host_scene* hs, uint64_t i) { return hs && i < hs->scene->warnings.size() ? hs->scene->warnings[i].c_str() : ""; }

void cray_set_image_decoder(cray::image_decoder_fn fn) { cray::set_image_decoder(fn); }
void cray_register_standin_mesh(const char* file_name, int kind, uint64_t triangles, uint64_t seed) { cray::register_standin_mesh(file_name, kind, triangles, seed); }
void cray_clear_standin_meshes(void) { cray::clear_standin_meshes(); }

// Parser introspection: JSON of tokenize(input) / RawValue::from_tokens(tokenize(input)); caller frees with cray_free.
int cray_debug_tokenize(const char* input, char** json_out) {
    g_line = g_column = 0;
    try {
        *json_out = dup_string(cray::tokens_to_json(cray::tokenize(input)));
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    }
}
int cray_debug_parse_raw_value(const char* input, char** json_out) {
    g_line = g_column = 0;
    try {
        auto tokens = cray::tokenize(input);
        size_t pos = 0;
        auto v = cray::parse_raw_value(tokens, pos);
        *json_out = dup_string(cray::raw_value_to_json(*v));
        return CRAY_OK;
    } catch (const cray::ParserError& e) {
        return fail_parse(e);
    }
}

}  // extern "C"
