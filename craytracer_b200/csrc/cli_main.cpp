// cray_b200: the reference's command line (struct Cli / main, src/bin/craytracer.rs:321-374) over libcray_b200.
//
//   cray_b200 --scene scenes/dragon.cry [--output out.exr] [--seed 0] [--preview]
//             [--spp N] [--mode fast|exact|f32] [--gpus N] [--base-dir DIR]
//
// The first four flags are the reference's (`--preview` is accepted and ignored: the minifb window is out of scope, SURVEY
// section 2).  Everything goes through the C ABI of include/cray_b200.h -- this file is also the worked example of a host
// that binds it.  Log lines mimic env_logger's "[INFO] ..." on stderr; a parse error is reported as
// "<message> at <file>:<line>:<column>" and, like the reference's main, is not a failing exit status.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "../../include/cray_b200.h"

namespace {

struct Args {
    std::string scene, output = "out.exr", mode = "fast", base_dir;
    uint64_t seed = 0;
    long spp = -1;
    long samples_per_call = 0;   // 0: as many as one render call takes
    int gpus = 1;
    bool preview = false;
};

void usage(FILE* f) {
    std::fputs("usage: cray_b200 --scene <file.cry> [--output out.exr] [--seed N] [--preview]\n"
               "                 [--spp N] [--mode fast|exact|f32] [--gpus N] [--base-dir DIR] [--samples-per-call N]\n", f);
}

bool parse_args(int argc, char** argv, Args& a) {
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        auto value = [&](const char* name) -> const char* {
            if (i + 1 >= argc) { std::fprintf(stderr, "error: %s needs a value\n", name); return nullptr; }
            return argv[++i];
        };
        const char* v = nullptr;
        if (k == "--help" || k == "-h") { usage(stdout); std::exit(0); }
        else if (k == "--preview") a.preview = true;
        else if (k == "--scene" || k == "-s") { if (!(v = value("--scene"))) return false; a.scene = v; }
        else if (k == "--output") { if (!(v = value("--output"))) return false; a.output = v; }
        else if (k == "--seed") { if (!(v = value("--seed"))) return false; a.seed = std::strtoull(v, nullptr, 10); }
        else if (k == "--spp") { if (!(v = value("--spp"))) return false; a.spp = std::strtol(v, nullptr, 10); }
        else if (k == "--mode") { if (!(v = value("--mode"))) return false; a.mode = v; }
        else if (k == "--gpus") { if (!(v = value("--gpus"))) return false; a.gpus = std::atoi(v); }
        else if (k == "--base-dir") { if (!(v = value("--base-dir"))) return false; a.base_dir = v; }
        else if (k == "--samples-per-call") { if (!(v = value("--samples-per-call"))) return false; a.samples_per_call = std::strtol(v, nullptr, 10); }
        else { std::fprintf(stderr, "error: unexpected argument '%s'\n", k.c_str()); return false; }
    }
    if (a.scene.empty()) { std::fputs("error: the following required arguments were not provided: --scene <SCENE>\n", stderr); return false; }
    if (a.mode != "fast" && a.mode != "exact" && a.mode != "f32") { std::fputs("error: --mode is one of fast, exact, f32\n", stderr); return false; }
    if (a.gpus < 1 || a.gpus > 64) { std::fputs("error: --gpus out of range\n", stderr); return false; }
    return true;
}

double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

int main(int argc, char** argv) {
    Args args;
    if (!parse_args(argc, argv, args)) { usage(stderr); return 2; }
    const auto start = std::chrono::steady_clock::now();

    // the two meshes the reference's repository does not ship fall back to the documented stand-ins (with a warning)
    cray_register_standin_mesh("objs/xyzrgb_dragon.obj", 0, 7219045ull, 0);
    cray_register_standin_mesh("objs/staircase/staircase.obj", 1, 1500000ull, 0);

    // Mesh paths are relative to the directory the reference is run from (its repository root).  Without --base-dir: the first of
    // the current directory, the scene file's parent directory and the packaged assets (<parent>/assets) that has an objs/
    // directory -- the same search as `python -m craytracer_b200`.
    std::string base = args.base_dir;
    if (base.empty()) {
        const size_t slash = args.scene.find_last_of('/');
        const std::string dir = slash == std::string::npos ? "." : args.scene.substr(0, slash);
        const std::string candidates[] = {".", dir + "/..", dir, dir + "/../assets", "assets"};
        base = candidates[1];
        for (const std::string& c : candidates) {
            struct stat st{};
            if (::stat((c + "/objs").c_str(), &st) == 0 && S_ISDIR(st.st_mode)) { base = c; break; }
        }
    }
    cray_host_scene* hs = nullptr;
    int rc = cray_host_scene_load(args.scene.c_str(), base.c_str(), &hs);
    if (rc == CRAY_E_PARSE) {  // craytracer.rs:346-355: logged, exit status 0
        uint32_t line = 0, column = 0;
        cray_last_error_location(&line, &column);
        if (line) std::fprintf(stderr, "[ERROR] %s at %s:%u:%u\n", cray_last_error(), args.scene.c_str(), line, column);
        else std::fprintf(stderr, "[ERROR] %s in %s\n", cray_last_error(), args.scene.c_str());
        return 0;
    }
    if (rc != CRAY_OK) { std::fprintf(stderr, "[ERROR] %s\n", cray_last_error()); return 1; }
    for (uint64_t i = 0; i < cray_host_scene_num_warnings(hs); ++i) std::fprintf(stderr, "[WARN] %s\n", cray_host_scene_warning(hs, i));
    const cray_scene_desc* desc = cray_host_scene_desc(hs);

    const int mode = args.mode == "exact" ? CRAY_TRAVERSE_EXACT : (args.mode == "f32" ? CRAY_TRAVERSE_F32 : CRAY_TRAVERSE_FAST);
    const uint32_t build = CRAY_BUILD_EXACT | CRAY_BUILD_FAST | (mode == CRAY_TRAVERSE_F32 ? CRAY_BUILD_F32 : 0u);
    std::vector<int> devices(args.gpus);
    for (int k = 0; k < args.gpus; ++k) devices[k] = k;
    std::vector<cray_scene*> scenes(args.gpus, nullptr);
    rc = cray_scene_create_multi(desc, devices.data(), args.gpus, build, scenes.data());
    if (rc != CRAY_OK) { std::fprintf(stderr, "[ERROR] %s\n", cray_last_error()); cray_host_scene_destroy(hs); return 1; }
    std::fprintf(stderr, "[INFO] Scene constructed in %.3fs\n", seconds_since(start));

    const uint32_t width = desc->camera.width, height = desc->camera.height;
    const uint32_t spp = args.spp >= 0 ? (uint32_t)args.spp : desc->num_samples;
    std::vector<float> pixels((size_t)width * height * 3, 0.0f);
    cray_render_stats stats{};
    const auto t_render = std::chrono::steady_clock::now();
    // One render call takes at most 2^32 - 1 samples per GPU (and keeps a path slot for each in flight): a frame
    // with more -- 720x1280 from 4661 spp on one GPU -- is rendered as several sample ranges, whose film sums add up.
    const uint64_t n_pixels = (uint64_t)width * height;
    uint64_t per_call = std::max<uint64_t>(1, 0xFFFFFFFFull / std::max<uint64_t>(n_pixels, 1)) * (uint64_t)args.gpus;
    if (args.samples_per_call > 0) per_call = (uint64_t)args.samples_per_call;
    std::vector<float> part;
    for (uint64_t begin = 0;; begin += per_call) {
        const uint32_t end = (uint32_t)std::min<uint64_t>(spp, begin + per_call);
        const bool whole = begin == 0 && end == spp;
        if (!whole && part.empty()) part.resize(pixels.size());
        float* dst = whole ? pixels.data() : part.data();
        cray_render_stats st{};
        if (args.gpus == 1) rc = cray_render(scenes[0], mode, args.seed, (uint32_t)begin, end, dst, &st);
        else rc = cray_render_multi(scenes.data(), args.gpus, mode, args.seed, (uint32_t)begin, end, dst, &st);
        if (rc != CRAY_OK) break;
        if (!whole)
            for (size_t i = 0; i < pixels.size(); ++i) pixels[i] += part[i];
        stats.samples += st.samples; stats.closest_rays += st.closest_rays; stats.shadow_rays += st.shadow_rays;
        stats.shadow_rays_traced += st.shadow_rays_traced; stats.contact_rays += st.contact_rays; stats.nan_samples += st.nan_samples;
        stats.iterations += st.iterations; stats.kernel_launches += st.kernel_launches;
        if (end >= spp) break;
    }
    int status = 0;
    if (rc != CRAY_OK) {
        std::fprintf(stderr, "[ERROR] %s\n", cray_last_error());
        status = 1;
    } else {
        if (spp)
            for (float& p : pixels) p /= (float)spp;  // pixels /= num_samples (craytracer.rs:253-259)
        const double dt = seconds_since(t_render);
        // "reference rays": the Scene::intersect + Scene::intersects calls the reference makes for these samples; "traced": the rays this
        // implementation traced (a light sample that cannot contribute needs no shadow ray)
        const double rays = (double)stats.closest_rays + (double)stats.shadow_rays, traced = (double)stats.closest_rays + (double)stats.shadow_rays_traced;
        std::fprintf(stderr, "[INFO] Rendering finished in %.3fs (%.1f M reference rays/s, %.1f M traced rays/s, %.1f Msamples/s", seconds_since(start), rays / dt / 1e6,
                     traced / dt / 1e6, (double)width * height * spp / dt / 1e6);
        if (stats.nan_samples) std::fprintf(stderr, ", %llu samples dropped where the reference would assert", (unsigned long long)stats.nan_samples);
        std::fputs(")\n", stderr);
        rc = cray_write_exr(args.output.c_str(), width, height, pixels.data());
        if (rc != CRAY_OK) { std::fprintf(stderr, "[ERROR] %s\n", cray_last_error()); status = 1; }
        else std::fprintf(stderr, "[INFO] Output written to %s\n", args.output.c_str());
    }
    for (cray_scene* sc : scenes) cray_scene_destroy(sc);
    cray_host_scene_destroy(hs);
    return status;
}
