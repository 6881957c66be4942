// cray_render_multi: one frame on several GPUs of ONE process (the reference's command line is a single process).
//
// (pixel, sample) units are independent, so the scene is replicated, GPU k renders the k-th contiguous slice of
// the sample range of every pixel, and the f32 sum films are combined with ONE ncclReduce(sum) onto the first GPU over NVLink.
// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it, and single-GPU users never load it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "scene_device.hpp"

extern "C" float* cray_scene_film_f32(cray_scene* sc);  // wavefront.cu: the scene's device staging film (W*H*3 f32)

// The handful of NCCL declarations used below, spelled out so that the library builds where the NCCL headers are absent (NCCL is
// only ever dlopen'd); values as in nccl.h 2.x.
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
typedef enum { ncclFloat = 7 } ncclDataType_t;
}

namespace cray {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(api.handle, "ncclCommInitAll"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(dlsym(api.handle, "ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.handle, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.handle, "ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
        api.ok = api.CommInitAll && api.Reduce && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api;
}

// one communicator clique per device list, created on first use and kept for the life of the process; its mutex is held across a
// whole group of collectives, so two callers never issue on the same communicators at once
struct Clique {
    std::vector<ncclComm_t> comms;
    std::mutex busy;
};
std::mutex g_comm_mutex;
std::map<std::vector<int>, Clique> g_comms;

int nccl_fail(ncclResult_t r, const char* what) {
    set_error(std::string("NCCL error: ") + nccl().GetErrorString(r) + " in " + what);
    return CRAY_E_CUDA;
}

}  // namespace
}  // namespace cray

using namespace cray;

extern "C" int cray_render_multi(cray_scene* const* scenes, int n, int mode, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                                 float* rgb_sum, cray_render_stats* stats) {
    if (!scenes || n <= 0 || !rgb_sum || sample_end < sample_begin) { set_error("bad arguments"); return CRAY_E_INVALID; }
    for (int k = 0; k < n; ++k) {
        if (!scenes[k]) { set_error("null scene"); return CRAY_E_INVALID; }
        if (scenes[k]->info.width != scenes[0]->info.width || scenes[k]->info.height != scenes[0]->info.height) { set_error("scenes differ in film size"); return CRAY_E_INVALID; }
        for (int j = 0; j < k; ++j)
            if (scenes[j]->device == scenes[k]->device) { set_error("two scenes on one device"); return CRAY_E_INVALID; }
    }
    if (n == 1) return cray_render(scenes[0], mode, seed, sample_begin, sample_end, rgb_sum, stats);
    NcclApi& api = nccl();
    if (!api.ok) { set_error("libnccl.so.2 could not be loaded: multi-GPU rendering needs NCCL"); return CRAY_E_UNSUPPORTED; }

    const uint32_t total = sample_end - sample_begin;
    std::vector<int> rcs(n, CRAY_OK);
    std::vector<std::string> errors(n);
    std::vector<cray_render_stats> st(n);
    std::vector<float*> films(n, nullptr);
    std::vector<std::thread> pool;
    for (int k = 0; k < n; ++k)
        pool.emplace_back([&, k] {
            // same split as craytracer_b200/distributed.py:shard_samples
            const uint32_t lo = sample_begin + (uint32_t)((uint64_t)total * k / n), hi = sample_begin + (uint32_t)((uint64_t)total * (k + 1) / n);
            if (cudaSetDevice(scenes[k]->device) != cudaSuccess) { rcs[k] = CRAY_E_CUDA; errors[k] = "cudaSetDevice failed"; return; }
            films[k] = cray_scene_film_f32(scenes[k]);
            if (!films[k]) { rcs[k] = CRAY_E_CUDA; errors[k] = cray_last_error(); return; }
            rcs[k] = cray_render_device(scenes[k], mode, seed, lo, hi, films[k], scenes[k]->stream, &st[k]);
            if (rcs[k] != CRAY_OK) errors[k] = cray_last_error();  // thread-local: carry it to the caller's thread
        });
    for (auto& th : pool) th.join();
    for (int k = 0; k < n; ++k)
        if (rcs[k] != CRAY_OK) { set_error(errors[k]); return rcs[k]; }

    std::vector<int> devices(n);
    for (int k = 0; k < n; ++k) devices[k] = scenes[k]->device;
    Clique* clique = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_comm_mutex);
        auto it = g_comms.find(devices);
        if (it == g_comms.end()) {
            std::vector<ncclComm_t> fresh(n);
            ncclResult_t r = api.CommInitAll(fresh.data(), n, devices.data());
            if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitAll");
            it = g_comms.try_emplace(devices).first;  // (std::map nodes never move: the pointer stays valid)
            it->second.comms = fresh;
        }
        clique = &it->second;
    }
    std::lock_guard<std::mutex> busy(clique->busy);
    const std::vector<ncclComm_t>& comms = clique->comms;
    const size_t count = (size_t)scenes[0]->info.width * scenes[0]->info.height * 3;
    ncclResult_t r = api.GroupStart();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupStart");
    for (int k = 0; k < n; ++k) {
        CRAY_CUDA(cudaSetDevice(devices[k]));
        r = api.Reduce(films[k], films[k], count, ncclFloat, ncclSum, 0, comms[k], scenes[k]->stream);  // in place on the root
        if (r != ncclSuccess) { api.GroupEnd(); return nccl_fail(r, "ncclReduce"); }
    }
    r = api.GroupEnd();
    if (r != ncclSuccess) return nccl_fail(r, "ncclGroupEnd");
    for (int k = 0; k < n; ++k) {
        CRAY_CUDA(cudaSetDevice(devices[k]));
        CRAY_CUDA(cudaStreamSynchronize(scenes[k]->stream));
    }
    CRAY_CUDA(cudaSetDevice(devices[0]));
    CRAY_CUDA(cudaMemcpy(rgb_sum, films[0], count * sizeof(float), cudaMemcpyDeviceToHost));
    if (stats) {
        *stats = st[0];
        for (int k = 1; k < n; ++k) {
            stats->samples += st[k].samples; stats->closest_rays += st[k].closest_rays; stats->shadow_rays += st[k].shadow_rays;
            stats->shadow_rays_traced += st[k].shadow_rays_traced; stats->contact_rays += st[k].contact_rays;
            stats->nan_samples += st[k].nan_samples; stats->kernel_launches += st[k].kernel_launches;
            stats->iterations = std::max(stats->iterations, st[k].iterations);
            stats->render_ms = std::max(stats->render_ms, st[k].render_ms);  // GPUs run side by side: the slowest one is the frame
            stats->trace_ms = std::max(stats->trace_ms, st[k].trace_ms); stats->shadow_ms = std::max(stats->shadow_ms, st[k].shadow_ms);
            stats->shade_ms = std::max(stats->shade_ms, st[k].shade_ms); stats->generate_ms = std::max(stats->generate_ms, st[k].generate_ms);
        }
    }
    return CRAY_OK;
}
