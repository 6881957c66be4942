// Host-side owner of one GPU-resident scene (the opaque `cray_scene` of include/cray_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "bvh_build.hpp"
#include "device_types.cuh"

struct cray_scene {
    int device = 0;
    uint32_t build_flags = 0;
    cray::SceneView view{};
    std::vector<void*> allocations;    // every cudaMalloc of this scene
    uint32_t* d_sobol = nullptr;       // Sobol byte tables [256][2][256] (sampler.cuh)
    uint32_t* d_pixel_order = nullptr; // tile-ordered pixel list (x | y << 16)
    cray_scene_info info{};
    cudaStream_t stream = nullptr;     // calls on one handle are serialised on this stream
    // lazily sized wavefront pool (see wavefront.cu)
    void* pool = nullptr;
};

namespace cray {
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define CRAY_CUDA(expr)                                            \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::cray::cuda_fail(_e, #expr); \
    } while (0)
}  // namespace cray
