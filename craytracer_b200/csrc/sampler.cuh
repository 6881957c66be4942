// Sampler and sampling warps (src/sampling.rs).
//
// The sample values come from two dependencies that are not part of the reference tree (sobol_burley 0.5.0 and
// Rust's DefaultHasher = SipHash-1-3 with a zero key); both are restated from their published algorithms, see
// DESIGN.md "sampler".  Everything here is integer arithmetic plus one u32 -> f32 conversion, so the GPU draws
// bit-identical values to the CPU oracle.
#pragma once
#include "cray_math.cuh"

#ifndef CRAY_SOBOL_BYTES
#define CRAY_SOBOL_BYTES 1
#endif

namespace cray {

// ---- SipHash-1-3, k0 = k1 = 0, over (seed, x, y) as three little-endian u64 words (sampling.rs:223-228) ----
CRAY_HD uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
CRAY_HD void sip_round(uint64_t& v0, uint64_t& v1, uint64_t& v2, uint64_t& v3) {
    v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
    v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
}
CRAY_HD uint32_t pixel_hash(uint64_t seed, uint64_t x, uint64_t y) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL, v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    const uint64_t words[3] = {seed, x, y};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        v3 ^= words[i];
        sip_round(v0, v1, v2, v3);
        v0 ^= words[i];
    }
    const uint64_t last = 24ULL << 56;  // message length in the top byte, no tail bytes
    v3 ^= last;
    sip_round(v0, v1, v2, v3);
    v0 ^= last;
    v2 ^= 0xff;
    sip_round(v0, v1, v2, v3);
    sip_round(v0, v1, v2, v3);
    sip_round(v0, v1, v2, v3);
    return (uint32_t)(v0 ^ v1 ^ v2 ^ v3);
}

// ---- hash-based Owen-scrambled Sobol (Burley 2020), the algorithm of sobol_burley::sample ----
CRAY_HD uint32_t reverse_bits32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
#endif
}
CRAY_HD uint32_t sobol_hash(uint32_t n) {
    n ^= 0x79c68e4au;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        n *= 0x736caf6fu;
        n ^= n >> 16;
    }
    return n;
}
CRAY_HD uint32_t owen_scramble_rev(uint32_t n_rev, uint32_t scramble) {
    scramble = sobol_hash(scramble);
    n_rev ^= n_rev * 0x3d20adeau;
    n_rev += scramble;
    n_rev *= (scramble >> 16) | 1u;
    n_rev ^= n_rev * 0x05526c56u;
    n_rev ^= n_rev * 0x53a22864u;
    return n_rev;
}
// `table` = the Joe-Kuo direction vectors folded into byte tables, [256 dimensions][2][256] (scene_device.cu): entry
// [d][h][b] is the XOR of the vectors k = 8h .. 8h+7 of dimension d whose index bit (0x80 >> (k - 8h)) is set in b.  Only the
// top 16 bits of the reversed index are used (2^16 indices), so a sample is two table reads instead of sixteen.
__device__ __forceinline__ float sobol_sample(const uint32_t* __restrict__ table, uint32_t shuffled_rev_index, uint32_t dimension, uint32_t seed) {
#if CRAY_SOBOL_BYTES
    const uint32_t* t = table + 512u * (dimension & 255u);
    const uint32_t sobol = __ldg(t + (shuffled_rev_index >> 24)) ^ __ldg(t + 256u + ((shuffled_rev_index >> 16) & 255u));
#else  // tuning build: the plain direction vectors, [256][32]
    const uint32_t* vecs = table + 32u * (dimension & 255u);
    uint32_t sobol = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (shuffled_rev_index & (0x80000000u >> k)) sobol ^= vecs[k];
#endif
    const uint32_t dim_seed = seed ^ (dimension * 0x9e3779b9u + 0x7f4a7c15u);
    const uint32_t scrambled = reverse_bits32(owen_scramble_rev(reverse_bits32(sobol), dim_seed));
    return (float)(scrambled >> 8) * (1.0f / 16777216.0f);
}

struct PixelSampler {  // SobolSampler, sampling.rs:197-247
    uint32_t hash;           // per-pixel scramble seed
    uint32_t shuffled_rev;   // Owen-shuffled, bit-reversed sample index (depends on hash and sample_index only)
    uint32_t dimension;
    __device__ __forceinline__ void start_pixel(uint64_t seed, uint32_t x, uint32_t y, uint32_t sample_index) {
        hash = pixel_hash(seed, x, y);
        shuffled_rev = owen_scramble_rev(reverse_bits32(sample_index), hash);
        dimension = 0;
    }
    __device__ __forceinline__ double sample_1d(const uint32_t* table) {
        const float v = sobol_sample(table, shuffled_rev, dimension, hash);
        dimension += 1;
        return (double)v;
    }
};

// The eight sampler values of one path vertex, PathSegmentSamples::from (path_integrator.rs:25-36).  The reference draws all
// eight up front; a value is a pure function of (sample index, dimension, pixel hash), so each is computed where -- and only
// if -- it is consumed.
struct VertexSamples {
    const uint32_t* table;
    uint32_t shuffled_rev, hash, first_dimension;
    enum : uint32_t { MATERIAL_1D = 0, MATERIAL_U = 1, MATERIAL_V = 2, LIGHT_INDEX = 3, LIGHT_1D = 4, LIGHT_U = 5, LIGHT_V = 6, ROULETTE = 7 };
    __device__ __forceinline__ double get(uint32_t k) const { return (double)sobol_sample(table, shuffled_rev, first_dimension + k, hash); }
};

// ---- sampling_fns sampling.rs:1-66 ----
CRAY_HD double power_heuristic(double pdf_f, double pdf_g) {  // n_f = n_g = 1 at every call site
    const double f = 1.0 * pdf_f, g = 1.0 * pdf_g;
    return (f * f) / (f * f + g * g);
}
CRAY_HD void sample_disk(double u, double v, double& x, double& y) {
    if (u == 0.0 || v == 0.0) { x = 0.0; y = 0.0; return; }
    u = 2.0 * u - 1.0;
    v = 2.0 * v - 1.0;
    double r, theta;
    if (fabs(u) > fabs(v)) { r = u; theta = kFracPi4 * v / u; }
    else { r = v; theta = kFracPi2 - kFracPi4 * u / v; }
    x = cos(theta) * r;
    y = sin(theta) * r;
}
CRAY_HD V3 sample_sphere(double u, double v) {
    const double z = 1.0 - 2.0 * u;
    const double r = sqrt(rmax(1.0 - z * z, 0.0));
    const double phi = 2.0 * kPi * v;
    return mk(r * cos(phi), r * sin(phi), z);
}
CRAY_HD V3 sample_hemisphere(double u, double v, V3 normal) {
    const V3 r = sample_sphere(u, v);
    return dot(r, normal) > 0.0 ? r : neg(r);
}
CRAY_HD V3 cosine_sample_hemisphere(double u, double v, V3 normal, bool& assert_failed) {
    V3 tangent, bitangent;
    generate_tangents(normal, tangent, bitangent);
    double x, y;
    sample_disk(u, v, x, y);
    const double z = sqrt(rmax(1.0 - x * x - y * y, 0.0));
    const V3 a = tangent * x + bitangent * y + normal * z;
    if (!(dot(a, normal) >= 0.0)) assert_failed = true;  // assert! at sampling.rs:63
    return normalized(a);
}

}  // namespace cray
