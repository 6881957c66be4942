// ORACLE — TEST INFRASTRUCTURE ONLY (see geometry.hpp).
// Restatement of src/sampling.rs: sampling_fns (:1-66) and SobolSampler (:197-247).
//
// PARITY UNPINNED for the sampler values: the arithmetic lives in two dependencies
// that are not in the reference tree and that no reference test pins:
//   * sobol_burley 0.5.0 (Cargo.lock:1019) -- `sample(index, dim, seed) -> f32`,
//     Burley 2020 "Practical Hash-based Owen Scrambling": index shuffled by a
//     Laine-Karras style Owen scramble of its reversed bits, top 16 index bits used,
//     Joe-Kuo direction vectors, per-dimension Owen scramble, 24-bit float in [0,1).
//     Restated below from the published algorithm; the crate's exact hash constants
//     cannot be checked offline.  The direction vectors are rebuilt from the Joe-Kuo
//     data by tools/gen_sobol_table.py and verified against scipy.
//   * std::collections::hash_map::DefaultHasher == SipHash-1-3 with k0 = k1 = 0 over
//     the 24 bytes (seed, x, y as little-endian u64).  SipHash is restated from its
//     specification and pinned by (a) the SipHash-2-4 reference vector of the paper
//     and (b) CPython's siphash13 with PYTHONHASHSEED=0 (zero key) -- see tests.
// Product and oracle implement the same integer arithmetic, so GPU and oracle draw
// bit-identical sample values; parity with the reference's images is statistical.
#pragma once
#include "geometry.hpp"
#include "../include/cray_sobol_directions.h"  // generated constant table (tools/gen_sobol_table.py), shared with the product: data, not code

namespace orc {

// ---- sampling_fns  sampling.rs:1-66 -------------------------------------------------
inline double power_heuristic(uint64_t n_f, double pdf_f, uint64_t n_g, double pdf_g) {  // :11-15
    double f = (double)n_f * pdf_f;
    double g = (double)n_g * pdf_g;
    return (f * f) / (f * f + g * g);
}
inline void sample_disk(double u, double v, double& x, double& y) {  // :17-29
    if (u == 0.0 || v == 0.0) { x = 0.0; y = 0.0; return; }
    u = 2.0 * u - 1.0;
    v = 2.0 * v - 1.0;
    double r, theta;
    if (std::fabs(u) > std::fabs(v)) { r = u; theta = FRAC_PI_4 * v / u; }
    else { r = v; theta = FRAC_PI_2 - FRAC_PI_4 * u / v; }
    x = std::cos(theta) * r;
    y = std::sin(theta) * r;
}
inline V3 sample_sphere(double u, double v) {  // :31-39
    double z = 1.0 - 2.0 * u;
    double r = std::sqrt(rmax(1.0 - z * z, 0.0));  // z.powf(2.0) -> z*z
    double phi = 2.0 * PI * v;
    double x = r * std::cos(phi);
    double y = r * std::sin(phi);
    return {x, y, z};
}
inline V3 sample_hemisphere(double u, double v, V3 normal) {  // :41-48
    V3 r = sample_sphere(u, v);
    if (dot(r, normal) > 0.0) return r;
    return neg(r);
}
inline void sample_triangle(double u, double v, double& b1, double& b2) {  // :51-55
    double su = std::sqrt(u);
    b1 = 1.0 - su;
    b2 = v * su;
}
inline V3 cosine_sample_hemisphere(double u, double v, V3 normal, bool* assert_failed = nullptr) {  // :57-65
    V3 tangent, bitangent;
    generate_tangents(normal, tangent, bitangent);
    double x, y;
    sample_disk(u, v, x, y);
    double z = std::sqrt(rmax(1.0 - x * x - y * y, 0.0));
    V3 a = tangent * x + bitangent * y + normal * z;
    if (assert_failed && !(dot(a, normal) >= 0.0)) *assert_failed = true;  // assert! at :63
    return normalized(a);
}

// ---- SipHash-c-d (Aumasson & Bernstein).  DefaultHasher = SipHash-1-3, key (0,0). ----
inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
inline uint64_t siphash(int c_rounds, int d_rounds, uint64_t k0, uint64_t k1, const uint8_t* in, size_t len) {
    uint64_t v0 = 0x736f6d6570736575ULL ^ k0, v1 = 0x646f72616e646f6dULL ^ k1;
    uint64_t v2 = 0x6c7967656e657261ULL ^ k0, v3 = 0x7465646279746573ULL ^ k1;
    auto round = [&]() {
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
    };
    size_t full = len / 8;
    for (size_t i = 0; i < full; ++i) {
        uint64_t m = 0;
        for (int b = 0; b < 8; ++b) m |= (uint64_t)in[i * 8 + b] << (8 * b);
        v3 ^= m;
        for (int r = 0; r < c_rounds; ++r) round();
        v0 ^= m;
    }
    uint64_t last = (uint64_t)(len & 0xff) << 56;
    for (size_t b = 0; b < (len & 7); ++b) last |= (uint64_t)in[full * 8 + b] << (8 * b);
    v3 ^= last;
    for (int r = 0; r < c_rounds; ++r) round();
    v0 ^= last;
    v2 ^= 0xff;
    for (int r = 0; r < d_rounds; ++r) round();
    return v0 ^ v1 ^ v2 ^ v3;
}
// SobolSampler::start_pixel  sampling.rs:223-228: hash(seed), hash(x), hash(y) as usize -> 8 LE bytes each.
inline uint32_t pixel_hash(uint64_t seed, uint64_t x, uint64_t y) {
    uint8_t buf[24];
    uint64_t w[3] = {seed, x, y};
    for (int i = 0; i < 3; ++i)
        for (int b = 0; b < 8; ++b) buf[i * 8 + b] = (uint8_t)(w[i] >> (8 * b));
    return (uint32_t)siphash(1, 3, 0, 0, buf, 24);
}

// ---- sobol_burley::sample restated (see header comment) --------------------------------
inline uint32_t reverse_bits32(uint32_t x) {
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    x = ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
    return x;
}
inline uint32_t sobol_hash(uint32_t n) {
    n ^= 0x79c68e4au;
    for (int i = 0; i < 3; ++i) {
        n *= 0x736caf6fu;
        n ^= n >> 16;
    }
    return n;
}
// Owen scramble of a bit-reversed integer (Laine-Karras permutation, Vegdahl's constants).
inline uint32_t owen_scramble_rev(uint32_t n_rev, uint32_t scramble) {
    scramble = sobol_hash(scramble);
    n_rev ^= n_rev * 0x3d20adeau;
    n_rev += scramble;
    n_rev *= (scramble >> 16) | 1u;
    n_rev ^= n_rev * 0x05526c56u;
    n_rev ^= n_rev * 0x53a22864u;
    return n_rev;
}
inline float sobol_burley_sample(uint32_t sample_index, uint32_t dimension, uint32_t seed) {
    // Shuffle the index: Owen scramble of the reversed index bits, keyed by the seed.
    uint32_t shuffled_rev = owen_scramble_rev(reverse_bits32(sample_index), seed);
    // Sobol point from the low 16 index bits (= top 16 bits of the reversed index).
    uint32_t sobol = 0;
    const uint32_t* vecs = SOBOL_DIRECTIONS_INIT[dimension & (SOBOL_NUM_DIMENSIONS - 1)];
    for (int k = 0; k < 16; ++k)
        if (shuffled_rev & (0x80000000u >> k)) sobol ^= vecs[k];
    // Per-dimension Owen scramble of the reversed Sobol bits.
    uint32_t dim_seed = seed ^ (dimension * 0x9e3779b9u + 0x7f4a7c15u);
    uint32_t scrambled = reverse_bits32(owen_scramble_rev(reverse_bits32(sobol), dim_seed));
    return (float)(scrambled >> 8) * (1.0f / 16777216.0f);
}

struct SobolSampler {  // sampling.rs:197-247
    uint64_t seed = 0;
    uint32_t hash = 0, sample_index = 0, dimension = 0;
    void start_pixel(uint64_t x, uint64_t y, uint64_t si) {
        hash = pixel_hash(seed, x, y);
        sample_index = (uint32_t)si;
        dimension = 0;
    }
    double sample_1d() {
        float s = sobol_burley_sample(sample_index, dimension, hash);
        dimension += 1;
        return (double)s;
    }
    void sample_2d(double& a, double& b) {
        a = sample_1d();
        b = sample_1d();
    }
};

}  // namespace orc
