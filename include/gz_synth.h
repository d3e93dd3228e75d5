/* gz_synth.h — deterministic synthetic policy/value "network" used for parity runs.
 *
 * The reference's trained weights are absent (data/.MISSING_LARGE_BLOBS), so PUCT
 * parity between the reference C++ search (oracle/_ref) and the CUDA search is established with a
 * network that both sides can evaluate bit-identically: an order-independent 64-bit hash of the
 * encoded input planes, expanded into per-action policy values and value-head outputs with exact
 * or round-to-nearest fp32 arithmetic. Host (oracle/ref_harness.cpp) and device (csrc/gz_synth.cu)
 * include this same file.
 */
#ifndef GZ_SYNTH_H
#define GZ_SYNTH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GZS_HD __host__ __device__ __forceinline__
#else
#define GZS_HD static inline
#endif

GZS_HD uint64_t gz_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}

/* contribution of plane element e holding a non-zero float with bit pattern `bits` */
GZS_HD uint64_t gz_synth_elem(uint32_t e, uint32_t bits, uint64_t seed) {
    return gz_mix64(seed ^ ((((uint64_t)e + 1) << 32) | bits));
}

GZS_HD float gz_u24(uint64_t z) { return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f); }

/* raw policy output for (role r, action a): u^4 with u uniform in [0,1) */
GZS_HD float gz_synth_policy(uint64_t h, uint32_t r, uint32_t a) {
    const float u = gz_u24(gz_mix64(h ^ ((uint64_t)(r + 1) * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)(a + 1) * 0xD6E8FEB86659FD93ULL)));
#if defined(__CUDA_ARCH__)
    const float u2 = __fmul_rn(u, u);
    return __fmul_rn(u2, u2);
#else
    const float u2 = u * u;
    return u2 * u2;
#endif
}

/* value head: 2 outputs (v, 1-v) or 3 outputs (win0, win1, draw) summing to ~1 */
GZS_HD void gz_synth_values(uint64_t h, int n_rewards, float* out) {
    const float u = gz_u24(gz_mix64(h ^ 0x5151515151515151ULL));
    if (n_rewards == 3) {
        const uint32_t j = (uint32_t)(gz_mix64(h ^ 0x7272727272727272ULL) >> 48);   /* 16 bits */
        const float d = (float)j * (0.25f / 65536.0f);
        const float rem = 1.0f - d;
#if defined(__CUDA_ARCH__)
        const float a = __fmul_rn(rem, u);
        out[0] = a; out[1] = __fsub_rn(rem, a); out[2] = d;
#else
        const float a = rem * u;
        out[0] = a; out[1] = rem - a; out[2] = d;
#endif
    } else {
        out[0] = u;
        out[1] = 1.0f - u;
        for (int i = 2; i < n_rewards; i++) out[i] = 0.0f;
    }
}

#if !defined(__CUDA_ARCH__)
static inline uint64_t gz_synth_hash_planes_f32(const float* p, int total, uint64_t seed) {
    uint64_t h = 0;
    for (int e = 0; e < total; e++) {
        if (p[e] != 0.0f) {
            uint32_t bits;
            memcpy(&bits, &p[e], 4);
            h += gz_synth_elem((uint32_t)e, bits, seed);
        }
    }
    return h;
}
#endif

#endif
