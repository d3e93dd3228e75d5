/* TEST INFRASTRUCTURE — pins galvanise_zero_b200/csrc/gz_libm.h against the libraries it restates:
 * the host's glibc logf / powf and libstdc++'s std::gamma_distribution<float> driven by the same xoroshiro
 * generator the reference uses (oracle/shims/k273/rng.h). Every entry returns the number of mismatching
 * results (bit comparison). Built by tests/hostsim/Makefile, called from tests/test_libm_restatement.py. */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "../../galvanise_zero_b200/csrc/gz_libm.h"
#include "../../oracle/shims/k273/rng.h"

uint64_t K273::gzref_next_seed() { return 0x5EED0000ULL; }

static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline bool same(float a, float b) { return bits(a) == bits(b) || (a != a && b != b); }

static inline uint64_t mix(uint64_t& s) {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* logf over `n` random bit patterns (all exponents, both signs) + every float in [lo_bits, hi_bits) */
extern "C" long libmcheck_logf(long n, uint64_t seed, uint32_t lo_bits, uint32_t hi_bits, uint32_t* first_bad) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        const uint32_t u = (uint32_t)mix(seed);
        float x; memcpy(&x, &u, 4);
        if (!same(gz_logf(x), std::log(x))) { if (!bad && first_bad) *first_bad = u; bad++; }
    }
    for (uint32_t u = lo_bits; u < hi_bits; u++) {
        float x; memcpy(&x, &u, 4);
        if (!same(gz_logf(x), std::log(x))) { if (!bad && first_bad) *first_bad = u; bad++; }
    }
    return bad;
}

/* powf(u, 1/alpha) on the sampler's domain: u = canonical draw in (0,1), alpha = 10.83f / nc, nc in [2, max_nc] */
extern "C" long libmcheck_powf(long n, uint64_t seed, int max_nc, uint32_t* first_bad) {
    long bad = 0;
    for (long i = 0; i < n; i++) {
        const uint32_t d = (uint32_t)mix(seed);
        const float u = gz_canonical_float(d ? d : 1u);
        const int nc = 2 + (int)(mix(seed) % (uint64_t)(max_nc - 1));
        const float alpha = 10.83f / nc;
        const float y = 1.0f / alpha;
        if (!same(gz_powf(u, y), std::pow(u, y))) { if (!bad && first_bad) { first_bad[0] = bits(u); first_bad[1] = bits(y); } bad++; }
    }
    /* general positive x, moderate y */
    for (long i = 0; i < n; i++) {
        uint32_t xb = (uint32_t)mix(seed) & 0x7fffffffu;
        if (xb >= 0x7f800000u) xb = 0x3f000000u;
        float x; memcpy(&x, &xb, 4);
        if (x == 0.0f) x = 1.0f;
        const float y = (float)((double)(int64_t)(mix(seed) % 2000001ULL) / 1000.0 - 1000.0);
        if (!same(gz_powf(x, y), std::pow(x, y))) { if (!bad && first_bad) { first_bad[0] = bits(x); first_bad[1] = bits(y); } bad++; }
    }
    return bad;
}

struct Xo { K273::xoroshiro128plus32* r; uint32_t operator()() { return r->next32(); } };

/* for nc in [2, max_nc]: alpha = 10.83f / nc, `nc` draws from one distribution object (the reference's
 * setDirichletNoise loop, evaluator.cpp:1249-1258), `rounds` seeds each */
extern "C" long libmcheck_gamma(int max_nc, int rounds, uint64_t seed, long* n_draws) {
    long bad = 0, draws = 0;
    for (int nc = 2; nc <= max_nc; nc++) {
        for (int k = 0; k < rounds; k++) {
            const uint64_t s = mix(seed);
            K273::xoroshiro128plus32 ra(s), rb(s);
            const float alpha = 10.83f / nc;
            std::gamma_distribution<float> gamma(alpha, 1.0f);
            GzGammaF g;
            gz_gamma_init(g, alpha);
            Xo xb{&rb};
            for (int i = 0; i < nc; i++) {
                const float a = gamma(ra);
                const float b = gz_gamma_float(g, xb);
                draws++;
                if (!same(a, b)) bad++;
            }
            if (ra.next32() != rb.next32()) bad++;      /* same number of draws consumed */
        }
    }
    if (n_draws) *n_draws = draws;
    return bad;
}

extern "C" long libmcheck_canonical(long n, uint64_t seed) {
    long bad = 0;
    struct One { uint32_t v; typedef uint32_t result_type; static constexpr uint32_t min() { return 0; }
                 static constexpr uint32_t max() { return 0xFFFFFFFFu; } uint32_t operator()() { return v; } };
    for (long i = 0; i < n + 4; i++) {
        One o{(uint32_t)mix(seed)};
        const uint32_t edge[4] = {0u, 0xFFFFFFFFu, 0xFFFFFF80u, 0xFFFFFF7Fu};
        if (i < 4) o.v = edge[i];
        const float a = std::generate_canonical<float, 24>(o);
        if (!same(a, gz_canonical_float(o.v))) bad++;
    }
    return bad;
}
