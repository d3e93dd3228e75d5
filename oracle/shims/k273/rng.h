/* ORACLE shim for k273/rng.h: xoroshiro128+ with 32-bit output (upper half), usable as a standard
 * URBG (src/cpp/puct/evaluator.cpp:1256). K273's own seeding and get()/getWithMax() mapping are not
 * in the reference; here: get() = next32 / 2^32 in [0,1), getWithMax(n) = next32 % n, and
 * default-constructed generators are seeded from a process-wide deterministic counter that
 * gzref_seed() resets (so runs are reproducible). */
#pragma once
#include <cstdint>

namespace K273 {

uint64_t gzref_next_seed();

class xoroshiro128plus32 {
public:
    typedef uint32_t result_type;
    xoroshiro128plus32() { seed(gzref_next_seed()); }
    explicit xoroshiro128plus32(uint64_t s) { seed(s); }
    void seed(uint64_t s) {
        // splitmix64 expansion
        for (int i = 0; i < 2; i++) {
            s += 0x9E3779B97F4A7C15ULL;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            st[i] = z ^ (z >> 31);
        }
    }
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    result_type operator()() { return next32(); }
    uint32_t next32() {
        const uint64_t s0 = st[0];
        uint64_t s1 = st[1];
        const uint64_t r = s0 + s1;
        s1 ^= s0;
        st[0] = ((s0 << 55) | (s0 >> 9)) ^ s1 ^ (s1 << 14);
        st[1] = (s1 << 36) | (s1 >> 28);
        return (uint32_t)(r >> 32);
    }
    double get() { return next32() * (1.0 / 4294967296.0); }
    uint32_t getWithMax(uint32_t n) { return n ? next32() % n : 0; }
private:
    uint64_t st[2];
};

}  // namespace K273
