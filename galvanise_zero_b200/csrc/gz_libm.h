// gz_libm.h — bit-level restatements of the library routines the reference's Dirichlet noise goes through.
//
// The reference draws its root noise with libstdc++'s std::gamma_distribution<float> over its xoroshiro
// generator (src/cpp/puct/evaluator.cpp:1249-1258). For the device search to pick the same children as the
// reference, the same floats have to come out, so the whole chain is restated here as host/device code:
//
//   libstdc++ 13  bits/random.tcc   generate_canonical<float, 24>   (one 32-bit draw / 2^32, clamped below 1)
//                                   normal_distribution<float>      (Marsaglia polar, keeps its second value)
//                                   gamma_distribution<float>       (Marsaglia-Tsang, alpha < 1 boosted by pow)
//   glibc 2.39    libm              logf, powf                      (ARM optimized-routines algorithms:
//                                   sysdeps/ieee754/flt-32/e_logf.c, e_powf.c; double arithmetic, table driven)
//
// Neither library is part of /root/reference (they are the toolchain the reference is built with); the tables
// are read out of the build container's libm.so.6 by tools/gen_libm_tables.py. tests/test_libm_restatement.py
// pins the restatement: gz_logf / gz_powf against the host's logf / powf bit for bit, and the samplers against
// std::gamma_distribution<float> driven by the same generator. CUDA's own logf / powf are NOT used on this
// path (they differ from glibc's in the last bit often enough to change a noise vector).
//
// Compile without floating-point contraction (nvcc -fmad=false, g++ -ffp-contract=off).
#ifndef GZ_LIBM_H
#define GZ_LIBM_H

#include <stdint.h>
#include <string.h>

#include "gz_libm_tables.h"

#ifdef __CUDACC__
#define GZM_HD __host__ __device__ __forceinline__
#define GZM_HDN __host__ __device__
#else
#define GZM_HD static inline
#define GZM_HDN static inline
#endif

#ifdef __CUDACC__
static __device__ const double gzm_logf_tab_d[32] = GZ_LOGF_TAB;
static __device__ const double gzm_logf_poly_d[3] = GZ_LOGF_POLY;
static __device__ const double gzm_powf_log2_tab_d[32] = GZ_POWF_LOG2_TAB;
static __device__ const double gzm_powf_log2_poly_d[5] = GZ_POWF_LOG2_POLY;
static __device__ const unsigned long long gzm_exp2f_tab_d[32] = GZ_EXP2F_TAB;
static __device__ const double gzm_exp2f_poly_d[3] = GZ_EXP2F_POLY;
#endif
static const double gzm_logf_tab_h[32] = GZ_LOGF_TAB;
static const double gzm_logf_poly_h[3] = GZ_LOGF_POLY;
static const double gzm_powf_log2_tab_h[32] = GZ_POWF_LOG2_TAB;
static const double gzm_powf_log2_poly_h[5] = GZ_POWF_LOG2_POLY;
static const unsigned long long gzm_exp2f_tab_h[32] = GZ_EXP2F_TAB;
static const double gzm_exp2f_poly_h[3] = GZ_EXP2F_POLY;

#ifdef __CUDA_ARCH__
#define GZM_TAB(name) name##_d
#else
#define GZM_TAB(name) name##_h
#endif

GZM_HD uint32_t gzm_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
GZM_HD float gzm_asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
GZM_HD uint64_t gzm_asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
GZM_HD double gzm_asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

// glibc logf (e_logf.c): x = 2^k z, z in [OFF, 2 OFF) split into 16 intervals with centre c;
// log(x) = log1p(z/c - 1) + log(c) + k ln2, degree-3 polynomial, everything in double, one rounding at the end.
GZM_HDN float gz_logf(float x) {
    const double* T = GZM_TAB(gzm_logf_tab);
    const double* A = GZM_TAB(gzm_logf_poly);
    uint32_t ix = gzm_asuint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2u == 0u) return gzm_asfloat(0xff800000u);                    // log(0) = -inf
        if (ix == 0x7f800000u) return x;                                    // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return gzm_asfloat(0x7fc00000u);
        ix = gzm_asuint(x * 0x1p23f);                                       // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t OFF = 0x3f330000u;
    const uint32_t tmp = ix - OFF;
    const int i = (int)((tmp >> (23 - 4)) % 16u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    const double invc = T[2 * i], logc = T[2 * i + 1];
    const double z = (double)gzm_asfloat(iz);
    const double r = z * invc - 1;
    const double y0 = logc + (double)k * GZ_LOGF_LN2;
    const double r2 = r * r;
    double y = A[1] * r + A[2];
    y = A[0] * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

// glibc powf (e_powf.c) for x > 0 finite and y finite: 2^(y log2 x) with log2 from a 16-entry table and a
// degree-5 polynomial, exp2 from a 32-entry table and a degree-3 polynomial, in double. The branches for
// negative / infinite / NaN arguments are not needed by the callers and not restated.
GZM_HDN float gz_powf(float x, float y) {
    const double* T = GZM_TAB(gzm_powf_log2_tab);
    const double* A = GZM_TAB(gzm_powf_log2_poly);
    const unsigned long long* E = GZM_TAB(gzm_exp2f_tab);
    const double* C = GZM_TAB(gzm_exp2f_poly);
    uint32_t ix = gzm_asuint(x);
    const uint32_t iy = gzm_asuint(y);
    if (2u * iy == 0u) return 1.0f;                                         // pow(x, +-0) = 1
    if (ix == 0x3f800000u) return 1.0f;                                     // pow(1, y) = 1
    if (ix < 0x00800000u) {                                                 // subnormal x: normalise
        ix = gzm_asuint(x * 0x1p23f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    // log2_inline
    const uint32_t OFF = 0x3f330000u;
    const uint32_t tmp = ix - OFF;
    const int i = (int)((tmp >> (23 - 4)) % 16u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int32_t)top >> 23;
    const double invc = T[2 * i], logc = T[2 * i + 1];
    const double z = (double)gzm_asfloat(iz);
    const double r = z * invc - 1;
    const double y0 = logc + (double)k;
    const double r2 = r * r;
    double yy = A[0] * r + A[1];
    const double p = A[2] * r + A[3];
    const double r4 = r2 * r2;
    double q = A[4] * r + y0;
    q = p * r2 + q;
    yy = yy * r4 + q;
    const double logx = yy;
    const double ylogx = (double)y * logx;
    if (((gzm_asuint64(ylogx) >> 47) & 0xffffu) >= (gzm_asuint64(126.0) >> 47)) {
        if (ylogx > 0x1.fffffffd1d571p+6) return gzm_asfloat(0x7f800000u);  // overflow
        if (ylogx <= -150.0) return 0.0f;                                   // underflow
        if (ylogx < -149.0) return 0x1p-149f;                               // __math_may_uflowf: 0x1.4p-75f squared
    }
    // exp2_inline
    const double SHIFT = GZ_EXP2F_SHIFT_SCALED;
    double kd = ylogx + SHIFT;
    const uint64_t ki = gzm_asuint64(kd);
    kd -= SHIFT;
    const double rr = ylogx - kd;
    uint64_t t = E[ki % 32u];
    t += ki << (52 - 5);
    const double s = gzm_asdouble(t);
    const double zz = C[0] * rr + C[1];
    const double rr2 = rr * rr;
    double e = C[2] * rr + 1;
    e = zz * rr2 + e;
    e = e * s;
    return (float)e;
}

// libstdc++ generate_canonical<float, 24>(urng) for a generator with a 32-bit range: one draw, divided by 2^32
// in float, and the value that would round to 1 replaced by the largest float below 1.
GZM_HD float gz_canonical_float(uint32_t draw) {
    float ret = (float)draw / 4294967296.0f;
    if (ret >= 1.0f) ret = 0x1.fffffep-1f;      // nextafter(1.0f, 0.0f)
    return ret;
}

// libstdc++ normal_distribution<float> (mean 0, stddev 1): polar method, the partner value is kept for the
// next call. `Rng` provides uint32_t operator()().
struct GzNormalF {
    float saved;
    bool saved_available;
};
template <typename Rng>
GZM_HDN float gz_normal_float(GzNormalF& nd, Rng& rng) {
    float ret;
    if (nd.saved_available) {
        nd.saved_available = false;
        ret = nd.saved;
    } else {
        float x, y, r2;
        do {
            x = (float)((double)(2.0f * gz_canonical_float(rng())) - 1.0);
            y = (float)((double)(2.0f * gz_canonical_float(rng())) - 1.0);
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const float mult = sqrtf(-2 * gz_logf(r2) / r2);
        nd.saved = x * mult;
        nd.saved_available = true;
        ret = y * mult;
    }
    ret = ret * 1.0f + 0.0f;
    return ret;
}

// libstdc++ gamma_distribution<float>(alpha, 1): parameters as param_type::_M_initialize computes them, then
// operator(). One GzGammaF lives as long as the reference's distribution object (its normal keeps state).
struct GzGammaF {
    float alpha, malpha, a2;
    GzNormalF nd;
};
GZM_HD void gz_gamma_init(GzGammaF& g, float alpha) {
    g.alpha = alpha;
    g.malpha = alpha < 1.0 ? alpha + 1.0f : alpha;
    const float a1 = g.malpha - 1.0f / 3.0f;
    g.a2 = 1.0f / sqrtf(9.0f * a1);
    g.nd.saved = 0.0f;
    g.nd.saved_available = false;
}
template <typename Rng>
GZM_HDN float gz_gamma_float(GzGammaF& g, Rng& rng) {
    float u, v, n;
    const float a1 = g.malpha - 1.0f / 3.0f;
    do {
        do {
            n = gz_normal_float(g.nd, rng);
            v = 1.0f + g.a2 * n;
        } while (v <= 0.0);
        v = v * v * v;
        u = gz_canonical_float(rng());
    } while ((double)u > (double)1.0f - 0.0331 * n * n * n * n &&
             ((double)gz_logf(u) > (0.5 * n * n + (double)a1 * (1.0 - (double)v + (double)gz_logf(v)))));
    if (g.alpha == g.malpha) return a1 * v * 1.0f;
    do {
        u = gz_canonical_float(rng());
    } while (u == 0.0);
    return gz_powf(u, 1.0f / g.alpha) * a1 * v * 1.0f;
}

#endif  // GZ_LIBM_H
