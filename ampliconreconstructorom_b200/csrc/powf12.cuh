// Bit-exact reproduction of glibc 2.39 powf(x, 1.2f) for x >= 0 (the `delta = powf(discrep, 1.2f)` term of
// SegAligner's score_f, /root/reference/SegAligner/SegAligner.cpp:88).
//
// The reference links the host libm dynamically; on every FMA-capable x86-64 the ifunc resolver selects
// __powf_fma (sysdeps/x86_64/fpu/multiarch/e_powf.c built from sysdeps/ieee754/flt-32/e_powf.c with
// -mfma -mavx2).  That routine is the public ARM "optimized-routines" powf: log2 via a 16-entry
// (invc, logc) table + degree-5 polynomial, y*log2(x), exp2 via a 32-entry table + cubic -- all in fp64,
// 18 fp64 operations, one final round to fp32.  It is NOT correctly rounded, so a generic pow cannot
// substitute: the exact operation sequence (including which multiply-adds are fused) must be replayed.
// The sequence below was read off the disassembly of __powf_fma in this image's libm.so.6 and is
// checked exhaustively against the host powf: the host-compiled replay in tests/test_powf12_host.py (csrc/tools/powf_check_host.cpp),
// both device paths in tests/test_gpu_parity.py::test_powf12_device_exhaustive (every non-negative float).
//
// Compiles as host code (SA_HD empty, uses fma()) and as device code (DFMA is IEEE-754 fused).
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define SA_HD __host__ __device__ __forceinline__
#else
#define SA_HD static inline
#endif

namespace sa {

// ---- published table data (glibc sysdeps/ieee754/flt-32/e_powf_log2_data.c, e_exp2f_data.c) ----------
#define SA_POWF_OFF 0x3f330000u
// T[i] = {invc, logc}: c near the centre of the i-th subinterval of [OFF, 2*OFF), invc = 1/c, logc = log2(c)
#define SA_POWF_LOG2_TAB                                                                                         \
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},               \
        {0x1.49539f0f010bp+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},             \
        {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8eap+0, -0x1.97c1d1b3b7afp-3},              \
        {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},            \
        {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1p+0, 0x0p+0},                                         \
        {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4}, {0x1.ca4b31f026aap-1, 0x1.476a9543891bap-3},               \
        {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},              \
        {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2}, {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}
// log2 polynomial A[0..4]
#define SA_POWF_A0 0x1.27616c9496e0bp-2
#define SA_POWF_A1 -0x1.71969a075c67ap-2
#define SA_POWF_A2 0x1.ec70a6ca7baddp-2
#define SA_POWF_A3 -0x1.7154748bef6c8p-1
#define SA_POWF_A4 0x1.71547652ab82bp0
// exp2: tab[i] = asuint64(2^(i/32)) - (i << 47); SHIFT = 0x1.8p52/32; C = poly
#define SA_EXP2F_TAB                                                                                             \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,                 \
        0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,             \
        0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,             \
        0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,             \
        0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,             \
        0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,             \
        0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,             \
        0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull
#define SA_EXP2F_SHIFT 0x1.8p+47
#define SA_EXP2F_C0 0x1.c6af84b912394p-5
#define SA_EXP2F_C1 0x1.ebfce50fac4f3p-3
#define SA_EXP2F_C2 0x1.62e42ff0c52d6p-1
// (double)1.2f
#define SA_Y12 0x1.3333340000000p+0

struct PowLogEntry {
    double invc, logc;
};

SA_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
SA_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
SA_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
SA_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
// explicit single-rounding fp64 primitives (no contraction / re-association by either compiler)
SA_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
SA_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
SA_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// log2 part given the 16-entry table entry and k; returns y*log2(x) (10 fp64 ops incl. the y multiply)
SA_HD double powf12_ylogx(uint32_t iz, double invc, double y0) {
    double z = (double)u2f(iz);
    double r = dfma(z, invc, -1.0);       // vfmadd132sd (%rcx),-1.0
    double y = dfma(r, SA_POWF_A0, SA_POWF_A1);
    double p = dfma(r, SA_POWF_A2, SA_POWF_A3);
    double r2 = dmul(r, r);
    double q = dfma(r, SA_POWF_A4, y0);
    double r4 = dmul(r2, r2);
    q = dfma(r2, p, q);
    y = dfma(y, r4, q);
    return dmul(SA_Y12, y);
}

// exp2 part: 8 fp64 ops + table value t = tab[ki % 32] supplied through the functor-like macro below
SA_HD float powf12_exp2(double xd, const uint64_t *tab) {
    double kd = dadd(xd, SA_EXP2F_SHIFT);
    uint64_t ki = d2u(kd);
    kd = dadd(kd, -SA_EXP2F_SHIFT);
    double r = dadd(xd, -kd);
    uint64_t t = tab[ki & 31];
    t += ki << 47;
    double s = u2d(t);
    double z = dfma(r, SA_EXP2F_C0, SA_EXP2F_C1);
    double r2 = dmul(r, r);
    double y = dfma(r, SA_EXP2F_C2, 1.0);
    y = dfma(z, r2, y);
    y = dmul(y, s);
    return (float)y;
}

// Generic path: any x >= 0 incl. 0, subnormals, inf, NaN (negative x never occurs: the argument is a fabs()).
// logtab: 16 entries {invc, logc}; exptab: 32 entries.
SA_HD float powf12_generic(float x, const PowLogEntry *logtab, const uint64_t *exptab) {
    uint32_t ix = f2u(x) & 0x7fffffffu;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix == 0) return 0.0f;             // pow(+0, y>0) = +0
        if (ix == 0x7f800000u) return u2f(ix);  // pow(+inf, y>0) = +inf
        if (ix > 0x7f800000u) return x + 1.2f;  // NaN
        // subnormal: normalise (e_powf.c: ix = asuint(x * 0x1p23f) & 0x7fffffff; ix -= 23 << 23)
        ix = f2u(u2f(ix) * 0x1p23f);
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - SA_POWF_OFF;
    int i = (tmp >> 19) & 15;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double y0 = dadd(logtab[i].logc, (double)k);
    double ylogx = powf12_ylogx(iz, logtab[i].invc, y0);
    // |y*log2(x)| >= 126: overflow / underflow handling of e_powf.c
    if (((d2u(ylogx) >> 47) & 0xffff) >= (d2u(126.0) >> 47)) {
        if (ylogx > 0x1.fffffffd1d571p+6) return u2f(0x7f800000u);  // __math_oflowf(0)
        if (ylogx <= -150.0) return 0.0f;                            // __math_uflowf(0)
        if (ylogx < -149.0) return 0x1p-149f;                        // __math_may_uflowf(0): 0x1.4p-75f squared
    }
    return powf12_exp2(ylogx, exptab);
}

}  // namespace sa
