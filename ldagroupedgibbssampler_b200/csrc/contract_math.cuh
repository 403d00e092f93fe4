// contract_math.cuh -- device-side "contract" arithmetic of libldagpu.
//
// Everything is composed of IEEE-754 correctly rounded primitives in a fixed order
// (explicit _rn intrinsics: nothing here may be re-associated or fused by the compiler),
// so the CPU restatement in oracle/ reproduces the results bit for bit.  No libdevice
// transcendental is used.
//
// Gamma sampler: Marsaglia-Tsang with the alpha<1 boost, as the reference states it in
//   src/main/java/cc/mallet/util/ParallelRandoms.java:60-70 (rgamma), :148-159 (prgamma)
// with its unseedable generators (ParallelRandoms.java:16-24,65,153-155) replaced by one
// Philox4x32-10 block per attempt.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ldagpu {

enum : uint32_t { STREAM_Z = 1, STREAM_THETA = 2, STREAM_PHI = 3, STREAM_POISSON = 4 };

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// The same generator with the ten round keys (k + r * Weyl constant) formed once on the host and passed as a
// kernel argument: they reach the xor as constant-bank operands, and each 32x32 product is ONE wide multiply --
// about half the instructions of the version above, bit-identical output.
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
};
__host__ __device__ inline PhiloxKeys philox_keys(uint32_t k0, uint32_t k1)
{
    PhiloxKeys r;
    for (int i = 0; i < 10; ++i) {
        r.k0[i] = k0 + 0x9E3779B9u * (uint32_t)i;
        r.k1[i] = k1 + 0xBB67AE85u * (uint32_t)i;
    }
    return r;
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys &rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k0[r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k1[r];
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// open-interval uniform from the top 23 bits: (n + 1/2) 2^-23, exact in fp32
__device__ __forceinline__ float uniform23(uint32_t w)
{
    return __fmul_rn(__fadd_rn(__uint2float_rn(w >> 9), 0.5f), 0x1p-23f);
}

// ---------------------------------------------------------------------------------------
// per-type primitives
// ---------------------------------------------------------------------------------------
template <typename T> struct CM;

template <> struct CM<float> {
    using real = float;
    using U = uint32_t;
    using S = int32_t;
    static constexpr int MANT = 23, BIAS = 127, LN_TERMS = 5, EXP_DEG = 7, TRIG_DEG = 5, DENORM_SHIFT = 23;
    static constexpr U SQRT_HALF = 0x3f3504f3u, MANT_MASK = 0x007fffffu, MIN_NORMAL = 0x00800000u;
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ float rint(float a) { return rintf(a); }
    static __device__ __forceinline__ U bits(float a) { return __float_as_uint(a); }
    static __device__ __forceinline__ float from(U a) { return __uint_as_float(a); }
    static __device__ __forceinline__ float from_int(S a) { return __int2float_rn(a); }
    static __device__ __forceinline__ S to_int(float a) { return __float2int_rn(a); }
    static __device__ __forceinline__ float denorm_scale() { return 0x1p23f; }
    static __device__ __forceinline__ float exp_cutoff() { return -104.0f; }
    static __device__ __forceinline__ float ln2_hi() { return 0x1.62ep-1f; }
    static __device__ __forceinline__ float ln2_lo() { return 0x1.0bfbe8p-15f; }
    static __device__ __forceinline__ float uni(uint32_t w) { return uniform23(w); }
    static __device__ __forceinline__ float ang_frac(uint32_t w, bool odd)
    {
        uint32_t b = ((w >> 6) & 0x7fffffu) ^ (odd ? 0x7fffffu : 0u);
        return __fmul_rn(__fadd_rn(__uint2float_rn(b), 0.5f), 0x1p-23f);
    }
};

template <> struct CM<double> {
    using real = double;
    using U = unsigned long long;
    using S = long long;
    static constexpr int MANT = 52, BIAS = 1023, LN_TERMS = 11, EXP_DEG = 14, TRIG_DEG = 9, DENORM_SHIFT = 54;
    static constexpr U SQRT_HALF = 0x3fe6a09e667f3bcdull, MANT_MASK = 0x000fffffffffffffull,
                       MIN_NORMAL = 0x0010000000000000ull;
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double rint(double a) { return ::rint(a); }
    static __device__ __forceinline__ U bits(double a) { return (U)__double_as_longlong(a); }
    static __device__ __forceinline__ double from(U a) { return __longlong_as_double((long long)a); }
    static __device__ __forceinline__ double from_int(S a) { return __ll2double_rn(a); }
    static __device__ __forceinline__ S to_int(double a) { return __double2ll_rn(a); }
    static __device__ __forceinline__ double denorm_scale() { return 0x1p54; }
    static __device__ __forceinline__ double exp_cutoff() { return -746.0; }
    static __device__ __forceinline__ double ln2_hi() { return 0x1.62e42feep-1; }
    static __device__ __forceinline__ double ln2_lo() { return 0x1.a39ef35793c76p-33; }
    static __device__ __forceinline__ double uni(uint32_t w)
    {
        return __dmul_rn(__dadd_rn(__uint2double_rn(w), 0.5), 0x1p-32);
    }
    static __device__ __forceinline__ double ang_frac(uint32_t w, bool odd)
    {
        uint32_t b = (w & 0x1fffffffu) ^ (odd ? 0x1fffffffu : 0u);
        return __dmul_rn(__dadd_rn(__uint2double_rn(b), 0.5), 0x1p-29);
    }
};

// 1/(2n+1)
template <typename T> __device__ __forceinline__ T coef_odd(int n)
{
    switch (n) {
    case 0: return T(1.0);
    case 1: return T(1.0 / 3.0);
    case 2: return T(1.0 / 5.0);
    case 3: return T(1.0 / 7.0);
    case 4: return T(1.0 / 9.0);
    case 5: return T(1.0 / 11.0);
    case 6: return T(1.0 / 13.0);
    case 7: return T(1.0 / 15.0);
    case 8: return T(1.0 / 17.0);
    case 9: return T(1.0 / 19.0);
    default: return T(1.0 / 21.0);
    }
}
// 1/n!
template <typename T> __device__ __forceinline__ T coef_invfact(int n)
{
    switch (n) {
    case 0: return T(1.0);
    case 1: return T(1.0);
    case 2: return T(1.0 / 2.0);
    case 3: return T(1.0 / 6.0);
    case 4: return T(1.0 / 24.0);
    case 5: return T(1.0 / 120.0);
    case 6: return T(1.0 / 720.0);
    case 7: return T(1.0 / 5040.0);
    case 8: return T(1.0 / 40320.0);
    case 9: return T(1.0 / 362880.0);
    case 10: return T(1.0 / 3628800.0);
    case 11: return T(1.0 / 39916800.0);
    case 12: return T(1.0 / 479001600.0);
    case 13: return T(1.0 / 6227020800.0);
    default: return T(1.0 / 87178291200.0);
    }
}
// (-1)^n/(2n+1)!
template <typename T> __device__ __forceinline__ T coef_sin(int n)
{
    switch (n) {
    case 0: return T(1.0);
    case 1: return T(-1.0 / 6.0);
    case 2: return T(1.0 / 120.0);
    case 3: return T(-1.0 / 5040.0);
    case 4: return T(1.0 / 362880.0);
    case 5: return T(-1.0 / 39916800.0);
    case 6: return T(1.0 / 6227020800.0);
    case 7: return T(-1.0 / 1307674368000.0);
    case 8: return T(1.0 / 355687428096000.0);
    default: return T(-1.0 / 121645100408832000.0);
    }
}
// (-1)^n/(2n)!
template <typename T> __device__ __forceinline__ T coef_cos(int n)
{
    switch (n) {
    case 0: return T(1.0);
    case 1: return T(-1.0 / 2.0);
    case 2: return T(1.0 / 24.0);
    case 3: return T(-1.0 / 720.0);
    case 4: return T(1.0 / 40320.0);
    case 5: return T(-1.0 / 3628800.0);
    case 6: return T(1.0 / 479001600.0);
    case 7: return T(-1.0 / 87178291200.0);
    case 8: return T(1.0 / 20922789888000.0);
    default: return T(-1.0 / 6402373705728000.0);
    }
}

// natural logarithm, x > 0 finite
template <typename T> __device__ __forceinline__ T c_ln(T x)
{
    using M = CM<T>;
    typename M::U ix = M::bits(x);
    typename M::S e = 0;
    if (ix < M::MIN_NORMAL) {
        x = M::mul(x, M::denorm_scale());
        ix = M::bits(x);
        e = -(typename M::S)M::DENORM_SHIFT;
    }
    typename M::U t = ix - M::SQRT_HALF;
    e += (typename M::S)t >> M::MANT;
    T m = M::from((t & M::MANT_MASK) + M::SQRT_HALF);
    T s = M::div(M::sub(m, T(1.0)), M::add(m, T(1.0)));
    T s2 = M::mul(s, s);
    T p = coef_odd<T>(M::LN_TERMS - 1);
#pragma unroll
    for (int n = M::LN_TERMS - 2; n >= 0; --n) p = M::fma(p, s2, coef_odd<T>(n));
    T lnm = M::mul(M::mul(T(2.0), s), p);
    return M::fma(M::from_int(e), T(0.693147180559945309417232121458), lnm);
}

// exponential for y <= 0
template <typename T> __device__ __forceinline__ T c_exp_neg(T y)
{
    using M = CM<T>;
    if (y < M::exp_cutoff()) return T(0.0);
    T n = M::rint(M::mul(y, T(1.44269504088896340735992468100)));
    T r = M::fma(-n, M::ln2_hi(), y);
    r = M::fma(-n, M::ln2_lo(), r);
    T p = coef_invfact<T>(M::EXP_DEG);
#pragma unroll
    for (int k = M::EXP_DEG - 1; k >= 0; --k) p = M::fma(p, r, coef_invfact<T>(k));
    typename M::S ni = M::to_int(n);
    typename M::S n1 = ni >> 1;
    typename M::S n2 = ni - n1;
    T s1 = M::from((typename M::U)(n1 + M::BIAS) << M::MANT);
    T s2 = M::from((typename M::U)(n2 + M::BIAS) << M::MANT);
    return M::mul(M::mul(p, s1), s2);
}

// cos(2 pi t), t given by a 32-bit word: octant (3 bits) + fraction
template <typename T> __device__ __forceinline__ T c_cos2pi(uint32_t w)
{
    using M = CM<T>;
    uint32_t o = w >> 29;
    bool odd = (o & 1u) != 0;
    bool use_sin = (((o + 1u) >> 1) & 1u) != 0;
    bool neg = (o >= 2u && o <= 5u);
    T f = M::ang_frac(w, odd);
    T a = M::mul(f, T(0.785398163397448309615660845820));
    T a2 = M::mul(a, a);
    T p = use_sin ? coef_sin<T>(M::TRIG_DEG) : coef_cos<T>(M::TRIG_DEG);
#pragma unroll
    for (int k = M::TRIG_DEG - 1; k >= 0; --k)
        p = M::fma(p, a2, use_sin ? coef_sin<T>(k) : coef_cos<T>(k));
    if (use_sin) p = M::mul(p, a);
    return neg ? -p : p;
}

// One Marsaglia-Tsang attempt.  Returns true and sets g on acceptance.
// d, c, inva are the per-cell constants: d = aa - 1/3, c = 1/sqrt(9 d), inva = 1/a when a < 1 (boost).
template <typename T>
__device__ __forceinline__ bool gamma_attempt(bool boost, T d, T c, T inva, uint4 w, T &g)
{
    using M = CM<T>;
    T u1 = M::uni(w.x);
    T x = M::mul(M::sqrt(M::mul(T(-2.0), c_ln<T>(u1))), c_cos2pi<T>(w.y));
    T v = M::fma(c, x, T(1.0));
    if (!(v > T(0.0))) return false;
    v = M::mul(M::mul(v, v), v);
    T x2 = M::mul(x, x);
    T x4 = M::mul(x2, x2);
    T u = M::uni(w.z);
    bool accept = u < M::fma(T(-0.0331), x4, T(1.0));
    if (!accept) {
        T t = M::add(M::sub(T(1.0), v), c_ln<T>(v));
        accept = c_ln<T>(u) < M::fma(T(0.5), x2, M::mul(d, t));
    }
    if (!accept) return false;
    g = M::mul(d, v);
    if (boost) {
        T ub = M::uni(w.w);
        g = M::mul(g, c_exp_neg<T>(M::mul(c_ln<T>(ub), inva)));
    }
    return true;
}

// Attempt with the squeeze test only: true (and g) when `u < 1 - 0.0331 x^4` accepts the candidate,
// false when the candidate needs the full test or was rejected -- the caller then runs c_gamma
// for the cell from attempt 0 (same operations, same result).
template <typename T>
__device__ __forceinline__ bool gamma_attempt_squeeze(bool boost, T d, T c, T inva, uint4 w, T &g)
{
    using M = CM<T>;
    T u1 = M::uni(w.x);
    T x = M::mul(M::sqrt(M::mul(T(-2.0), c_ln<T>(u1))), c_cos2pi<T>(w.y));
    T v = M::fma(c, x, T(1.0));
    T x2 = M::mul(x, x);
    T x4 = M::mul(x2, x2);
    T u = M::uni(w.z);
    if (!(v > T(0.0)) || !(u < M::fma(T(-0.0331), x4, T(1.0)))) return false;
    v = M::mul(M::mul(v, v), v);
    g = M::mul(d, v);
    if (boost) {
        T ub = M::uni(w.w);
        g = M::mul(g, c_exp_neg<T>(M::mul(c_ln<T>(ub), inva)));
    }
    return true;
}

// ---------------------------------------------------------------------------------------
// The squeeze-only attempt for W cells at once (fp32).  Operation for operation the same arithmetic as
// gamma_attempt_squeeze<float> -- every cell sees exactly the scalar sequence, so the values are bit-identical --
// but written with the W cells side by side: the polynomial and Philox chains of one cell are strictly
// dependent, and a warp that works on one cell at a time spends most of its cycles waiting on fixed-latency
// results (ncu: stall_wait 28 % of the fused z kernel).  W independent chains in one basic block let the
// scheduler overlap them.  Inputs of the logarithms are uniforms (n + 1/2) 2^-23 >= 2^-24: never denormal, so
// the denormal branch of c_ln is not needed here.
// ---------------------------------------------------------------------------------------
template <int W> __device__ __forceinline__ void c_ln_normal_w(const float (&x)[W], float (&out)[W])
{
    using M = CM<float>;
    float m[W], s[W], s2[W], p[W], ef[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t t = M::bits(x[w]) - M::SQRT_HALF;
        ef[w] = M::from_int((int32_t)t >> M::MANT);
        m[w] = M::from((t & M::MANT_MASK) + M::SQRT_HALF);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) s[w] = M::div(M::sub(m[w], 1.0f), M::add(m[w], 1.0f));
#pragma unroll
    for (int w = 0; w < W; ++w) { s2[w] = M::mul(s[w], s[w]); p[w] = coef_odd<float>(M::LN_TERMS - 1); }
#pragma unroll
    for (int n = M::LN_TERMS - 2; n >= 0; --n)
#pragma unroll
        for (int w = 0; w < W; ++w) p[w] = M::fma(p[w], s2[w], coef_odd<float>(n));
#pragma unroll
    for (int w = 0; w < W; ++w)
        out[w] = M::fma(ef[w], 0.693147180559945309417232121458f, M::mul(M::mul(2.0f, s[w]), p[w]));
}

template <int W> __device__ __forceinline__ void c_exp_neg_w(const float (&y)[W], float (&out)[W])
{
    using M = CM<float>;
    float n[W], r[W], p[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        n[w] = M::rint(M::mul(y[w], 1.44269504088896340735992468100f));
        r[w] = M::fma(-n[w], M::ln2_hi(), y[w]);
        r[w] = M::fma(-n[w], M::ln2_lo(), r[w]);
        p[w] = coef_invfact<float>(M::EXP_DEG);
    }
#pragma unroll
    for (int k = M::EXP_DEG - 1; k >= 0; --k)
#pragma unroll
        for (int w = 0; w < W; ++w) p[w] = M::fma(p[w], r[w], coef_invfact<float>(k));
#pragma unroll
    for (int w = 0; w < W; ++w) {
        // below the cutoff the scalar version returns 0 before it forms the scale factors; clamp n so that the
        // exponent arithmetic stays in range on the discarded path
        const bool tiny = y[w] < M::exp_cutoff();
        const int32_t ni = tiny ? 0 : M::to_int(n[w]);
        const int32_t n1 = ni >> 1, n2 = ni - n1;
        const float s1 = M::from((uint32_t)(n1 + M::BIAS) << M::MANT);
        const float s2 = M::from((uint32_t)(n2 + M::BIAS) << M::MANT);
        const float v = M::mul(M::mul(p[w], s1), s2);
        out[w] = tiny ? 0.0f : v;
    }
}

template <int W> __device__ __forceinline__ void c_cos2pi_w(const uint32_t (&wd)[W], float (&out)[W])
{
    using M = CM<float>;
    float a[W], a2[W], p[W];
    bool use_sin[W], neg[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const uint32_t o = wd[w] >> 29;
        const bool odd = (o & 1u) != 0;
        use_sin[w] = (((o + 1u) >> 1) & 1u) != 0;
        neg[w] = (o >= 2u && o <= 5u);
        a[w] = M::mul(M::ang_frac(wd[w], odd), 0.785398163397448309615660845820f);
        a2[w] = M::mul(a[w], a[w]);
        p[w] = use_sin[w] ? coef_sin<float>(M::TRIG_DEG) : coef_cos<float>(M::TRIG_DEG);
    }
#pragma unroll
    for (int k = M::TRIG_DEG - 1; k >= 0; --k)
#pragma unroll
        for (int w = 0; w < W; ++w) p[w] = M::fma(p[w], a2[w], use_sin[w] ? coef_sin<float>(k) : coef_cos<float>(k));
#pragma unroll
    for (int w = 0; w < W; ++w) {
        if (use_sin[w]) p[w] = M::mul(p[w], a[w]);
        out[w] = neg[w] ? -p[w] : p[w];
    }
}

// done[w] = the squeeze accepted cell w's attempt-0 candidate, g[w] its Gamma value (boost applied when inva > 0)
template <int W>
__device__ __forceinline__ void gamma_attempt_squeeze_w(const float (&d)[W], const float (&c)[W], const float (&inva)[W],
                                                        const uint4 (&rnd)[W], bool (&done)[W], float (&g)[W])
{
    using M = CM<float>;
    float u1[W], l1[W], cs[W], x[W], v[W], ub[W], lb[W], eb[W];
    uint32_t ang[W];
#pragma unroll
    for (int w = 0; w < W; ++w) { u1[w] = M::uni(rnd[w].x); ang[w] = rnd[w].y; ub[w] = M::uni(rnd[w].w); }
    c_ln_normal_w<W>(u1, l1);
    c_cos2pi_w<W>(ang, cs);
    c_ln_normal_w<W>(ub, lb);
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = M::mul(M::sqrt(M::mul(-2.0f, l1[w])), cs[w]);
#pragma unroll
    for (int w = 0; w < W; ++w) lb[w] = M::mul(lb[w], inva[w]);
    c_exp_neg_w<W>(lb, eb);
#pragma unroll
    for (int w = 0; w < W; ++w) {
        v[w] = M::fma(c[w], x[w], 1.0f);
        const float x2 = M::mul(x[w], x[w]);
        const float x4 = M::mul(x2, x2);
        const float u = M::uni(rnd[w].z);
        done[w] = (v[w] > 0.0f) && (u < M::fma(-0.0331f, x4, 1.0f));
        const float v3 = M::mul(M::mul(v[w], v[w]), v[w]);
        const float gv = M::mul(d[w], v3);
        g[w] = inva[w] > 0.0f ? M::mul(gv, eb[w]) : gv;
    }
}

template <typename T> __device__ __forceinline__ void gamma_setup(T a, bool &boost, T &d, T &c, T &inva)
{
    using M = CM<T>;
    boost = a < T(1.0);
    T aa = boost ? M::add(a, T(1.0)) : a;
    d = M::sub(aa, T(1.0 / 3.0));
    c = M::div(T(1.0), M::sqrt(M::mul(T(9.0), d)));
    inva = boost ? M::div(T(1.0), a) : T(0.0);
}

// Complete draw for one cell (loops over attempts).
template <typename T>
__device__ __forceinline__ T c_gamma(T a, uint32_t k0, uint32_t k1, unsigned long long cell,
                                     uint32_t sweep, uint32_t stream)
{
    bool boost;
    T d, c, inva, g = T(0.0);
    gamma_setup<T>(a, boost, d, c, inva);
    for (uint32_t attempt = 0;; ++attempt) {
        uint4 w = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), sweep, (stream << 24) | attempt, k0, k1);
        if (gamma_attempt<T>(boost, d, c, inva, w, g)) return g;
    }
}

// ---------------------------------------------------------------------------------------
// Poisson draw of the Polya-urn Phi sampler (reference: types/PolyaUrnDirichletFixedCoeffPoisson.java:17-44 draws
// X_w ~ Poisson(beta + n_wk) per cell through types/PoissonFixedCoeffSampler.java:45-51: an alias table over the
// Poisson pmf truncated to [0, 2L) when n < L, and round(sqrt(lambda) N(0,1) + lambda) -- the normal approximation
// of PolyaUrnDirichlet.java:102-107 -- from L on; L = alias_poisson_threshold).  Contract: one Philox block per
// cell, fp64, libm-free:
//   n <  L   inversion by sequential search over the same truncated pmf with ONE 52-bit uniform:
//            p = exp(-lambda), F = p, k = 0; while (u > F && k < 2L-1) { k++; p *= lambda / k; F += p; }
//            (exactly the distribution the alias table encodes; a zero-count cell, lambda = beta << 1, stops at
//            k = 0 with probability exp(-beta) after one comparison)
//   n >= L   floor(sqrt(lambda) * x + lambda + 1/2) with x = sqrt(-2 ln u1) cos(2 pi t) (Java's Math.round), at least 0
// p0 = exp(-beta) is passed in so that the zero-count cells do not re-evaluate it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double uniform52(uint32_t hi, uint32_t lo)
{
    const unsigned long long n = ((unsigned long long)hi << 20) | (unsigned long long)(lo >> 12);
    return __dmul_rn(__dadd_rn(__ull2double_rn(n), 0.5), 0x1p-52);
}

__device__ __forceinline__ int32_t c_poisson(double beta, int32_t n, int32_t L, double p0, uint4 w)
{
    const double lambda = __dadd_rn(beta, __int2double_rn(n));
    if (n < L) {
        const double u = uniform52(w.x, w.y);
        double p = n == 0 ? p0 : c_exp_neg<double>(-lambda);
        double F = p;
        int32_t k = 0;
        const int32_t kmax = 2 * L - 1;
        while (u > F && k < kmax) {
            ++k;
            p = __dmul_rn(p, __ddiv_rn(lambda, __int2double_rn(k)));
            F = __dadd_rn(F, p);
        }
        return k;
    }
    const double u1 = CM<double>::uni(w.x);
    const double x = __dmul_rn(__dsqrt_rn(__dmul_rn(-2.0, c_ln<double>(u1))), c_cos2pi<double>(w.y));
    const double v = __dadd_rn(__fma_rn(__dsqrt_rn(lambda), x, lambda), 0.5);
    const double r = floor(v);
    return r > 0.0 ? (int32_t)__double2int_rz(r) : 0;
}

}  // namespace ldagpu
