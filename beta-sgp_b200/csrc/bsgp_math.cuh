// Scalar helpers shared by every kernel of the beta-SGP path.
//
// The headers under csrc/ compile in two modes:
//   * nvcc (sm_100a): everything is __device__ code; this is the product.
//   * g++ with -DBSGP_HOST_EMUL: a single-"thread" emulation used ONLY by tests/host_emul to check
//     index arithmetic and controller logic on machines without a GPU.  It is never linked into
//     libbsgp.so and the Python package cannot reach it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef BSGP_HOST_EMUL
#define BSGP_DEV inline
#define BSGP_NOINLINE inline
#else
#define BSGP_DEV __device__ __forceinline__
#define BSGP_NOINLINE __device__ __noinline__
#endif

namespace bsgp {

// Base of the kernel's dynamic shared memory.  Non-inlined device functions receive shared-memory
// buffers as byte OFFSETS from this base and rebuild the pointer with smem_at(): the compiler then
// knows the address space and emits 32-bit LDS/STS addressing instead of generic 64-bit loads.
#ifdef BSGP_HOST_EMUL
static unsigned char* g_emul_smem = nullptr;      // set by the emulation before each call
BSGP_DEV unsigned char* dyn_smem() { return g_emul_smem; }
#else
BSGP_DEV unsigned char* dyn_smem() {
    extern __shared__ __align__(128) unsigned char bsgp_dyn_smem[];
    return bsgp_dyn_smem;
}
#endif
template <typename P> BSGP_DEV P* smem_at(unsigned off) { return reinterpret_cast<P*>(dyn_smem() + off); }

// ---------------------------------------------------------------------------------------------
// "numpy-faithful" arithmetic: numpy evaluates every ufunc separately, so a*b+c is two roundings.
// nvcc would contract it into one FMA; the *_rn intrinsics are never contracted.  Used in the
// elementwise parts of the solver (sgp.py:311,329-334,355-365); the FFT butterflies use plain
// operators and may contract.
// ---------------------------------------------------------------------------------------------
#ifdef BSGP_HOST_EMUL
BSGP_DEV double nmul(double a, double b) { volatile double r = a * b; return r; }
BSGP_DEV double nadd(double a, double b) { volatile double r = a + b; return r; }
BSGP_DEV double nsub(double a, double b) { volatile double r = a - b; return r; }
BSGP_DEV double ndiv(double a, double b) { volatile double r = a / b; return r; }
BSGP_DEV float nmul(float a, float b) { volatile float r = a * b; return r; }
BSGP_DEV float nadd(float a, float b) { volatile float r = a + b; return r; }
BSGP_DEV float nsub(float a, float b) { volatile float r = a - b; return r; }
BSGP_DEV float ndiv(float a, float b) { volatile float r = a / b; return r; }
#else
BSGP_DEV double nmul(double a, double b) { return __dmul_rn(a, b); }
BSGP_DEV double nadd(double a, double b) { return __dadd_rn(a, b); }
BSGP_DEV double nsub(double a, double b) { return __dsub_rn(a, b); }
BSGP_DEV double ndiv(double a, double b) { return __ddiv_rn(a, b); }
BSGP_DEV float nmul(float a, float b) { return __fmul_rn(a, b); }
BSGP_DEV float nadd(float a, float b) { return __fadd_rn(a, b); }
BSGP_DEV float nsub(float a, float b) { return __fsub_rn(a, b); }
BSGP_DEV float ndiv(float a, float b) { return __fdiv_rn(a, b); }
#endif

// ---------------------------------------------------------------------------------------------
// pow(x, y) for the beta-divergence terms den^(beta-1), gn^beta (sgp.py:457-458,495,499).
//
// The CUDA library pow() is a real function call (__internal_accurate_pow is not inlined): inside a
// pixel loop every call spills the loop's live registers to local memory.  pow_inline() is a
// call-free replacement for positive, finite, normal x and results far from over/underflow:
//   log(x) in double-double  (x = m 2^e, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh((m-1)/(m+1)),
//   the quotient carried as head + tail, the odd series tail in plain double),
//   t = y log(x) in double-double, exp(t) by Cody-Waite reduction and a degree-13 polynomial.
// Measured against 80-bit long-double pow on the CPU (tests/test_host_logic.py): <= 1 ulp (CUDA's pow: 2 ulp).
// Anything else (x <= 0, NaN, Inf, subnormal, |y log x| > 600) takes the library function on a cold path.
// ---------------------------------------------------------------------------------------------
#ifdef BSGP_HOST_EMUL
BSGP_DEV double bits_hi_lo(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &u, 8); return d; }
BSGP_DEV int hi_word(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
BSGP_DEV int lo_word(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u & 0xffffffffu); }
BSGP_DEV double fast_rcp(double a) { return 1.0 / a; }
BSGP_DEV double mfma(double a, double b, double c) { return fma(a, b, c); }
BSGP_DEV double mrint(double a) { return nearbyint(a); }
inline double pow_slow(double a, double b) { return pow(a, b); }
#else
BSGP_DEV double bits_hi_lo(int hi, int lo) { return __hiloint2double(hi, lo); }
BSGP_DEV int hi_word(double d) { return __double2hiint(d); }
BSGP_DEV int lo_word(double d) { return __double2loint(d); }
BSGP_DEV double fast_rcp(double a) {          // ~2^-20 seed + two Newton steps: relative error ~2^-52, a in [1.7, 2.5]
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
BSGP_DEV double mfma(double a, double b, double c) { return fma(a, b, c); }
BSGP_DEV double mrint(double a) { return rint(a); }
static __device__ __noinline__ double pow_slow(double a, double b) { return pow(a, b); }
#endif

// polynomial coefficients of pow_inline
#define BSGP_LOG_COEFFS {2.0 / 23.0, 2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0}
#define BSGP_EXP_COEFFS {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, \
                         1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5}
// (as constexpr locals the coefficients become immediates; a __constant__ table measured 15 % slower)
#define BSGP_POW_TABLES constexpr double kLogC[11] = BSGP_LOG_COEFFS; constexpr double kExpC[12] = BSGP_EXP_COEFFS;

BSGP_DEV double pow_inline(double x, double y) {
    BSGP_POW_TABLES
    if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return pow_slow(x, y);
    // ---- log(x) = (lh, ll)
    int hx = hi_word(x);
    int e = (hx >> 20) - 1023;
    int mh = (hx & 0x000fffff) | 0x3ff00000;
    if (mh >= 0x3ff6a09e) { mh -= 0x00100000; e += 1; }          // m >= sqrt(2) (to 21 bits) -> m / 2
    const double m = bits_hi_lo(mh, lo_word(x));
    const double f = m - 1.0;                                      // exact
    const double uh = m + 1.0;
    const double ul = (m - (uh - 1.0));                            // uh + ul = m + 1 exactly
    const double r = fast_rcp(uh);
    const double qh = f * r;
    double rem = mfma(-qh, uh, f);
    rem = mfma(-qh, ul, rem);
    const double ql = rem * r;                                     // q = qh + ql = f / (m + 1)
    const double q2 = qh * qh;
    // 2 atanh(q) - 2q = 2 q^3 (1/3 + q^2/5 + ... + q^20/23)
    double p = kLogC[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = mfma(p, q2, kLogC[i]);
    const double tail = (p * q2) * qh;
    const double ed = (double)e;
    const double a = ed * 6.93147180369123816490e-01;              // exact: ln2_hi has 21 trailing zero bits
    const double b2 = qh + qh;
    const double sh = a + b2;                                      // two-sum
    const double bb = sh - a;
    const double se = (a - (sh - bb)) + (b2 - bb);
    const double lo = se + (mfma(ed, 1.90821492927058770002e-10, ql + ql) + tail);
    const double lh = sh + lo;
    const double ll = lo - (lh - sh);
    // ---- t = y * log(x) = (th, tl)
    const double th = y * lh;
    const double tl = mfma(y, ll, mfma(y, lh, -th));
    if (!(th > -600.0 && th < 600.0)) return pow_slow(x, y);
    // ---- exp(th + tl)
    const double n = mrint(th * 1.4426950408889634);
    double rr = mfma(-n, 6.93147180559945286227e-01, th);
    rr = mfma(-n, 2.31904681384629955842e-17, rr) + tl;
    // exp(rr) = 1 + (rr + rr^2 (1/2 + rr/6 + ... + rr^11/13!)); the last addition carries the only 1/2-ulp rounding
    double z = kExpC[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) z = mfma(z, rr, kExpC[i]);
    double sm = mfma(z * rr, rr, rr);                              // exp(rr) - 1
    z = 1.0 + sm;
    return bits_hi_lo(hi_word(z) + ((int)n << 20), lo_word(z));
}

BSGP_DEV double mpow(double a, double b) { return pow_inline(a, b); }
BSGP_DEV float mpow(float a, float b) { return powf(a, b); }
BSGP_DEV double mlog(double a) { return log(a); }
BSGP_DEV float mlog(float a) { return logf(a); }
// The same two functions as ONE shared copy each, for the kernels of small images (DeviceCtxSmall, bsgp_device.cuh): those
// are bound by instruction fetch, every pixel loop runs once or twice per pass, and an inlined pow() per pixel of every
// pair and row makes the objective phases the largest functions of an iteration.  CALL is a compile-time property of
// the execution context, so the large-image kernels keep their call-free loops.
#ifdef BSGP_HOST_EMUL
inline double pow_call(double a, double b) { return pow_inline(a, b); }
inline double log_call(double a) { return log(a); }
#else
static __device__ __noinline__ double pow_call(double a, double b) { return pow_inline(a, b); }
static __device__ __noinline__ double log_call(double a) { return log(a); }
#endif
template <bool CALL> BSGP_DEV double mpow_sel(double a, double b) { return CALL ? pow_call(a, b) : pow_inline(a, b); }
template <bool CALL> BSGP_DEV float mpow_sel(float a, float b) { return powf(a, b); }
template <bool CALL> BSGP_DEV double mlog_sel(double a) { return CALL ? log_call(a) : log(a); }
template <bool CALL> BSGP_DEV float mlog_sel(float a) { return logf(a); }
BSGP_DEV double mexp(double a) { return exp(a); }
BSGP_DEV float mexp(float a) { return expf(a); }
BSGP_DEV double mabs(double a) { return fabs(a); }
BSGP_DEV float mabs(float a) { return fabsf(a); }
BSGP_DEV double msqrt(double a) { return sqrt(a); }
BSGP_DEV float msqrt(float a) { return sqrtf(a); }

// Python's builtin max(a, b) / min(a, b) on two scalars: returns `a` unless b compares greater /
// smaller (so a NaN in `b` never wins and a NaN in `a` always does).  sgp.py:366-375.
template <typename T> BSGP_DEV T py_max(T a, T b) { return (b > a) ? b : a; }
template <typename T> BSGP_DEV T py_min(T a, T b) { return (b < a) ? b : a; }
// np.max([a, b]) / np.min([a, b]): NaN propagates.  flux_conserve_proj.py:44,66,116,118,131,133.
template <typename T> BSGP_DEV T np_max2(T a, T b) { return (a != a || b != b) ? (a + b) : ((a > b) ? a : b); }
template <typename T> BSGP_DEV T np_min2(T a, T b) { return (a != a || b != b) ? (a + b) : ((a < b) ? a : b); }
template <typename T> BSGP_DEV bool is_finite(T a) { return (a - a) == (T)0; }

template <typename T> struct Eps;
template <> struct Eps<double> { static BSGP_DEV double v() { return 2.220446049250313e-16; } };
// the reference is fp64 throughout; the fp32 mode keeps the reference's constants where they are
// representable and uses float epsilon where the constant is "machine epsilon"
template <> struct Eps<float> { static BSGP_DEV float v() { return 1.1920929e-07f; } };

// Compensated (two-sum) accumulator.  The beta-divergence is the difference of three sums that are
// each 1e2..1e5 times larger than the result (sgp.py:457-458), so the accumulation error of every
// partial sum is kept at the 1-ulp level regardless of how many pixels a thread owns.
struct KSum {
    double s, c;
    BSGP_DEV void clear() { s = 0.0; c = 0.0; }
    BSGP_DEV void add(double x) {
        const double t = nadd(s, x);
        const double bp = nsub(t, s);
        c = nadd(c, nadd(nsub(s, nsub(t, bp)), nsub(x, bp)));
        s = t;
    }
    BSGP_DEV double value() const { return nadd(s, c); }
};

// ---------------------------------------------------------------------------------------------
// complex
// ---------------------------------------------------------------------------------------------
template <typename T> struct alignas(2 * sizeof(T)) cplx {
    T re, im;
};

template <typename T> BSGP_DEV cplx<T> cmake(T re, T im) { cplx<T> r; r.re = re; r.im = im; return r; }
template <typename T> BSGP_DEV cplx<T> cadd(cplx<T> a, cplx<T> b) { return cmake<T>(a.re + b.re, a.im + b.im); }
template <typename T> BSGP_DEV cplx<T> csub(cplx<T> a, cplx<T> b) { return cmake<T>(a.re - b.re, a.im - b.im); }
template <typename T> BSGP_DEV cplx<T> cmul(cplx<T> a, cplx<T> b) {
    return cmake<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// a * conj(b)
template <typename T> BSGP_DEV cplx<T> cmulc(cplx<T> a, cplx<T> b) {
    return cmake<T>(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
template <typename T> BSGP_DEV cplx<T> cconj(cplx<T> a) { return cmake<T>(a.re, -a.im); }
template <typename T> BSGP_DEV cplx<T> cscale(cplx<T> a, T s) { return cmake<T>(a.re * s, a.im * s); }
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename T> BSGP_DEV cplx<T> crot(cplx<T> a) {
    return INV ? cmake<T>(-a.im, a.re) : cmake<T>(a.im, -a.re);
}

}  // namespace bsgp
