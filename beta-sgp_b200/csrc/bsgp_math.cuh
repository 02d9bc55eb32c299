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

#ifdef BSGP_HOST_EMUL
#define BSGP_DEV inline
#else
#define BSGP_DEV __device__ __forceinline__
#endif

namespace bsgp {

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

BSGP_DEV double mpow(double a, double b) { return pow(a, b); }
BSGP_DEV float mpow(float a, float b) { return powf(a, b); }
BSGP_DEV double mlog(double a) { return log(a); }
BSGP_DEV float mlog(float a) { return logf(a); }
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
