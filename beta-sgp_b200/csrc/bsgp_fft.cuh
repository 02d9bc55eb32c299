// In-place shared-memory FFT building blocks (power-of-two lengths, radix 2/4/8 stages).
//
// Replaces the numpy.fft (pocketfft) calls behind the reference's PSF operator,
// sgp.py:109-117 / 571-579:  real(ifftn(TF * fftn(x))).
//
// Layout.  A batch of independent 1-D transforms lives in one shared-memory workspace; transform f
// starts at f*fstride complex elements and element i sits at fpad(i) = i + (i >> pad_shift)
// (one padding slot every 2^pad_shift elements, pad_shift = log2(last radix)), which makes both the
// strided first stages and the unit-stride last stage bank-conflict free for 16-byte elements.
//
// Forward = decimation in frequency, natural order in -> digit-reversed order out; inverse = the
// exact adjoint stage sequence, digit-reversed in -> natural out.  The pointwise work between the
// two (two-for-one untangle, PSF-spectrum multiply) addresses frequencies through pos_of_freq(),
// so no reordering pass is ever executed.  Every stage task reads and writes the same R slots, so
// one block barrier per stage is enough.
#pragma once
#include "bsgp_math.cuh"

namespace bsgp {

struct FftPlan {
    int n;          // transform length (power of two, >= 4)
    int log2n;
    int nstages;
    int log2r[6];   // log2 of the radix of each forward stage
    int pad_shift;  // log2 of the last radix
    int plen;       // padded length: n + (n >> pad_shift)
};

BSGP_DEV int fpad(int i, int ps) { return i + (i >> ps); }

// position (unpadded) of frequency k after the forward stages
BSGP_DEV int pos_of_freq(const FftPlan& pl, int k) {
    int p = 0, lg = pl.log2n;
    for (int s = 0; s < pl.nstages; ++s) {
        lg -= pl.log2r[s];
        p += (k & ((1 << pl.log2r[s]) - 1)) << lg;
        k >>= pl.log2r[s];
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// register-resident DFTs.  X[q] = sum_r v[r] w^(q r), w = exp(-+2 pi i / R); result left in v[q].
// ---------------------------------------------------------------------------------------------
template <bool INV, typename T> BSGP_DEV void dft2(cplx<T>& a, cplx<T>& b) {
    cplx<T> t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <bool INV, typename T> BSGP_DEV void dft4(cplx<T>& a0, cplx<T>& a1, cplx<T>& a2, cplx<T>& a3) {
    cplx<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = crot<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// multiply by exp(-+ 2 pi i m / 16) for the constants needed by the radix-8 kernel
template <int M16, bool INV, typename T> BSGP_DEV cplx<T> mul_w16(cplx<T> a) {
    const T c1 = (T)0.92387953251128673848, s1 = (T)0.38268343236508978178, c2 = (T)0.70710678118654752440;
    // forward twiddle = (cr, -si); inverse = (cr, +si)
    T cr, si;
    if (M16 == 0) return a;
    if (M16 == 4) return crot<INV>(a);
    if (M16 == 1) { cr = c1; si = s1; }
    else if (M16 == 2) { cr = c2; si = c2; }
    else if (M16 == 3) { cr = s1; si = c1; }
    else if (M16 == 6) { cr = -c2; si = c2; }
    else /* 9 */ { cr = -c1; si = -s1; }
    if (INV) return cmake<T>(a.re * cr - a.im * si, a.re * si + a.im * cr);
    return cmake<T>(a.re * cr + a.im * si, a.im * cr - a.re * si);
}

template <bool INV, typename T> BSGP_DEV void dft8(cplx<T>* v) {
    // even / odd radix-4 halves, then the W8^k combine
    dft4<INV>(v[0], v[2], v[4], v[6]);
    dft4<INV>(v[1], v[3], v[5], v[7]);
    cplx<T> o0 = v[1], o1 = mul_w16<2, INV>(v[3]), o2 = mul_w16<4, INV>(v[5]), o3 = mul_w16<6, INV>(v[7]);
    cplx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <int LOG2R, bool INV, typename T> BSGP_DEV void dft_r(cplx<T>* v) {
    if (LOG2R == 1) dft2<INV>(v[0], v[1]);
    else if (LOG2R == 2) dft4<INV>(v[0], v[1], v[2], v[3]);
    else dft8<INV>(v);
}

// ---------------------------------------------------------------------------------------------
// one butterfly task of one stage: block length 2^lgL, radix 2^LOG2R, element j of block blk.
// forward:  V = DFT_R(v);  out[q] = V[q] * W_L^(q j)
// inverse:  V[q] = in[q] * conj(W_L^(q j));  out = IDFT_R(V)          (unnormalised)
// tw[] holds W_n^k, k in [0, n); W_L^(q j) = tw[q j (n / L)].
// ---------------------------------------------------------------------------------------------
// Addressing.  With the padding period 2^ps and every non-final stage stride S a multiple of it
// (make_fft_plan guarantees this), the R elements of a task sit at a CONSTANT padded stride:
//   fpad(p0 + r S) = fpad(p0) + r (S + S / 2^ps)      (S >= 2^ps),      fpad(p0 + r) = fpad(p0) + r   (final stage),
// so one base address per task replaces two shifts and two adds per element.
// SPLIT: tw is a two-level table [W^0 .. W^63][W^0, W^64, W^128, ...] and W^k = tw[64 + (k >> 6)] * tw[k & 63] (long transforms,
// whose full table does not fit shared memory: one extra complex multiply instead of an L2 round trip per twiddle).
template <typename T, bool SPLIT> BSGP_DEV cplx<T> twiddle_at(const cplx<T>* tw, int k) {
    if (SPLIT) return cmul(tw[64 + (k >> 6)], tw[k & 63]);
    return tw[k];
}

template <int LOG2R, bool INV, bool SPLIT, typename T>
BSGP_DEV void stage_task(cplx<T>* a, int spad, bool twiddle, int j, const cplx<T>* tw, int lg_twstep) {
    constexpr int R = 1 << LOG2R;
    cplx<T> v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = a[r * spad];
    if (!INV) {
        dft_r<LOG2R, false>(v);
        if (twiddle) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], twiddle_at<T, SPLIT>(tw, (q * j) << lg_twstep));
        }
    } else {
        if (twiddle) {
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], twiddle_at<T, SPLIT>(tw, (q * j) << lg_twstep));
        }
        dft_r<LOG2R, true>(v);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) a[r * spad] = v[r];
}

template <int LOG2R, bool INV, bool SPLIT, class Ctx, typename T>
BSGP_DEV void run_stage(Ctx& ctx, cplx<T>* ws, int nfft, int fstride, int log2n, int ps, int lgL, const cplx<T>* tw) {
    const int lg_per = log2n - LOG2R;             // butterflies per transform
    const int lgS = lgL - LOG2R;
    const int total = nfft << lg_per;
    const int lg_tw = log2n - lgL;
    const int S = 1 << lgS;
    const int spad = (lgS >= ps) ? S + (S >> ps) : S;
    for (int t = ctx.tid; t < total; t += ctx.nt) {
        const int f = t >> lg_per, u = t & ((1 << lg_per) - 1);
        const int j = u & (S - 1);
        const int p0 = ((u >> lgS) << lgL) + j;
        stage_task<LOG2R, INV, SPLIT>(ws + f * fstride + p0 + (p0 >> ps), spad, lgS > 0, j, tw, lg_tw);
    }
    ctx.sync();
}

template <bool INV, bool SPLIT, class Ctx, typename T>
BSGP_DEV void run_stages(Ctx& ctx, cplx<T>* ws, int nfft, int fstride, const FftPlan& pl, const cplx<T>* tw) {
    const int log2n = pl.log2n, ps = pl.pad_shift, ns = pl.nstages;
    int lgL = INV ? 0 : log2n;
    for (int i = 0; i < ns; ++i) {
        const int lg_r = pl.log2r[INV ? ns - 1 - i : i];
        if (INV) lgL += lg_r;
        switch (lg_r) {
            case 1: run_stage<1, INV, SPLIT>(ctx, ws, nfft, fstride, log2n, ps, lgL, tw); break;
            case 2: run_stage<2, INV, SPLIT>(ctx, ws, nfft, fstride, log2n, ps, lgL, tw); break;
            default: run_stage<3, INV, SPLIT>(ctx, ws, nfft, fstride, log2n, ps, lgL, tw); break;
        }
        if (!INV) lgL -= lg_r;
    }
}

constexpr unsigned kNoSmem = 0xffffffffu;

// nfft transforms of length pl.n, transform f at ws + f*fstride (padded layout), ws = shared memory at byte
// offset ws_off.
// Ends with a barrier.
// Not inlined: the solver runs five convolutions, all sharing one copy of each direction.
// tw_off: byte offset of the full twiddle table in shared memory (kNoSmem: read it through the generic pointer tw).
template <bool INV, class Ctx, typename T>
BSGP_NOINLINE void fft_batch(Ctx ctx, unsigned ws_off, int nfft, int fstride, const FftPlan& pl, const cplx<T>* tw, unsigned tw_off) {
    cplx<T>* ws = smem_at<cplx<T>>(ws_off);
    if (tw_off != kNoSmem) run_stages<INV, false>(ctx, ws, nfft, fstride, pl, (const cplx<T>*)smem_at<cplx<T>>(tw_off));
    else run_stages<INV, false>(ctx, ws, nfft, fstride, pl, tw);
}

// The same with the two-level twiddle table at tw_off (long transforms).  A separate function so that the common
// case above keeps its code size: the small-image kernels are instruction-fetch bound.
template <bool INV, class Ctx, typename T>
BSGP_NOINLINE void fft_batch_split(Ctx ctx, unsigned ws_off, int nfft, int fstride, const FftPlan& pl, unsigned tw_off) {
    run_stages<INV, true>(ctx, smem_at<cplx<T>>(ws_off), nfft, fstride, pl, (const cplx<T>*)smem_at<cplx<T>>(tw_off));
}

}  // namespace bsgp
