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
    int dft_n;      // 0: radix stages over all n slots.  Otherwise the transform is a DENSE DFT of length dft_n <= n over
                    // slots [0, dft_n), frequencies in natural order (small sides that are not a power of two, see
                    // dense_dft_batch); the slots beyond keep whatever they hold
};

// transform length, as opposed to the number of workspace slots pl.n
BSGP_DEV int fft_len(const FftPlan& pl) { return pl.dft_n ? pl.dft_n : pl.n; }

BSGP_DEV int fpad(int i, int ps) { return i + (i >> ps); }

// position (unpadded) of frequency k after the forward stages
BSGP_DEV int pos_of_freq(const FftPlan& pl, int k) {
    if (pl.dft_n) return k;
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

// ---------------------------------------------------------------------------------------------
// Dense DFT for short transforms whose length is not a power of two (the reference's 31 x 31 cut-outs,
// application_sgp_star_stamps.py:24,58): X[k] = sum_j x[j] w^(j k), w = exp(-+2 pi i / n), n = pl.dft_n <= 32, straight
// from the table tw[m] = exp(-2 pi i m / n) (the exponent j k mod n is carried incrementally, so every twiddle is an
// exact table entry).  In place like the radix stages: every thread first accumulates its outputs in registers, one
// barrier, then stores them.  n^2 complex multiply-adds per transform instead of ~n log n, but the image
// keeps its own size: all per-image arrays of a 31 x 31 solve stay resident in shared memory, where the alternative
// (linear convolution on a 64 x 64 grid and a fold, "wrapped plans") works on arrays four times as large.
// Ends with a barrier.
// ---------------------------------------------------------------------------------------------
// One task = one transform f and one frequency k <= n / 2; it delivers X[k] and X[n - k] from the same products:
//   C = sum_j x[j] cos(2 pi j k / n),  S = sum_j x[j] sin(2 pi j k / n)   (complex x: four real multiply-adds per j)
//   forward X[k] = C - i S, X[n - k] = C + i S;  inverse (unnormalised) the other way round.
// R = tasks per thread (compile time, so the accumulators stay in registers); a slot beyond the last task repeats it.
template <bool INV, int R, class Ctx, typename T>
BSGP_DEV void dense_dft_tasks(Ctx& ctx, cplx<T>* ws, int nfft, int fstride, int n, int ps, const cplx<T>* tw) {
    const int per = (n >> 1) + 1, total = nfft * per;
    cplx<T> C[R], S[R];
    cplx<T>* src[R];
    int kk[R], idx[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
        int t = ctx.tid + u * ctx.nt;
        t = t < total ? t : total - 1;                  // clamped, not guarded: the tile stays in registers
        const int f = t / per;
        kk[u] = t - f * per; idx[u] = 0;
        src[u] = ws + f * fstride;
        C[u] = S[u] = cmake<T>(0, 0);
    }
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
        const int pj = fpad(j, ps);
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const cplx<T> v = src[u][pj], w = tw[idx[u]];        // w = (cos, -sin)
            C[u].re += v.re * w.re; C[u].im += v.im * w.re;
            S[u].re -= v.re * w.im; S[u].im -= v.im * w.im;
            idx[u] += kk[u];
            idx[u] = idx[u] >= n ? idx[u] - n : idx[u];
        }
    }
    ctx.sync();
#pragma unroll
    for (int u = 0; u < R; ++u) {
        if (ctx.tid + u * ctx.nt < total) {
            // C - i S = (C.re + S.im, C.im - S.re),  C + i S = (C.re - S.im, C.im + S.re)
            const cplx<T> lo = cmake<T>(C[u].re + S[u].im, C[u].im - S[u].re), hi = cmake<T>(C[u].re - S[u].im, C[u].im + S[u].re);
            src[u][fpad(kk[u], ps)] = INV ? hi : lo;
            if (kk[u] != 0 && 2 * kk[u] != n) src[u][fpad(n - kk[u], ps)] = INV ? lo : hi;
        }
    }
    ctx.sync();
}

template <bool INV, class Ctx, typename T>
BSGP_DEV void dense_dft_batch(Ctx& ctx, cplx<T>* ws, int nfft, int fstride, const FftPlan& pl, const cplx<T>* tw) {
    const int n = pl.dft_n, ps = pl.pad_shift;
#ifdef BSGP_HOST_EMUL
    // one emulated thread: plain out-of-place sums per transform
    cplx<T> tmp[64];
    for (int f = 0; f < nfft; ++f) {
        cplx<T>* a = ws + f * fstride;
        for (int k = 0; k < n; ++k) {
            cplx<T> acc = cmake<T>(0, 0);
            int idx = 0;
            for (int j = 0; j < n; ++j) {
                const cplx<T> v = a[fpad(j, ps)], w = tw[idx];
                acc = cadd(acc, INV ? cmulc(v, w) : cmul(v, w));
                idx += k; if (idx >= n) idx -= n;
            }
            tmp[k] = acc;
        }
        for (int k = 0; k < n; ++k) a[fpad(k, ps)] = tmp[k];
    }
#else
    const int total = nfft * ((n >> 1) + 1);
    if (total <= ctx.nt) dense_dft_tasks<INV, 1>(ctx, ws, nfft, fstride, n, ps, tw);
    else if (total <= 2 * ctx.nt) dense_dft_tasks<INV, 2>(ctx, ws, nfft, fstride, n, ps, tw);
    else {
        // more tasks than two per thread (narrow CTAs): groups of whole transforms, so that no group overwrites the inputs
        // of a later one (a transform has at most 17 tasks; every CTA of this library has at least 128 threads)
        const int per = (n >> 1) + 1, fpg = (2 * ctx.nt) / per;
        for (int f0 = 0; f0 < nfft; f0 += fpg)
            dense_dft_tasks<INV, 2>(ctx, ws + f0 * fstride, nfft - f0 < fpg ? nfft - f0 : fpg, fstride, n, ps, tw);
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// The same dense DFT on the fp64 tensor cores (mma.sync m8n8k4, DMMA).  A batch of nfft complex transforms of length n
// is one real matrix product  Out[2 f + part][2 k + t] = sum_j In[2 f + part][j] * Wt[j][2 k + t]  with
//   In[2 f + part][j] = re (part 0) or im (part 1) of element j of transform f   (M = 2 nfft rows, K = n),
//   Wt[j][2 k + t]    = cos (t 0) or sin (t 1) of 2 pi j k / n, k <= n / 2         (N = 2 (n / 2 + 1) columns),
// i.e. the sums C and S of dense_dft_tasks for every (f, k), re and im parts in neighbouring rows.  An 8 x 8 output tile
// takes ceil(n / 4) DMMA instructions; the B operand comes from a table in shared memory that is laid out in fragment
// order (fill_dense_table: one coalesced 8-byte load per lane and step), the A operand straight from the workspace.
// A lane ends up with C and S of one (f, part, k); its partner for the other part is lane ^ 4, so one shuffle completes
//   X[k] = C - i S = (Cr + Si, Ci - Sr),   X[n - k] = C + i S = (Cr - Si, Ci + Sr)      (forward; swapped for the inverse).
// 31-point transforms: 16 tiles x 8 steps = 128 DMMA per batch of 16 transforms, against 256 threads x 31 steps x 14
// instructions of the scalar version, and an eighth of its shared-memory traffic (ncu: the scalar version spent 36 % of
// a 31 x 31 solve in this function at 60 % shared-memory pipe utilisation, 2.4 G bank-conflict wavefronts).
// ---------------------------------------------------------------------------------------------
constexpr int kDenseTilesPerWarp = 4;          // 16 tiles over the four warps of the narrowest CTA

// fragment-order table for a grid of side ng (table of (ng / 8) x (ng / 4) x 32 doubles = 8 ng^2 bytes): entry
// [(nt * (ng / 4) + s) * 32 + lane] = Wt[4 s + (lane & 3)][8 nt + (lane >> 2)], zero outside j < n, k <= n / 2
template <class Ctx, typename T> BSGP_DEV void fill_dense_table(Ctx& ctx, int ng, int n, const cplx<T>* tw, double* dst) {
    const int S = ng >> 2, total = (ng >> 3) * S * 32;
    for (int e = ctx.tid; e < total; e += ctx.nt) {
        const int lane = e & 31, s = (e >> 5) % S, nt = (e >> 5) / S;
        const int j = 4 * s + (lane & 3), nn = 8 * nt + (lane >> 2), k = nn >> 1;
        double v = 0.0;
        if (j < n && 2 * k <= n) {
            const cplx<T> w = tw[(j * k) % n];                  // (cos, -sin)
            v = (nn & 1) ? -(double)w.im : (double)w.re;
        }
        dst[e] = v;
    }
}

#ifndef BSGP_HOST_EMUL
// Preconditions (checked by fft_batch): grid side pl.n = 16 or 32 with padding period 8 (pad_shift 3), nfft a multiple of 4,
// whole warps, at most kDenseTilesPerWarp tiles per warp.  The step loop is unrolled over the grid's 8 (or 4) steps so that
// every address is base + immediate: fpad(4 s + jl) = 4 s + jl + (s >> 1) for a padding period of 8.
// FAST (compile time): the transform fills the grid's last step (n > grid - 4: the 31-point case) and the tiles divide evenly
// among the warps, TPW each: no step or tile predicates, only the last step's A operand is guarded.
template <bool INV, int TPW, bool FAST, class Ctx>
BSGP_DEV void dense_dft_mma(Ctx& ctx, cplx<double>* ws, int nfft, int fstride, const FftPlan& pl, const double* tab) {
    const int n = pl.dft_n;
    const int lane = ctx.tid & 31, warp = ctx.tid >> 5, nwarps = ctx.nt >> 5;
    const int per = (n >> 1) + 1;
    const int lgNT = pl.log2n - 3, tiles = (nfft >> 2) << lgNT;          // 8 x 8 tiles: 4 frequencies (cos, sin) x 4 transforms (re, im); all of the table's column tiles
    const int S = FAST ? 8 : (n + 3) >> 2, Sfull = FAST ? 8 : pl.n >> 2;
    double* wsd = reinterpret_cast<double*>(ws);
    const int part = (lane >> 2) & 1, jl = lane & 3;
    double c0[TPW], c1[TPW];
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        c0[i] = 0.0; c1[i] = 0.0;
        const int tile = warp + i * nwarps;                   // tile = nt + NT * mt
        if (FAST || tile < tiles) {
            const int mt = tile >> lgNT, nt = tile & ((1 << lgNT) - 1);
            const double* arow = wsd + 2 * ((4 * mt + (lane >> 3)) * fstride + jl) + part;
            const double* brow = tab + (nt * Sfull) * 32 + lane;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                if (FAST || s < S) {
                    double a = arow[2 * (4 * s + (s >> 1))];
                    if (!FAST || s == 7) a = (4 * s + jl < n) ? a : 0.0;           // a slot beyond the transform may hold anything
                    const double b = brow[s * 32];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
                }
            }
        }
    }
    ctx.sync();                                               // every warp has read its inputs
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int tile = warp + i * nwarps;
        if (FAST || tile < tiles) {                           // warp-uniform
            const int mt = tile >> lgNT, nt = tile & ((1 << lgNT) - 1);
            const double other = __shfl_xor_sync(0xffffffffu, c1[i], 4);     // the S sum of the other part
            const double lo = part ? c0[i] - other : c0[i] + other;
            const double hi = part ? c0[i] + other : c0[i] - other;
            const int k = 4 * nt + jl;
            double* orow = wsd + 2 * ((4 * mt + (lane >> 3)) * fstride) + part;
            if (k < per) {
                orow[2 * fpad(k, 3)] = INV ? hi : lo;
                if (k != 0 && 2 * k != n) orow[2 * fpad(n - k, 3)] = INV ? lo : hi;
            }
        }
    }
    ctx.sync();
}
#endif

// nfft transforms of length pl.n, transform f at ws + f*fstride (padded layout), ws = shared memory at byte
// offset ws_off.
// Ends with a barrier.
// Not inlined: the solver runs five convolutions, all sharing one copy of each direction.
// tw_off: byte offset of the full twiddle table in shared memory (kNoSmem: read it through the generic pointer tw).
// GEN: the plan may ask for a dense DFT instead (embedded plans only, see row_geom in bsgp_conv.cuh).
template <bool INV, bool GEN = false, class Ctx, typename T>
BSGP_NOINLINE void fft_batch(Ctx ctx, unsigned ws_off, int nfft, int fstride, const FftPlan& pl, const cplx<T>* tw, unsigned tw_off) {
    cplx<T>* ws = smem_at<cplx<T>>(ws_off);
    if (GEN && pl.dft_n) {
        // dense transform: on the tensor cores where the kernel has built the fragment table (fp64 solve kernels: tw_off
        // addresses that table, not a twiddle table), else the scalar version with the twiddles in global memory
#ifndef BSGP_HOST_EMUL
        if constexpr (sizeof(T) == 8) {
            if (tw_off != kNoSmem && pl.n <= 32 && pl.pad_shift == 3 && (nfft & 3) == 0 && (ctx.nt & 31) == 0 && ((nfft >> 2) << (pl.log2n - 3)) <= kDenseTilesPerWarp * (ctx.nt >> 5)) {
                const double* tab = (const double*)smem_at<double>(tw_off);
                const int tiles = (nfft >> 2) << (pl.log2n - 3), nwarps = ctx.nt >> 5;
                const bool full = pl.n == 32 && pl.dft_n > 28;
                if (full && tiles == 2 * nwarps) dense_dft_mma<INV, 2, true>(ctx, ws, nfft, fstride, pl, tab);
                else if (full && tiles == 4 * nwarps) dense_dft_mma<INV, 4, true>(ctx, ws, nfft, fstride, pl, tab);
                else dense_dft_mma<INV, kDenseTilesPerWarp, false>(ctx, ws, nfft, fstride, pl, tab);
                return;
            }
        }
#endif
        dense_dft_batch<INV>(ctx, ws, nfft, fstride, pl, tw);
        return;
    }
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
