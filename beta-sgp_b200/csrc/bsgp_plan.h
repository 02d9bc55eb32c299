// Host-side planning shared by libbsgp (bsgp_kernels.cu) and the test-only emulation
// (tests/host_emul): FFT stage factorisation, twiddle tables, cluster / tile geometry.
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>

#include "bsgp_conv.cuh"

namespace bsgp {

inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// Stage radices (2, 4 or 8: a radix-8 butterfly of complex doubles fits a 128-register thread): every
// stage has an in-block stride >= 8 (conflict-free 16-byte quarter-warp access) except the last,
// whose radix (8) sets the padding period.
inline bool make_fft_plan(int n, FftPlan* pl, int dft_n = 0) {
    static const int table[16][6] = {
        {0}, {0}, {0},
        {3, 0},              // 8
        {1, 3, 0},           // 16   = 2 x 8
        {2, 3, 0},           // 32   = 4 x 8
        {3, 3, 0},           // 64   = 8 x 8
        {2, 2, 3, 0},        // 128  = 4 x 4 x 8
        {2, 3, 3, 0},        // 256  = 4 x 8 x 8
        {3, 3, 3, 0},        // 512
        {2, 2, 3, 3, 0},     // 1024
        {2, 3, 3, 3, 0},     // 2048
        {3, 3, 3, 3, 0},     // 4096
        {2, 2, 3, 3, 3, 0},  // 8192
        {2, 3, 3, 3, 3, 0},  // 16384
        {3, 3, 3, 3, 3, 0}}; // 32768
    if (!is_pow2(n) || n < 8 || n > 32768) return false;
    if (dft_n < 0 || dft_n > n || dft_n > 64) return false;
    pl->n = n;
    pl->dft_n = dft_n;
    pl->log2n = ilog2(n);
    int ns = 0;
    for (int s = 0; s < 6; ++s) pl->log2r[s] = 0;
    while (table[pl->log2n][ns] != 0) { pl->log2r[ns] = table[pl->log2n][ns]; ++ns; }
    pl->nstages = ns;
    pl->pad_shift = pl->log2r[ns - 1];
    pl->plen = n + (n >> pl->pad_shift);
    // run_stage's constant-stride addressing needs every non-final stage stride to be a multiple of the padding period
    int lgL = pl->log2n;
    for (int s = 0; s + 1 < ns; ++s) {
        lgL -= pl->log2r[s];
        if (lgL < pl->pad_shift) return false;
    }
    return true;
}

// W_n^k = exp(-2 pi i k / n), k in [0, n), computed in long double and rounded once.
// A dense transform of length dft_n < n keeps the table size n (the kernels copy n entries to shared memory) and
// fills its first dft_n entries with W_dft_n^k.
template <typename T> inline void make_twiddles(int n, std::vector<cplx<T>>& tw, int dft_n = 0) {
    if (dft_n) {
        tw.assign(n, cplx<T>{(T)0, (T)0});
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int k = 0; k < dft_n; ++k) {
            const long double ang = two_pi * (long double)k / (long double)dft_n;
            tw[k].re = (T)cosl(ang);
            tw[k].im = (T)(-sinl(ang));
        }
        tw[0].re = 1; tw[0].im = 0;
        if (dft_n % 2 == 0) { tw[dft_n / 2].re = -1; tw[dft_n / 2].im = 0; }
        if (dft_n % 4 == 0) { tw[dft_n / 4].re = 0; tw[dft_n / 4].im = -1; tw[3 * dft_n / 4].re = 0; tw[3 * dft_n / 4].im = 1; }
        return;
    }
    tw.resize(n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; ++k) {
        // reduce to the first octant for accuracy
        int kk = k % n;
        long double ang = two_pi * (long double)kk / (long double)n;
        tw[k].re = (T)cosl(ang);
        tw[k].im = (T)(-sinl(ang));
    }
    // exact values on the axes
    tw[0].re = 1; tw[0].im = 0;
    if (n % 4 == 0) { tw[n / 4].re = 0; tw[n / 4].im = -1; tw[n / 2].re = -1; tw[n / 2].im = 0; tw[3 * n / 4].re = 0; tw[3 * n / 4].im = 1; }
}

// Geometry for a cluster of G CTAs and a shared-memory FFT workspace of at most ws_limit bytes.
// Returns false if the shape cannot be handled.
// dft_ny / dft_nx: 0, or the length of the dense DFT that replaces the radix transform along that axis (bsgp_fft.cuh).
inline bool make_geom(int ny, int nx, int G, size_t elem_bytes /* sizeof(cplx<T>) */, size_t ws_limit, ConvGeom* g,
                      size_t* ws_bytes, int dft_ny = 0, int dft_nx = 0) {
    if (!is_pow2(ny) || !is_pow2(nx) || ny < 16 || nx < 16) return false;
    if (!make_fft_plan(nx, &g->px, dft_nx) || !make_fft_plan(ny, &g->py, dft_ny)) return false;
    if ((dft_ny || dft_nx) && G != 1) return false;          // dense transforms: small images, one CTA each
    g->ny = ny; g->nx = nx; g->hx = nx / 2;
    g->lg_nx = ilog2(nx); g->lg_ny = ilog2(ny); g->lg_hx = g->lg_nx - 1;
    g->G = G;
    g->wrap_ny = g->wrap_nx = 0;
    if (ny % (2 * G) != 0 || (nx / 2) % G != 0) return false;
    g->rows_per_cta = ny / G;
    g->cols_per_cta = g->hx / G;
    g->rowstride = g->px.plen;
    g->colstride = g->py.plen | 1;                   // odd: conflict-free transposing loads
    int rtp = g->rows_per_cta / 2;
    while (rtp > 1 && (size_t)rtp * g->rowstride * elem_bytes > ws_limit) rtp >>= 1;
    int ct = g->cols_per_cta;
    while (ct > 1 && (size_t)ct * g->colstride * elem_bytes > ws_limit) ct >>= 1;
    if ((size_t)rtp * g->rowstride * elem_bytes > ws_limit || (size_t)ct * g->colstride * elem_bytes > ws_limit) return false;
    g->row_tile_pairs = rtp;
    g->col_tile = ct;
    g->lg_col_tile = ilog2(ct);
    g->lg_cp = ilog2(g->cols_per_cta);              // panel width of the exchange buffer in frame mode (bsgp_conv.cuh)
    const size_t a = (size_t)rtp * g->rowstride * elem_bytes, b = (size_t)ct * g->colstride * elem_bytes;
    *ws_bytes = a > b ? a : b;
    return true;
}

}  // namespace bsgp
