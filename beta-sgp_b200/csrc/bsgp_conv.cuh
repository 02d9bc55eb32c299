// Circular PSF convolution of one real image, split over the G CTAs of a thread-block cluster.
//
// Replaces `afunction` (sgp.py:111-117 / 573-579): out = real(ifftn(TF * fftn(x))) with
// TF = fftn(fftshift(psf)) (sgp.py:109 / 571) or its conjugate for A^T.
//
// Decomposition (real-to-complex / complex-to-real, half spectrum, fused PSF multiply):
//   rows_forward : CTA `rank` owns image rows [rank*ny/G, (rank+1)*ny/G).  Two real rows are
//                  transformed by one complex FFT (row 2p -> real part, row 2p+1 -> imaginary
//                  part) and untangled into two half spectra.  Values come from a producer functor
//                  (so the elementwise step that creates the input is fused in) and the half
//                  spectra go to the cluster's global (L2-resident) exchange buffer
//                  spec[ny][hx], hx = nx/2; column 0 packs the two real columns kx = 0 and
//                  kx = nx/2 as (re, im).
//   cols         : CTA `rank` owns packed spectrum columns [rank*hx/G, (rank+1)*hx/G): forward
//                  column FFT, multiply by the precomputed PSF spectrum (or its conjugate),
//                  inverse column FFT, back to the exchange buffer.
//   rows_inverse : the mirror of rows_forward; real outputs are handed to a consumer functor (so
//                  the elementwise step that uses the result is fused in).
// A cluster barrier is required between the three phases (the caller issues it).
//
// Scaling: rows_forward stores 2*A[k]; the stored PSF spectrum carries 0.5/(nx*ny) (0.25/(nx*ny)
// for the two packed columns, whose untangle doubles once more), all powers of two, so the result
// equals numpy's normalised ifftn without a separate scaling pass.
#pragma once
#include "bsgp_fft.cuh"

namespace bsgp {

struct ConvGeom {
    int ny, nx, hx;          // hx = nx / 2 packed spectrum columns
    int lg_nx, lg_ny, lg_hx;
    int G;                   // CTAs per image (cluster size)
    int rows_per_cta;        // ny / G (even)
    int cols_per_cta;        // hx / G
    int row_tile_pairs;      // row pairs per workspace tile (divides rows_per_cta / 2), power of two
    int col_tile;            // columns per workspace tile (divides cols_per_cta), power of two
    int lg_col_tile;
    int rowstride;           // padded complex elements per row pair in the workspace
    int colstride;           // padded complex elements per column in the workspace (odd)
    FftPlan px, py;
};

enum ConvMode { CONV_TF = 0, CONV_CTF = 1, CONV_MAKE_TF = 2 };

// prod(local_row, col) -> T  for local_row in [0, rows_per_cta)
template <class Ctx, typename T, class Prod>
BSGP_DEV void conv_rows_forward(Ctx& ctx, const ConvGeom& g, cplx<T>* ws, const cplx<T>* twx, cplx<T>* spec, Prod& prod) {
    const int r0 = ctx.rank * g.rows_per_cta;
    const int ps = g.px.pad_shift;
    const int ntiles = (g.rows_per_cta >> 1) / g.row_tile_pairs;
    for (int tile = 0; tile < ntiles; ++tile) {
        const int pair0 = tile * g.row_tile_pairs;
        for (int e = ctx.tid; e < (g.row_tile_pairs << g.lg_nx); e += ctx.nt) {
            const int p = e >> g.lg_nx, c = e & (g.nx - 1);
            const int row = 2 * (pair0 + p);
            const T v0 = prod(row, c);
            const T v1 = prod(row + 1, c);
            ws[(size_t)p * g.rowstride + fpad(c, ps)] = cmake<T>(v0, v1);
        }
        ctx.sync();
        fft_batch<false>(ctx, ws, g.row_tile_pairs, g.rowstride, g.px, twx);
        for (int e = ctx.tid; e < (g.row_tile_pairs << g.lg_hx); e += ctx.nt) {
            const int p = e >> g.lg_hx, k = e & (g.hx - 1);
            const cplx<T>* a = ws + (size_t)p * g.rowstride;
            cplx<T> A, B;
            if (k == 0) {
                const cplx<T> z0 = a[fpad(pos_of_freq(g.px, 0), ps)], zh = a[fpad(pos_of_freq(g.px, g.hx), ps)];
                A = cmake<T>(z0.re + z0.re, zh.re + zh.re);
                B = cmake<T>(z0.im + z0.im, zh.im + zh.im);
            } else {
                const cplx<T> zk = a[fpad(pos_of_freq(g.px, k), ps)], zm = a[fpad(pos_of_freq(g.px, g.nx - k), ps)];
                A = cmake<T>(zk.re + zm.re, zk.im - zm.im);
                B = cmake<T>(zk.im + zm.im, zm.re - zk.re);
            }
            const size_t row = (size_t)(r0 + 2 * (pair0 + p));
            spec[row * g.hx + k] = A;
            spec[(row + 1) * g.hx + k] = B;
        }
        ctx.sync();
    }
}

// tf: [hx + 1][ny] complex in column-workspace position order (written by CONV_MAKE_TF);
// row 0 = kx 0, rows 1..hx-1 = kx, row hx = kx nx/2.
template <class Ctx, typename T>
BSGP_DEV void conv_cols(Ctx& ctx, const ConvGeom& g, cplx<T>* ws, const cplx<T>* twy, cplx<T>* spec, cplx<T>* tf, int mode) {
    const int c0 = ctx.rank * g.cols_per_cta;
    const int ps = g.py.pad_shift;
    const int ntiles = g.cols_per_cta / g.col_tile;
    const int ct = g.col_tile;
    for (int tile = 0; tile < ntiles; ++tile) {
        const int cc0 = c0 + tile * ct;
        for (int e = ctx.tid; e < (g.ny << g.lg_col_tile); e += ctx.nt) {
            const int row = e >> g.lg_col_tile, cl = e & (ct - 1);
            ws[(size_t)cl * g.colstride + fpad(row, ps)] = spec[(size_t)row * g.hx + cc0 + cl];
        }
        ctx.sync();
        fft_batch<false>(ctx, ws, ct, g.colstride, g.py, twy);
        if (mode == CONV_MAKE_TF) {
            // column FFT of 2*A holds 2*TF; store TF * 0.5/(nx ny)
            const T sc = (T)0.25 / ((T)g.nx * (T)g.ny);
            for (int e = ctx.tid; e < (ct << g.lg_ny); e += ctx.nt) {
                const int cl = e >> g.lg_ny, p = e & (g.ny - 1);
                if (cc0 + cl == 0) continue;
                tf[(size_t)(cc0 + cl) * g.ny + p] = cscale(ws[(size_t)cl * g.colstride + fpad(p, ps)], sc);
            }
            if (cc0 == 0) {
                const T sp = (T)0.0625 / ((T)g.nx * (T)g.ny);   // (C0 = 4 TF0) * 0.25/(nx ny)
                for (int ky = ctx.tid; ky <= (g.ny >> 1); ky += ctx.nt) {
                    const int pa = pos_of_freq(g.py, ky), pb = pos_of_freq(g.py, (g.ny - ky) & (g.ny - 1));
                    const cplx<T> za = ws[fpad(pa, ps)], zb = ws[fpad(pb, ps)];
                    const cplx<T> c0v = cmake<T>(za.re + zb.re, za.im - zb.im);          // za + conj(zb)
                    const cplx<T> chv = cmake<T>(za.im + zb.im, zb.re - za.re);          // -i (za - conj(zb))
                    tf[pa] = cscale(c0v, sp);
                    tf[(size_t)g.hx * g.ny + pa] = cscale(chv, sp);
                    tf[pb] = cscale(cconj(c0v), sp);
                    tf[(size_t)g.hx * g.ny + pb] = cscale(cconj(chv), sp);
                }
            }
            ctx.sync();
            continue;
        }
        for (int e = ctx.tid; e < (ct << g.lg_ny); e += ctx.nt) {
            const int cl = e >> g.lg_ny, p = e & (g.ny - 1);
            if (cc0 + cl == 0) continue;
            cplx<T>* z = ws + (size_t)cl * g.colstride + fpad(p, ps);
            const cplx<T> t = tf[(size_t)(cc0 + cl) * g.ny + p];
            *z = (mode == CONV_TF) ? cmul(*z, t) : cmulc(*z, t);
        }
        if (cc0 == 0) {
            for (int ky = ctx.tid; ky <= (g.ny >> 1); ky += ctx.nt) {
                const int pa = pos_of_freq(g.py, ky), pb = pos_of_freq(g.py, (g.ny - ky) & (g.ny - 1));
                const cplx<T> za = ws[fpad(pa, ps)], zb = ws[fpad(pb, ps)];
                const cplx<T> c0a = cmake<T>(za.re + zb.re, za.im - zb.im);
                const cplx<T> cha = cmake<T>(za.im + zb.im, zb.re - za.re);
                cplx<T> t0a = tf[pa], tha = tf[(size_t)g.hx * g.ny + pa], t0b = tf[pb], thb = tf[(size_t)g.hx * g.ny + pb];
                if (mode == CONV_CTF) { t0a = cconj(t0a); tha = cconj(tha); t0b = cconj(t0b); thb = cconj(thb); }
                const cplx<T> ya0 = cmul(t0a, c0a), yah = cmul(tha, cha);
                const cplx<T> yb0 = cmul(t0b, cconj(c0a)), ybh = cmul(thb, cconj(cha));
                // Y = Y0 + i Yh
                ws[fpad(pa, ps)] = cmake<T>(ya0.re - yah.im, ya0.im + yah.re);
                if (pb != pa) ws[fpad(pb, ps)] = cmake<T>(yb0.re - ybh.im, yb0.im + ybh.re);
            }
        }
        ctx.sync();
        fft_batch<true>(ctx, ws, ct, g.colstride, g.py, twy);
        for (int e = ctx.tid; e < (g.ny << g.lg_col_tile); e += ctx.nt) {
            const int row = e >> g.lg_col_tile, cl = e & (ct - 1);
            spec[(size_t)row * g.hx + cc0 + cl] = ws[(size_t)cl * g.colstride + fpad(row, ps)];
        }
        ctx.sync();
    }
}

// cons(local_row, col, value)
template <class Ctx, typename T, class Cons>
BSGP_DEV void conv_rows_inverse(Ctx& ctx, const ConvGeom& g, cplx<T>* ws, const cplx<T>* twx, const cplx<T>* spec, Cons& cons) {
    const int r0 = ctx.rank * g.rows_per_cta;
    const int ps = g.px.pad_shift;
    const int ntiles = (g.rows_per_cta >> 1) / g.row_tile_pairs;
    for (int tile = 0; tile < ntiles; ++tile) {
        const int pair0 = tile * g.row_tile_pairs;
        for (int e = ctx.tid; e < (g.row_tile_pairs << g.lg_hx); e += ctx.nt) {
            const int p = e >> g.lg_hx, k = e & (g.hx - 1);
            cplx<T>* a = ws + (size_t)p * g.rowstride;
            const size_t row = (size_t)(r0 + 2 * (pair0 + p));
            const cplx<T> A = spec[row * g.hx + k], B = spec[(row + 1) * g.hx + k];
            if (k == 0) {
                a[fpad(pos_of_freq(g.px, 0), ps)] = cmake<T>(A.re, B.re);
                a[fpad(pos_of_freq(g.px, g.hx), ps)] = cmake<T>(A.im, B.im);
            } else {
                a[fpad(pos_of_freq(g.px, k), ps)] = cmake<T>(A.re - B.im, A.im + B.re);
                a[fpad(pos_of_freq(g.px, g.nx - k), ps)] = cmake<T>(A.re + B.im, B.re - A.im);
            }
        }
        ctx.sync();
        fft_batch<true>(ctx, ws, g.row_tile_pairs, g.rowstride, g.px, twx);
        for (int e = ctx.tid; e < (g.row_tile_pairs << g.lg_nx); e += ctx.nt) {
            const int p = e >> g.lg_nx, c = e & (g.nx - 1);
            const cplx<T> z = ws[(size_t)p * g.rowstride + fpad(c, ps)];
            const int row = 2 * (pair0 + p);
            cons(row, c, z.re);
            cons(row + 1, c, z.im);
        }
        ctx.sync();
    }
}

// Whole convolution; `cluster_sync` separates the phases.  The leading block barrier orders the
// caller's earlier slab writes (made with a different pixel->thread mapping) before the producer
// reads them; between two convolutions no cluster barrier is needed because a CTA only rewrites
// the exchange-buffer rows it alone read in the previous rows_inverse.
template <class Ctx, typename T, class Prod, class Cons>
BSGP_DEV void conv_image(Ctx& ctx, const ConvGeom& g, cplx<T>* ws, const cplx<T>* twx, const cplx<T>* twy, cplx<T>* spec,
                         cplx<T>* tf, int mode, Prod& prod, Cons& cons) {
    ctx.sync();
    conv_rows_forward(ctx, g, ws, twx, spec, prod);
    ctx.cluster_sync();
    conv_cols(ctx, g, ws, twy, spec, tf, mode);
    ctx.cluster_sync();
    if (mode != CONV_MAKE_TF) conv_rows_inverse(ctx, g, ws, twx, spec, cons);
}

}  // namespace bsgp
