// One complete SGP / beta-SGP restoration of one image by one thread-block cluster.
//
// This is the body of the persistent solve kernel: every loop of the reference (outer iteration
// sgp.py:302-425 / 748-882, line search :328-349 / :776-800, projection root-find
// flux_conserve_proj.py:38-142) runs here with ordinary control flow.  All threads of all CTAs of
// the cluster execute the scalar controller redundantly on bit-identical all-reduced sums, so the
// control flow is uniform across the cluster and no scalar is ever broadcast or sent to the host.
//
// `Ctx` supplies: tid, nt, rank, G, sync(), cluster_sync(), allreduce_sum(double*, k),
// allreduce_min(double&), allreduce_max(double&), now().  DeviceCtx (bsgp_kernels.cu) is the
// product; tests/host_emul provides a one-thread emulation of the same code.
//
// Numerics follow the reference's operand order with non-contracted arithmetic (bsgp_math.cuh);
// deliberate deviations, all at rounding level:
//   * the projection evaluates (c + lambda) * X instead of (c + lambda) / D with D = 1/X,
//   * den^beta = den^(beta-1) * den and den^(beta-2) = den^(beta-1) / den (one pow per pixel),
//   * reductions are tree sums in fp64 (numpy: pairwise / BLAS dot).
#pragma once
#include "../../include/bsgp.h"
#include "bsgp_conv.cuh"
#include "bsgp_project.cuh"

namespace bsgp {

enum Buf { B_GN = 0, B_BKG, B_XA, B_XB, B_G, B_XTF, B_D, B_DTF, B_T1, NBUF };
constexpr int kMaxMem = 16;

template <typename T> struct SolveArgs {
    bsgp_params p;
    ConvGeom g;
    int batch;
    // inputs
    const T* gn; const T* bkg; int bkg_is_image; const double* flux; const double* beta0; const T* x0; const T* obj;
    // tables
    const cplx<T>* twx; const cplx<T>* twy; cplx<T>* tf; int n_psf;
    // per-cluster global scratch
    T* work; size_t work_stride; cplx<T>* spec; size_t spec_stride;
    int resident_mask;              // bit b: Buf b lives in shared memory
    // outputs
    T* x_out; int* iters; int* status; double* discr; double* times; double* stop_value; double* err;
    double* beta_final; int* proj_evals; int* ls_trials; double* scalars;
    double* tr_alpha; double* tr_lambda; double* tr_beta; int* tr_trials; int* tr_evals;
    int* queue;
};

// divergence constants for the current beta (sgp.py:452-458)
template <typename T> struct DivK {
    int kind;            // 0 KL (sgp), 1 beta generic, 2 beta == 0, 3 beta == 1
    T b, bm1, k, k2, k3;
};
template <typename T> BSGP_DEV DivK<T> make_divk(int divergence, double beta) {
    DivK<T> d;
    d.b = (T)beta; d.bm1 = (T)(beta - 1.0); d.k = d.k2 = d.k3 = (T)0;
    if (divergence == BSGP_DIV_KL) { d.kind = 0; return d; }
    if (beta == 0.0) { d.kind = 2; return d; }
    if (beta == 1.0) { d.kind = 3; return d; }
    d.kind = 1;
    const double k = 1.0 / nmul(beta, beta - 1.0);
    d.k = (T)k; d.k2 = (T)nmul(k, beta - 1.0); d.k3 = (T)nmul(k, beta);
    return d;
}

// One pixel of the objective at den = x_tf_try + bkg.  Adds its terms to acc[0..3] and returns the
// quantity cached for the gradient: gn/den (KL) or den^(beta-1) (beta-divergence).
//   KL    : acc0 += gn*log(gn/den), acc1 += x_tf_try                               sgp.py:333-334
//   beta  : acc0 += k*gn^b (only if want_s1), acc1 += k(b-1)*den^b, acc2 += k*b*gn*den^(b-1)   sgp.py:457-458
//   b == 0: acc0 += gn/den, acc1 += log(gn/den)                                     sgp.py:453
//   b == 1: acc0 += gn*log(gn/den), acc1 += gn, acc2 += den                         sgp.py:455
template <typename T> BSGP_DEV T objective_pixel(const DivK<T>& dk, T gnv, T den, T xtf_try, bool want_s1, KSum* acc) {
    if (dk.kind == 0) {
        const T ratio = ndiv(gnv, den);
        acc[0].add((double)nmul(gnv, mlog(ratio)));
        acc[1].add((double)xtf_try);
        return ratio;
    }
    if (dk.kind == 1) {
        const T p1 = mpow(den, dk.bm1);
        acc[1].add((double)nmul(dk.k2, nmul(p1, den)));
        acc[2].add((double)nmul(nmul(dk.k3, gnv), p1));
        if (want_s1) acc[0].add((double)nmul(dk.k, mpow(gnv, dk.b)));
        return p1;
    }
    const T ratio = ndiv(gnv, den);
    if (dk.kind == 2) {
        acc[0].add((double)ratio);
        acc[1].add((double)mlog(ratio));
        return ndiv((T)1, den);
    }
    acc[0].add((double)nmul(gnv, mlog(ratio)));
    acc[1].add((double)gnv);
    acc[2].add((double)den);
    return (T)1;
}

template <typename T> BSGP_DEV double objective_value(const DivK<T>& dk, const double* acc, double s1, double flux, double npix) {
    if (dk.kind == 0) return (acc[0] + acc[1]) - flux;
    if (dk.kind == 1) return (s1 + acc[1]) - acc[2];
    if (dk.kind == 2) return (acc[0] - acc[1]) - npix;
    return (acc[0] - acc[1]) + acc[2];
}

// per-pixel d D_beta / d beta, sgp.py:495 (term order kept)
template <typename T> BSGP_DEV T dbeta_pixel(T x, T y, T b) {
    const T bm1 = b - (T)1;
    const T ypb1 = mpow(y, bm1), ypb = nmul(ypb1, y), xpb = mpow(x, b), ly = mlog(y), lx = mlog(x);
    const T bm1sq = nmul(bm1, bm1), bbm1 = nmul(b, bm1), bsq = nmul(b, b);
    T s = ndiv(nmul(nmul(-x, ypb1), ly), bm1);
    s = nadd(s, ndiv(nmul(x, ypb1), bm1sq));
    s = nadd(s, ndiv(nmul(xpb, lx), bbm1));
    s = nsub(s, ndiv(xpb, nmul(b, bm1sq)));
    s = nadd(s, ndiv(nmul(ypb, ly), b));
    s = nsub(s, ndiv(xpb, nmul(bsq, bm1)));
    s = nsub(s, ndiv(ypb, bsq));
    return s;
}

template <typename T> BSGP_DEV T clip_bounds(T v, T lo, T hi) {   // sgp.py:355-357
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

template <typename T, class Ctx>
BSGP_DEV void solve_image(Ctx& ctx, const SolveArgs<T>& a, T* const* buf, cplx<T>* ws, cplx<T>* spec, cplx<T>* tf, int img) {
    const bsgp_params& P = a.p;
    const ConvGeom& g = a.g;
    const int nx = g.nx, lg_nx = g.lg_nx;
    const int nslab = g.rows_per_cta * nx;
    const size_t npix = (size_t)g.ny * nx;
    const double npix_d = (double)npix;
    const size_t goff = (size_t)img * npix + (size_t)ctx.rank * nslab;
    const size_t toff = (size_t)img * (P.maxit + 1);
    const bool leader = (ctx.rank == 0 && ctx.tid == 0);
    const bool pflag = P.proj_type == 1;
    const bool is_beta = P.divergence == BSGP_DIV_BETA;

    T* gn = buf[B_GN]; T* bkgb = buf[B_BKG]; T* xcur = buf[B_XA]; T* xnext = buf[B_XB]; T* gr = buf[B_G];
    T* xtf = buf[B_XTF]; T* dbuf = buf[B_D]; T* dtf = buf[B_DTF]; T* t1 = buf[B_T1];
    const T* gn_raw = a.gn + goff;
    const bool bkg_img = a.bkg_is_image != 0;
    const T* bkg_raw_img = bkg_img ? a.bkg + goff : nullptr;
    const T bkg_raw_s = bkg_img ? (T)0 : a.bkg[img];
    const double t_start = ctx.now();

    // ------------------------------------------------------------------ setup (sgp.py:166-217)
    double v3[3];
    double mx = -INFINITY;
    v3[0] = v3[1] = 0.0;
    for (int i = ctx.tid; i < nslab; i += ctx.nt) {
        const T v = gn_raw[i];
        mx = ((double)v > mx) ? (double)v : mx;
        v3[0] += (double)v;
        v3[1] += (double)nsub(v, bkg_img ? bkg_raw_img[i] : bkg_raw_s);
    }
    ctx.allreduce_sum(v3, 2);
    ctx.allreduce_max(mx);
    const double sum_raw = v3[0], sum_gb_raw = v3[1];
    const T scaling = P.scale_data ? (T)mx : (T)1;
    const T bkg_s = ndiv(bkg_raw_s, scaling);

    double vmin = INFINITY;
    for (int i = ctx.tid; i < nslab; i += ctx.nt) {
        const T v = ndiv(gn_raw[i], scaling);
        gn[i] = v;
        if (v > (T)0 && (double)v < vmin) vmin = (double)v;
    }
    ctx.allreduce_min(vmin);
    const T eps = Eps<T>::v();
    const T null_fill = nmul(nmul((T)vmin, eps), eps);
    const double flux_in = P.has_flux ? a.flux[img] : 0.0;
    const T x_const = ndiv(nmul((T)ndiv(P.has_flux ? flux_in : sum_gb_raw, npix_d), (T)1), scaling);
    v3[0] = 0.0;
    for (int i = ctx.tid; i < nslab; i += ctx.nt) {
        T v = gn[i];
        if (v <= (T)0) { v = null_fill; gn[i] = v; }
        T bk = bkg_s;
        if (bkg_img) { bk = ndiv(bkg_raw_img[i], scaling); bkgb[i] = bk; }
        v3[0] += (double)nsub(v, bk);
        T xv;
        if (P.init_recon == 0) xv = (T)0;
        else if (P.init_recon == 1) xv = ndiv(a.x0[goff + i], scaling);
        else if (P.init_recon == 2) xv = ndiv(gn_raw[i], scaling);
        else xv = x_const;
        xcur[i] = xv;
    }
    ctx.allreduce_sum(v3, 1);
    const double flux = P.has_flux ? ndiv(flux_in, (double)scaling) : v3[0];
    auto bkgv = [&](int i) -> T { return bkg_img ? bkgb[i] : bkg_s; };

    double tol = 0.0;                                                   // sgp.py:185-190, 291-294
    if (P.stop_criterion == 2 || P.stop_criterion == 3) tol = P.tol_convergence;
    else if (P.stop_criterion == 4) tol = 1.0 + 1.0 / (sum_raw / npix_d);
    if (P.verbose && P.stop_criterion == 2) tol = nmul(tol, tol);
    const double discr_coeff = nmul(2.0 / npix_d, (double)scaling);
    const bool has_cap = P.has_sat != 0;
    const T cap = has_cap ? nsub(ndiv((T)P.ccd_sat_level, scaling), eps) : (T)0;

    int status = BSGP_ST_OK;
    int total_evals = 0, total_trials = 0;
    if (pflag && !(flux > 0.0 && is_finite(flux))) status = BSGP_ST_BAD_FLUX;

    // x(lambda) of the projection for slab pixel i: min(cap, max(0, (c + lambda) * X))
    auto proj_point = [&](T c, T X, T lam) -> T {
        T v = nmul(nadd(c, lam), X);
        v = (v <= (T)0) ? (T)0 : v;
        if (has_cap) v = (v >= cap) ? cap : v;
        return v;
    };
    T* cbuf = dbuf;   // c = y * D during the root-find, then d in place
    T* Xbuf = t1;     // scaling-matrix diagonal during the root-find, objective cache afterwards
    auto proj_eval = [&](double lam) -> double {
        double s = 0.0;
        const T l = (T)lam;
        for (int i = ctx.tid; i < nslab; i += ctx.nt) s += (double)proj_point(cbuf[i], Xbuf[i], l);
        ctx.allreduce_sum(&s, 1);
        return s - flux;
    };

    // ------------------------------------------------------------------ initial projection (:248-253)
    if (status == BSGP_ST_OK) {
        if (!pflag) {
            for (int i = ctx.tid; i < nslab; i += ctx.nt) { const T v = xcur[i]; xcur[i] = (v < (T)0) ? (T)0 : v; }
        } else {
            for (int i = ctx.tid; i < nslab; i += ctx.nt) { cbuf[i] = xcur[i]; Xbuf[i] = (T)1; }
            const ProjResult pr = flux_rootfind(proj_eval, flux, P.max_projs);
            total_evals += pr.evals;
            if (pr.status != PROJ_OK) status = BSGP_ST_PROJ_NO_BRACKET;
            const T l = (T)pr.lambda;
            for (int i = ctx.tid; i < nslab; i += ctx.nt) xcur[i] = proj_point(cbuf[i], (T)1, l);
        }
    }

    double truth_sq = 1.0;
    const bool want_err = P.errflag && a.obj != nullptr && a.err != nullptr;
    const T* truth = want_err ? a.obj + goff : nullptr;
    if (want_err && status == BSGP_ST_OK) {                             // :240-244, 255-257
        double e2[2] = {0.0, 0.0};
        for (int i = ctx.tid; i < nslab; i += ctx.nt) {
            const T t = ndiv(truth[i], scaling);
            const T e = nsub(xcur[i], t);
            e2[0] += (double)nmul(e, e);
            e2[1] += (double)nmul(t, t);
        }
        ctx.allreduce_sum(e2, 2);
        truth_sq = e2[1];
        if (leader) a.err[(size_t)img * (P.maxit + 2)] = sqrt(e2[0] / truth_sq);
    }

    double beta_p = is_beta ? a.beta0[img] : 1.0;
    DivK<T> dk = make_divk<T>(P.divergence, beta_p);
    double s1 = 0.0;          // sum k*gn^beta for the current beta
    bool s1_valid = false;
    double fv = 0.0, x_low = 0.0, x_upp = 0.0;
    KSum osum[3];
    double acc[4];

    if (status == BSGP_ST_OK) {
        // ---------------------------------------------------------------- x_tf = A(x), objective (:260-265)
        osum[0].clear(); osum[1].clear(); osum[2].clear();
        {
            auto prod = [&](int row, int c) -> T { return xcur[(row << lg_nx) + c]; };
            auto cons = [&](int row, int c, T v) {
                const int i = (row << lg_nx) + c;
                xtf[i] = v;
                const T den = nadd(v, bkgv(i));
                t1[i] = objective_pixel(dk, gn[i], den, v, !s1_valid, osum);
            };
            conv_image(ctx, g, ws, a.twx, a.twy, spec, tf, CONV_TF, prod, cons);
        }
        acc[0] = osum[0].value(); acc[1] = osum[1].value(); acc[2] = osum[2].value();
        ctx.allreduce_sum(acc, 3);
        if (dk.kind == 1) { s1 = acc[0]; s1_valid = true; }
        fv = objective_value(dk, acc, s1, flux, npix_d);
        // ---------------------------------------------------------------- gradient
        {
            auto prod = [&](int row, int c) -> T {
                const int i = (row << lg_nx) + c;
                if (dk.kind == 0) return t1[i];
                const T den = nadd(xtf[i], bkgv(i));
                return nmul(gn[i], ndiv(t1[i], den));
            };
            auto cons = [&](int row, int c, T w) {
                const int i = (row << lg_nx) + c;
                gr[i] = (dk.kind == 0) ? nsub((T)1, w) : nsub(t1[i], w);
            };
            conv_image(ctx, g, ws, a.twx, a.twy, spec, tf, CONV_CTF, prod, cons);
        }
        // ---------------------------------------------------------------- scaling-matrix bounds (:268-273)
        double lo = INFINITY, hi = -INFINITY;
        {
            auto prod = [&](int row, int c) -> T { return gn[(row << lg_nx) + c]; };
            auto cons = [&](int row, int c, T w) {
                const int i = (row << lg_nx) + c;
                const T ratio = (T)ndiv(flux, nadd(flux, (double)bkgv(i)));
                const double yv = (double)nmul(ratio, w);
                if (yv > 0.0 && yv < lo) lo = yv;
                if (yv > hi) hi = yv;
            };
            conv_image(ctx, g, ws, a.twx, a.twy, spec, tf, CONV_CTF, prod, cons);
        }
        ctx.allreduce_min(lo);
        ctx.allreduce_max(hi);
        if (!(lo < INFINITY)) status = BSGP_ST_EMPTY_BOUNDS;
        x_low = lo; x_upp = hi;
        if (x_upp / x_low < 50.0) { x_low = x_low / 10.0; x_upp = x_upp * 10.0; }
    }

    if (leader) {
        a.discr[toff] = nmul(discr_coeff, fv);
        a.times[toff] = 0.0;
        if (a.stop_value) a.stop_value[toff] = 0.0;
    }

    // ------------------------------------------------------------------ main loop (sgp.py:302-425)
    double alpha = P.alpha, tau = P.tau, lr = P.lr;
    const int M = P.m, MA = P.m_alpha;
    double alpha_hist[kMaxMem], f_hist[kMaxMem];      // Valpha / Fold (sgp.py:214-215); dynamically indexed -> local memory
    for (int k = 0; k < MA; ++k) alpha_hist[k] = P.alpha_max;
    for (int k = 0; k < M; ++k) f_hist[k] = -1e30;
    const T xlo = (T)x_low, xhi = (T)x_upp;
    bool X_is_ones = (P.init_recon == 0);
    int iter = 1;
    bool keep_going = (status == BSGP_ST_OK);
    T* x_final = xcur;

    while (keep_going) {
        // history shift (:306-308)
        for (int k = 0; k < MA - 1; ++k) alpha_hist[k] = alpha_hist[k + 1];
        for (int k = 0; k < M - 1; ++k) f_hist[k] = f_hist[k + 1];
        f_hist[M - 1] = fv;
        const double f_prev = fv;

        // ---- trial point y = x - alpha X g, projection (:311-318)
        const T al = (T)alpha;
        int evals = 0;
        T lam_proj = (T)0;
        if (pflag) {
            for (int i = ctx.tid; i < nslab; i += ctx.nt) {
                const T xv = xcur[i];
                const T X = X_is_ones ? (T)1 : clip_bounds(xv, xlo, xhi);
                const T y = nsub(xv, nmul(al, nmul(X, gr[i])));
                cbuf[i] = nmul(y, ndiv((T)1, X));
                Xbuf[i] = X;
            }
            const ProjResult pr = flux_rootfind(proj_eval, flux, P.max_projs);
            evals = pr.evals;
            total_evals += evals;
            if (pr.status != PROJ_OK) { status = BSGP_ST_PROJ_NO_BRACKET; break; }
            lam_proj = (T)pr.lambda;
        }

        // ---- d = y - x, gd = d.g, d_tf = A(d) with the first line-search trial fused (:318-334)
        double sums[4];   // [0..2] objective terms, [3] gd
        osum[0].clear(); osum[1].clear(); osum[2].clear();
        double gd_part = 0.0;
        double lam = 1.0;
        {
            auto prod = [&](int row, int c) -> T {
                const int i = (row << lg_nx) + c;
                const T xv = xcur[i], gv = gr[i];
                T y;
                if (pflag) {
                    y = proj_point(cbuf[i], Xbuf[i], lam_proj);
                } else {
                    const T X = X_is_ones ? (T)1 : clip_bounds(xv, xlo, xhi);
                    y = nsub(xv, nmul(al, nmul(X, gv)));
                    y = (y < (T)0) ? (T)0 : y;
                }
                const T d = nsub(y, xv);
                dbuf[i] = d;
                gd_part += (double)nmul(d, gv);
                return d;
            };
            auto cons = [&](int row, int c, T v) {
                const int i = (row << lg_nx) + c;
                dtf[i] = v;
                const T xt = nadd(xtf[i], v);                 // lam = 1
                const T den = nadd(xt, bkgv(i));
                t1[i] = objective_pixel(dk, gn[i], den, xt, !s1_valid, osum);
            };
            conv_image(ctx, g, ws, a.twx, a.twy, spec, tf, CONV_TF, prod, cons);
        }
        sums[0] = osum[0].value(); sums[1] = osum[1].value(); sums[2] = osum[2].value(); sums[3] = gd_part;
        ctx.allreduce_sum(sums, 4);
        const double gd = sums[3];
        if (dk.kind == 1 && !s1_valid) { s1 = sums[0]; s1_valid = true; }
        fv = objective_value(dk, sums, s1, flux, npix_d);
        double f_ref = f_hist[0];                                        // fr = max(Fold)
        for (int k = 1; k < M; ++k) f_ref = py_max(f_ref, f_hist[k]);
        int trials = 1;
        // ---- backtracking (:328-349 / :776-800): accept iff fv <= fr + gamma*lam*gd or lam < 1e-12
        while (!(fv <= nadd(f_ref, nmul(nmul(P.gamma, lam), gd)) || lam < 1e-12)) {
            if (is_beta && P.adapt_beta && dk.kind == 1) {               // :798-800, den of the rejected trial
                double db = 0.0;
                const T l = (T)lam;
                for (int i = ctx.tid; i < nslab; i += ctx.nt) {
                    const T den = nadd(nadd(xtf[i], nmul(l, dtf[i])), bkgv(i));
                    db += (double)dbeta_pixel(gn[i], den, dk.b);
                }
                ctx.allreduce_sum(&db, 1);
                beta_p = nsub(beta_p, nmul(lr, db / npix_d));
                dk = make_divk<T>(P.divergence, beta_p);
                s1_valid = false;
            }
            lam = nmul(lam, P.ls_beta);
            ++trials;
            osum[0].clear(); osum[1].clear(); osum[2].clear();
            const T l = (T)lam;
            for (int i = ctx.tid; i < nslab; i += ctx.nt) {
                const T xt = nadd(xtf[i], nmul(l, dtf[i]));
                const T den = nadd(xt, bkgv(i));
                t1[i] = objective_pixel(dk, gn[i], den, xt, !s1_valid, osum);
            }
            sums[0] = osum[0].value(); sums[1] = osum[1].value(); sums[2] = osum[2].value();
            ctx.allreduce_sum(sums, 3);
            if (dk.kind == 1 && !s1_valid) { s1 = sums[0]; s1_valid = true; }
            fv = objective_value(dk, sums, s1, flux, npix_d);
        }
        total_trials += trials;

        // ---- accept: x, x_tf, new gradient through A^T, BB sums (:337-347, 355-365, 402)
        double bb[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) bb[k] = 0.0;
        {
            const T l = (T)lam;
            auto prod = [&](int row, int c) -> T {
                const int i = (row << lg_nx) + c;
                xnext[i] = nadd(xcur[i], nmul(l, dbuf[i]));
                const T xt = nadd(xtf[i], nmul(l, dtf[i]));
                xtf[i] = xt;
                if (dk.kind == 0) return t1[i];
                const T den = nadd(xt, bkgv(i));
                return nmul(gn[i], ndiv(t1[i], den));
            };
            auto cons = [&](int row, int c, T w) {
                const int i = (row << lg_nx) + c;
                const T gnew = (dk.kind == 0) ? nsub((T)1, w) : nsub(t1[i], w);
                const T yk = nsub(gnew, gr[i]);
                gr[i] = gnew;
                const T sk = nmul(l, dbuf[i]);
                const T xn = xnext[i];
                const T X = clip_bounds(xn, xlo, xhi);
                const T sk2 = nmul(sk, ndiv((T)1, X));
                const T yk2 = nmul(yk, X);
                bb[0] += (double)sk2 * (double)yk;
                bb[1] += (double)yk2 * (double)sk;
                bb[2] += (double)sk2 * (double)sk2;
                bb[3] += (double)yk2 * (double)yk2;
                if (P.stop_criterion == 2) { bb[4] += (double)sk * (double)sk; bb[5] += (double)xn * (double)xn; }
                if (want_err) { const T e = nsub(xn, ndiv(truth[i], scaling)); bb[6] += (double)nmul(e, e); }
            };
            conv_image(ctx, g, ws, a.twx, a.twy, spec, tf, CONV_CTF, prod, cons);
        }
        ctx.allreduce_sum(bb, 7);
        X_is_ones = false;

        // ---- Barzilai-Borwein steps and their alternation (:366-386)
        const double bk = bb[0], ck = bb[1];
        double a1, a2;
        if (bk <= 0.0) a1 = py_min(nmul(10.0, alpha), P.alpha_max);
        else a1 = py_min(P.alpha_max, py_max(P.alpha_min, bb[2] / bk));
        if (ck <= 0.0) a2 = py_min(nmul(10.0, alpha), P.alpha_max);
        else a2 = py_min(P.alpha_max, py_max(P.alpha_min, ck / bb[3]));
        alpha_hist[MA - 1] = a2;
        double amin = alpha_hist[0];
        for (int k = 1; k < MA; ++k) amin = py_min(amin, alpha_hist[k]);
        if (iter <= 20) alpha = amin;
        else if (a2 / a1 < tau) { alpha = amin; tau = nmul(tau, 0.9); }
        else { alpha = a1; tau = nmul(tau, 1.1); }

        if (is_beta && P.schedule_lr) lr = nmul(P.lr, exp(nmul(-P.lr_exp_param, (double)iter)));   // :842-844, epoch == iter

        // ---- bookkeeping and stop rules (:390-425)
        ++iter;
        double stop_val = 0.0;
        if (P.stop_criterion == 2) {
            stop_val = bb[4] / bb[5];
            keep_going = stop_val > tol;
        } else if (P.stop_criterion == 3) {
            stop_val = (f_prev - fv) / fv;
            keep_going = (stop_val > tol) && (stop_val >= 0.0);
        } else if (P.stop_criterion == 4) {
            stop_val = nmul(discr_coeff, fv);
            keep_going = stop_val > tol;
        }
        if (iter > P.maxit) keep_going = false;
        if (leader) {
            const size_t o = toff + (size_t)(iter - 1);
            a.times[o] = ctx.now() - t_start;
            a.discr[o] = nmul(discr_coeff, fv);
            if (a.stop_value) a.stop_value[o] = stop_val;
            if (a.tr_alpha) a.tr_alpha[o] = alpha;
            if (a.tr_lambda) a.tr_lambda[o] = lam;
            if (a.tr_beta) a.tr_beta[o] = beta_p;
            if (a.tr_trials) a.tr_trials[o] = trials;
            if (a.tr_evals) a.tr_evals[o] = evals;
            if (want_err && iter <= P.maxit + 1) a.err[(size_t)img * (P.maxit + 2) + iter] = sqrt(bb[6] / truth_sq);   // :394-396
        }
        if (keep_going) { T* t = xcur; xcur = xnext; xnext = t; }       // else: the previous iterate is returned (:424-425)
    }
    x_final = xcur;

    // ------------------------------------------------------------------ epilogue (:427-438)
    for (int i = ctx.tid; i < nslab; i += ctx.nt) a.x_out[goff + i] = nmul(x_final[i], scaling);
    if (leader) {
        a.iters[img] = iter - 1;
        a.status[img] = status;
        if (a.beta_final) a.beta_final[img] = beta_p;
        if (a.proj_evals) a.proj_evals[img] = total_evals;
        if (a.ls_trials) a.ls_trials[img] = total_trials;
        if (a.scalars) {
            double* sc = a.scalars + (size_t)img * BSGP_NSCALARS;
            sc[0] = (double)scaling; sc[1] = flux; sc[2] = x_low; sc[3] = x_upp; sc[4] = tol; sc[5] = fv; sc[6] = alpha; sc[7] = tau;
        }
    }
}

}  // namespace bsgp
