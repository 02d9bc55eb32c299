// Root-find on the multiplier of the flux-conserving scaled projection
//     min 1/2 x' diag(dia) x - c' x   s.t.  sum(x) = b,  0 <= x (<= sat)
// (flux_conserve_proj.py:7-144, the Dai-Fletcher style bracketing + safeguarded secant that the
// reference inherited from SGP-dec).  Every thread of the cluster runs this scalar state machine
// redundantly on identical all-reduced residuals, so control flow stays uniform without any
// broadcast; `eval(lam)` is the only collective step: it returns r(lam) = sum_i x_i(lam) - b.
//
// Kept on purpose: no warm start (lambda = 0, dlambda = 1 on every call, :7); the secant ratio is
// NOT refreshed in the `r > 0, s > 2` branch (the reference assigns it to `x` at :122); the
// floating-point-exception escape of the downward bracketing (:68-72).
// Added: the reference's bracketing loops are unbounded (they spin forever when the target cannot
// be reached, e.g. flux above N*sat); here they stop after `kBracketCap` growth steps or on a
// non-finite multiplier and report PROJ_NO_BRACKET instead of hanging the GPU.
#pragma once
#include "bsgp_math.cuh"

namespace bsgp {

enum ProjStatus { PROJ_OK = 0, PROJ_NO_BRACKET = 1 };
constexpr int kBracketCap = 4096;

struct ProjResult {
    double lambda;   // multiplier of the returned point
    int evals;       // full evaluations of r(lambda)
    int status;
};

template <class Eval>
BSGP_DEV ProjResult flux_rootfind(Eval& eval, double b, int max_projs, double lambda = 0.0, double dlambda = 1.0,
                                  double tol_lam = 1e-11, int biter0 = 0, int siter0 = 0) {
    ProjResult out;
    out.evals = 0;
    out.status = PROJ_OK;
    const double tol_r = 1e-11 * b;
    // The caller's counters (:7) only matter through the secant budget: biter keeps growing during the bracketing
    // (:39, :64), maxit_s = max_projs - biter is taken after it (:103) and the loop runs while siter < maxit_s (:106).
    int biter = biter0, siter = siter0, nsteps = 0;
    double lam = lambda, dlam = dlambda, lam_lo, lam_hi, r_lo, r_hi, s;

    double r = eval(lam); ++out.evals;                                  // :22-25
    if (fabs(r) < tol_r) { out.lambda = lam; return out; }              // :27-28

    if (r < 0) {                                                        // :30-54
        lam_lo = lam; r_lo = r;
        lam = lam + dlam;
        r = eval(lam); ++out.evals;
        while (r < 0) {
            ++biter;
            lam_lo = lam;
            s = np_max2(r_lo / r - 1.0, 0.1);
            dlam = dlam + dlam / s;
            lam = lam + dlam;
            r_lo = r;
            if (++nsteps > kBracketCap || !is_finite(lam)) { out.status = PROJ_NO_BRACKET; out.lambda = lam_lo; return out; }
            r = eval(lam); ++out.evals;
        }
        lam_hi = lam; r_hi = r;
    } else {                                                            // :55-81
        lam_hi = lam; r_hi = r;
        lam = lam - dlam;
        r = eval(lam); ++out.evals;
        while (r > 0) {
            ++biter;
            lam_hi = lam;
            s = np_max2(r_hi / r - 1.0, 0.1);
            const double grown = dlam + dlam / s;
            if (!is_finite(grown)) break;                               // :68-72 FP-exception escape
            dlam = grown;
            lam = lam - dlam;
            r_hi = r;
            if (++nsteps > kBracketCap) { out.status = PROJ_NO_BRACKET; out.lambda = lam_hi; return out; }
            r = eval(lam); ++out.evals;
        }
        lam_lo = lam; r_lo = r;
    }

    if (fabs(r_hi) < tol_r) { out.lambda = lam_hi; return out; }        // :84-93
    if (fabs(r_lo) < tol_r) { out.lambda = lam_lo; return out; }

    s = 1.0 - r_lo / r_hi;                                              // :96-103
    dlam = dlam / s;
    lam = lam_hi - dlam;
    r = eval(lam); ++out.evals;
    const int budget = max_projs - biter;

    while (fabs(r) > tol_r && dlam > nmul(tol_lam, 1.0 + fabs(lam)) && siter < budget) {   // :106-142
        ++siter;
        if (r > 0) {
            if (s <= 2.0) {
                lam_hi = lam; r_hi = r;
                s = 1.0 - r_lo / r_hi;
                dlam = (lam_hi - lam_lo) / s;
                lam = lam_hi - dlam;
            } else {
                s = np_max2(r_hi / r - 1.0, 0.1);
                dlam = (lam_hi - lam) / s;
                const double lam_new = np_max2(lam - dlam, nadd(nmul(0.75, lam_lo), nmul(0.25, lam)));
                lam_hi = lam; r_hi = r;
                lam = lam_new;
                // :122 writes (lam_hi - lam_lo) / (lam_hi - lam) into `x`; `s` keeps its value
            }
        } else {
            if (s >= 2.0) {
                lam_lo = lam; r_lo = r;
                s = 1.0 - r_lo / r_hi;
                dlam = (lam_hi - lam_lo) / s;
                lam = lam_hi - dlam;
            } else {
                s = np_max2(r_lo / r - 1.0, 0.1);
                dlam = (lam - lam_lo) / s;
                const double lam_new = np_min2(lam + dlam, nadd(nmul(0.75, lam_hi), nmul(0.25, lam)));
                lam_lo = lam; r_lo = r;
                lam = lam_new;
                s = (lam_hi - lam_lo) / (lam_hi - lam);
            }
        }
        r = eval(lam); ++out.evals;
    }
    out.lambda = lam;
    return out;
}

}  // namespace bsgp
