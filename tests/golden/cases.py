"""The parity cases shared by make_golden.py (runs the unmodified reference), the oracle pin test and
the GPU parity tests.  Each case: (dataset, divergence, kwargs, store_full_image).

datasets: "ngc" / "sat" = the two SGP-dec simulations bundled with the reference
(restoration/simulated_test/data/*.mat, loaded at simulation_test_sgp.py:18,38), "stamp<i>" /
"tile<i>" = seeded synthetic inputs stored in fixtures.
"""
import numpy as np

NGC_BETA0 = 0.9887296104546054      # simulation_test_sgp.py:98
SAT_BETA0 = 1.0001                  # simulation_test_sgp.py:154

_stamp_kw = dict(gamma=1e-4, beta=0.4, alpha_min=1e-5, alpha_max=1e5, alpha=1e1, M_alpha=3, tau=0.5, M=1,
                 proj_type=1, max_projs=1000, init_recon=2, stop_criterion=3, verbose=True,
                 ccd_sat_level=65000, scale_data=True, lr=1e-3, lr_exp_param=0.1, schedule_lr=True,
                 adapt_beta=True)
_tile_kw = dict(_stamp_kw, adapt_beta=False, tol_convergence=1e-5)

CASES = {
    # --- reference functional runs (simulation_test_sgp.py:25,45,100-104,156-160) ---
    "ngc_kl_27": ("ngc", "kl", dict(init_recon=3, stop_criterion=1, MAXIT=27), True),
    "sat_kl_40": ("sat", "kl", dict(init_recon=3, stop_criterion=1, MAXIT=40), True),
    "sat_kl_332": ("sat", "kl", dict(init_recon=3, stop_criterion=1, MAXIT=332), False),
    "ngc_beta_27": ("ngc", "beta", dict(init_recon=3, stop_criterion=1, MAXIT=27, betaParam=NGC_BETA0, lr=1e-3,
                                         lr_exp_param=0.1, schedule_lr=True, adapt_beta=False), False),
    "sat_beta_332": ("sat", "beta", dict(init_recon=3, stop_criterion=1, MAXIT=332, betaParam=SAT_BETA0, lr=1e-3,
                                          lr_exp_param=0.1, schedule_lr=True, adapt_beta=False), False),
    # --- BASELINE config 2 family: beta-SGP with the flux-conserving projection ---
    "ngc_beta_p1_27": ("ngc", "beta", dict(init_recon=3, proj_type=1, stop_criterion=1, MAXIT=27,
                                            betaParam=NGC_BETA0, adapt_beta=False), True),
    "ngc_beta_p1_stop3": ("ngc", "beta", dict(init_recon=3, proj_type=1, stop_criterion=3, MAXIT=332,
                                               betaParam=NGC_BETA0, adapt_beta=False), True),
    "sat_beta_p1_40": ("sat", "beta", dict(init_recon=3, proj_type=1, stop_criterion=1, MAXIT=40,
                                            betaParam=SAT_BETA0, adapt_beta=False), True),
    "sat_beta_p1_stop3": ("sat", "beta", dict(init_recon=3, proj_type=1, stop_criterion=3, MAXIT=332,
                                               betaParam=SAT_BETA0, adapt_beta=False), True),
    "sat_beta_p1_332": ("sat", "beta", dict(init_recon=3, proj_type=1, stop_criterion=1, MAXIT=332,
                                             betaParam=SAT_BETA0, adapt_beta=False), False),
    # --- option coverage (every init / projection / stop rule / flag of SURVEY.md §8a) ---
    "ngc_kl_p1_stop2": ("ngc", "kl", dict(init_recon=2, proj_type=1, stop_criterion=2, MAXIT=60,
                                           tol_convergence=2e-2), False),
    "ngc_kl_stop2_quiet": ("ngc", "kl", dict(init_recon=2, stop_criterion=2, MAXIT=60, verbose=False,
                                              tol_convergence=1e-3), False),
    "ngc_kl_stop4": ("ngc", "kl", dict(init_recon=3, stop_criterion=4, MAXIT=60), False),
    "ngc_kl_init0": ("ngc", "kl", dict(init_recon=0, stop_criterion=1, MAXIT=12), False),
    "ngc_kl_init1_p1": ("ngc", "kl", dict(init_recon=1, proj_type=1, stop_criterion=1, MAXIT=12), False),
    "ngc_kl_noscale_flux": ("ngc", "kl", dict(init_recon=3, proj_type=1, stop_criterion=3, MAXIT=40,
                                               scale_data=False, flux=np.float64(2325942.0), ccd_sat_level=60000.0), False),
    "ngc_kl_nonmonotone": ("ngc", "kl", dict(init_recon=2, stop_criterion=1, MAXIT=25, M=3, M_alpha=5,
                                              alpha=5.0), False),
    "sat_kl_default_stop": ("sat", "kl", dict(init_recon=3, MAXIT=10), False),
    "ngc_beta_adapt": ("ngc", "beta", dict(init_recon=2, proj_type=1, stop_criterion=3, MAXIT=200, alpha=1e1,
                                            betaParam=1.0248357, adapt_beta=True, schedule_lr=True,
                                            ccd_sat_level=65000), False),
    "ngc_beta_is": ("ngc", "beta", dict(init_recon=3, stop_criterion=1, MAXIT=10, betaParam=0, adapt_beta=False), False),
    "ngc_beta_one": ("ngc", "beta", dict(init_recon=3, stop_criterion=1, MAXIT=10, betaParam=1, adapt_beta=False), False),
    "ngc_beta_two": ("ngc", "beta", dict(init_recon=3, proj_type=1, stop_criterion=1, MAXIT=10, betaParam=2.0,
                                          adapt_beta=False), False),
}
# config 3 (stamps) and config 4 (tiles): per-input kwargs supplied by make_golden (flux, betaParam)
N_STAMPS = 16
N_TILES = 2
for _i in range(N_STAMPS):
    CASES[f"stamp{_i:02d}"] = (f"stamp{_i}", "beta", dict(_stamp_kw), True)
for _i in range(N_TILES * 5):
    CASES[f"tile{_i:02d}"] = (f"tile{_i}", "beta", dict(_tile_kw), _i in (0, 7))

# cases where the strict north_star tolerances are expected to hold (SURVEY.md §7 hard part 1):
# identical iteration counts, discr rel. diff <= 1e-10, image ||dx||inf/||x||inf <= 1e-8
STRICT = ["ngc_kl_27", "sat_kl_40", "ngc_beta_27", "ngc_beta_p1_27", "ngc_beta_p1_stop3", "sat_beta_p1_40",
          "ngc_kl_init0", "ngc_kl_init1_p1", "ngc_kl_nonmonotone", "sat_kl_default_stop", "ngc_beta_is",
          "ngc_beta_one", "ngc_beta_two", "ngc_kl_stop4", "ngc_kl_noscale_flux", "ngc_kl_p1_stop2",
          "ngc_kl_stop2_quiet"]
