"""Where does the time of the pipelined host path go?  usage: python tools/pin_probe.py [workload]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import beta_sgp_b200 as bs
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "tiles256"
class A: pass
a = A(); a.workload = wl; a.dtype = "float64"; a.field = 2048; a.stamps = 8192; a.frame = 8192; a.maxit = 10
w, kw, wname, shared = bench.make_workload(a, 0)
dev = torch.device("cuda:0")
host = {k: torch.as_tensor(np.ascontiguousarray(w[k])).to(torch.float64).pin_memory() for k in ("gn", "psf", "bkg", "flux", "beta0")}
B, ny, nx = host["gn"].shape
plan = bs.Plan(ny, nx, "float64", 0)
def run():
    plan.set_psf(host["psf"].to(dev, non_blocking=True))
    return bs.solve_batch(host["gn"], None, host["bkg"], divergence="beta", flux=host["flux"].numpy(), betaParam=host["beta0"].numpy(), plan=plan, psf_is_set=True, **kw)
for mode in (0, 1, 2, 3):
    os.environ["BSGP_PIN_MODE"] = str(mode)
    for _ in range(2): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = run(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); x = torch.empty((B, ny, nx), dtype=torch.float64, pin_memory=True); t_alloc = (time.perf_counter() - t0) * 1e3
    print(f"mode {mode}: wall ms {[round(t, 2) for t in ts]}  pinned alloc ms {t_alloc:.2f}", flush=True)
