#!/usr/bin/env python
"""Benchmark of the beta-SGP restoration path (BASELINE.json metric: restored images/s and ms/iteration,
fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tiles256|stamps32]

Workload (config.workload): BASELINE config 4 — a synthetic 2048x2048 crowded field split into 64
256x256 subdivisions, each solved for the 5 beta initialisations of the reference's application script
(application_sgp_subdivisions.py:69-107), i.e. 320 independent beta-SGP restorations with the
flux-conserving projection, a 2-D background map and one shared PSF per GPU.  One "step" = one pass of
the hot path over that batch (PSF spectrum + 320 solves).  Multi-GPU: independent units are sharded, no
data-path collective; weak scaling (every rank solves its own 320-solve field).

`--impl reference` times the reference's own CPU algorithm (the oracle port, oracle/sgp_oracle.py: numpy,
same FFT calls and operand order as restoration/sgp.py; the Python reference itself cannot travel to the
GPU box) on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # the CPU arm runs one single-threaded numpy process per core; stop BLAS / OpenMP from oversubscribing them
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(_v, "1")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiles256", choices=["tiles256", "stamps32", "frame"])
    ap.add_argument("--field", type=int, default=2048, help="side of the synthetic field (tiles256)")
    ap.add_argument("--stamps", type=int, default=8192, help="number of stamps (stamps32)")
    ap.add_argument("--frame", type=int, default=8192, help="side of the single frame (workload frame, BASELINE config 5)")
    ap.add_argument("--maxit", type=int, default=10, help="iterations of the frame workload (stop_criterion=1)")
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--cluster", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=8)
    return ap.parse_args()


def make_workload(args, rank):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bsgp_synth", os.path.join(ROOT, "beta-sgp_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    if args.workload == "tiles256":
        w = synth.field_tiles(size=args.field, tile=256, seed=2024 + rank, n_beta=5)
        kw = dict(synth.TILE_KWARGS)
        name = (f"config4: {args.field}x{args.field} synthetic crowded field -> {len(w['gn']) // 5} subdivisions of 256x256 "
                f"(reference tiler utils.py:332-375, overlap 0) x 5 beta inits = {len(w['gn'])} beta-SGP solves, "
                "proj_type=1, 2-D bkg, shared PSF")
        return w, kw, name, True
    if args.workload == "frame":
        f = synth.single_frame(args.frame, seed=77 + rank)
        w = dict(gn=f["gn"][None], bkg=f["bkg"][None], psf=f["psf"], flux=np.array([f["flux"]]), beta0=np.array([1.0248357076505616]))
        kw = dict(synth.TILE_KWARGS, stop_criterion=1, MAXIT=args.maxit)
        name = (f"config5: one synthetic {args.frame}x{args.frame} crowded frame, beta-SGP, proj_type=1, 2-D bkg, frame-sized PSF, "
                f"{args.maxit} iterations (stop_criterion=1), frame mode (one image over the whole GPU)")
        return w, kw, name, True
    w = synth.star_stamps(args.stamps, 32, seed=12345 + rank)
    kw = dict(synth.STAMP_KWARGS)
    name = f"config3: {args.stamps} synthetic 32x32 star stamps, per-stamp PSF, adapt_beta, proj_type=1"
    return w, kw, name, False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def _nvml_loop(self):
        import pynvml as nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.nvml_rows.append((float(sm), [n for n, b in bits.items() if r & b]))
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def start(self):
        # NVML in-process (a sample every 20 ms: a 5-step timed region of 0.4 s gets ~20); nvidia-smi -lms as the fallback
        self.nvml_rows, self.h = [], None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.stop_flag = threading.Event()
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if getattr(self, "h", None) is not None:
            self.stop_flag.set()
            self.t.join(1.0)
            sm = [r[0] for r in self.nvml_rows]
            reasons = sorted({n for r in self.nvml_rows for n in r[1]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(npix, wbytes, iters, evals, trials, shared_psf, bkg_image):
    """B_iter = [26 + 2E + 3(T-1)] N w; shared TF: -2 N w; full-image bkg: +(3 + (T-1)) N w.
    Summed over the iterations of every image with the counted E (projection evaluations) and T (trials)."""
    per_iter = 26.0 - (2.0 if shared_psf else 0.0) + (3.0 if bkg_image else 0.0)
    extra_trials = np.maximum(trials - iters, 0)
    units = per_iter * iters + 2.0 * evals + (4.0 if bkg_image else 3.0) * extra_trials
    return float(units.sum()) * npix * wbytes


def _oracle_solve(job):
    from oracle import sgp_oracle as orc
    gn, psf, bkg, flux, b0, kw = job
    r = orc.solve(gn, psf, bkg, divergence="beta", flux=np.float64(flux), betaParam=float(b0), **kw)
    return r.iters


def cpu_jobs(w, kw, shared_psf, idx):
    jobs = []
    for i in idx:
        psf = w["psf"] if shared_psf else w["psf"][i]
        bkg = w["bkg"][i] if np.ndim(w["bkg"][i]) == 2 else np.float64(w["bkg"][i])
        jobs.append((w["gn"][i], psf, bkg, float(w["flux"][i]), float(w["beta0"][i]), kw))
    return jobs


def frame_cpu_baseline(w, kw, side):
    """The oracle needs ~40 s per iteration at 8192^2; time a 2048^2 crop for 3 iterations and scale by the pixel
    ratio (the loop is O(N log N): the scaled figure flatters the CPU slightly)."""
    from oracle import sgp_oracle as orc
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bsgp_synth2", os.path.join(ROOT, "beta-sgp_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    c = min(2048, side)
    gn = np.ascontiguousarray(w["gn"][0][:c, :c]); bkg = np.ascontiguousarray(w["bkg"][0][:c, :c])
    psf = synth.moffat_psf(c, c, 3.5)
    k = dict(kw, MAXIT=3)
    t0 = time.perf_counter()
    orc.solve(gn, psf, bkg, divergence="beta", flux=np.float64((gn - bkg).sum()), betaParam=float(w["beta0"][0]), **k)
    per_iter = (time.perf_counter() - t0) / 3.0 * (side * side) / float(c * c)
    return {"value": 1.0 / (per_iter * kw["MAXIT"]), "unit": "images/s", "cores": 1, "kind": "port",
            "sample": f"{c}x{c} crop, 3 iterations of the oracle port on 1 thread, scaled by the pixel ratio to {side}x{side} and {kw['MAXIT']} iterations"}


def run_reference(args, rank):
    """CPU arm: oracle port on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    w, kw, name, shared = make_workload(args, 0)
    if args.workload == "frame":
        cb = frame_cpu_baseline(w, kw, w["gn"].shape[-1])
        print(json.dumps({"impl": "reference", "metric": "beta-SGP restored images/s", "value": cb["value"], "unit": "images/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": name}, "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    procs = max(1, min(cores, 64))
    n = len(w["gn"])
    per_step = procs if args.workload == "tiles256" else procs * 16
    rng = np.random.default_rng(0)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        def step():
            idx = rng.choice(n, per_step, replace=per_step > n)
            return sum(pool.map(_oracle_solve, cpu_jobs(w, kw, shared, idx), chunksize=1 if args.workload == "tiles256" else 16))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        iters = 0
        for _ in range(args.steps):
            iters += step()
        dt = time.perf_counter() - t0
    images = per_step * args.steps
    val = images / dt
    line = {"impl": "reference", "metric": "beta-SGP restored images/s", "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "sample_per_step": per_step},
            "ms_per_image_iteration": 1e3 * dt * procs / max(iters, 1) / procs,
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": procs, "kind": "port",
                             "sample": f"{per_step} random solves of the workload per step on {procs} processes (1 numpy thread each)"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import beta_sgp_b200 as bs

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w, kw, wname, shared_psf = make_workload(args, rank)
    tdt = torch.float64 if args.dtype == "float64" else torch.float32
    B, ny, nx = w["gn"].shape
    wbytes = 8 if args.dtype == "float64" else 4
    bkg_image = np.ndim(w["bkg"]) == 3

    # pinned host copies (e2e leg) and resident device copies (value leg)
    host = {k: torch.as_tensor(np.ascontiguousarray(w[k])).to(tdt if k in ("gn", "psf", "bkg") else torch.float64).pin_memory()
            for k in ("gn", "psf", "bkg", "flux", "beta0")}
    devt = {k: v.to(dev) for k, v in host.items()}
    plan = bs.Plan(ny, nx, args.dtype, local_rank, cluster_size=args.cluster, threads=args.threads)
    info = plan.info()
    ev_k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_k1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def step(src, k=None):
        plan.set_psf(src["psf"])
        if k is not None:
            ev_k0[k].record()
        r = bs.solve_batch(src["gn"], None, src["bkg"], divergence="beta", flux=src["flux"], betaParam=src["beta0"], plan=plan,
                           psf_is_set=True, **kw)
        if k is not None:
            ev_k1[k].record()
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---------------- value: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        res = step(devt)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        res = step(devt, k)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(ev_k0, ev_k1)]))
    iters = res.iters.cpu().numpy().astype(np.float64)
    evals = res.proj_evals.cpu().numpy().astype(np.float64)
    trials = res.ls_trials.cpu().numpy().astype(np.float64)
    status = res.status.cpu().numpy()
    assert (status == 0).all(), "solver reported a failure status"
    # share of (clusters in flight x kernel time) spent inside solves: 1 - this is queue tail + launch overhead
    t_img = res.times.cpu().numpy()[np.arange(B), res.iters.cpu().numpy().astype(int)]
    slot_util = float(t_img.sum() / (min(info["num_clusters"], B) * kernel_ms * 1e-3))
    x_sum = res.x.sum(dim=(1, 2)).cpu().numpy()
    assert np.abs(x_sum - w["flux"]).max() <= 1e-8 * np.abs(w["flux"]).max() or args.dtype == "float32", "flux not conserved"

    # ---------------- e2e: page-locked host buffers in, page-locked host results out, everything inside the timed region ----------------
    # The public call for host data: solve_batch with pinned CPU tensors -> bsgp_solve_batch_pinned (upload in queue order on a
    # copy stream behind per-item ready flags, restored images stored zero-copy into the pinned output, small outputs copied back).
    x_host = torch.empty((B, ny, nx), dtype=tdt).pin_memory()

    def step_e2e():
        plan.set_psf(host["psf"].to(dev, non_blocking=True))
        return bs.solve_batch(host["gn"], None, host["bkg"], divergence="beta", flux=host["flux"].numpy(), betaParam=host["beta0"].numpy(),
                              plan=plan, psf_is_set=True, device=local_rank, x_out=x_host, **kw)

    for _ in range(min(args.warmup, 3)):
        r_e2e = step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        r_e2e = step_e2e()
    f1.record()
    barrier()
    ms_e2e = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
    assert np.array_equal(r_e2e.iters, res.iters.cpu().numpy()) and torch.equal(r_e2e.x, res.x.cpu()), "host path and resident path disagree"
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        tot = torch.tensor([iters.sum()], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        total_iters = float(tot.item())
    else:
        total_iters = float(iters.sum())
    ms, ms_e2e = float(ms.item()), float(ms_e2e.item())
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = r_e2e.x.numel() * r_e2e.x.element_size() + sum(getattr(r_e2e, k).nbytes for k in ("iters", "status", "discr", "times", "stop_value", "beta_final", "proj_evals", "ls_trials", "scalars"))

    if rank == 0:
        images = B * world * args.steps
        value = images / (ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        abytes = algorithmic_bytes(ny * nx, wbytes, iters, evals, trials, shared_psf, bkg_image)
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(args.workload)
        line = {
            "metric": "beta-SGP restored images/s", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32", "data": "synthetic",
            "config": {"workload": wname, "images_per_gpu_per_step": B, "cluster_size": info["cluster_size"],
                       "clusters_in_flight": info["num_clusters"], "threads": info["threads"], "smem_bytes": info["smem_bytes"],
                       "l2": f"inputs {h2d / 1e6:.0f} MB per step > 126 MB L2, no flush needed; per-cluster scratch "
                             f"{info['workspace_bytes'] / 1e6:.0f} MB " + ("streams through L2/HBM" if info['workspace_bytes'] > 100e6 else "fits L2"),
                       "mean_iterations": float(iters.mean()), "mean_proj_evals_per_iter": float(evals.sum() / iters.sum()),
                       "mean_trials_per_iter": float(trials.sum() / iters.sum()), "max_iterations": int(iters.max()),
                       "cluster_slot_utilisation": slot_util},
            "ms_per_image_iteration": ms * 1e-3 * 1e3 / (total_iters * args.steps) if total_iters else None,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "bsgp_frame_kernel" if info["cluster_size"] > 16 else "bsgp_solve_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": abytes,
                         "peak_source": peak_src},
            "e2e": {"value": images / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 2 * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, ROOT)
            if args.workload == "frame":
                line["cpu_baseline"] = frame_cpu_baseline(w, kw, ny)
            else:
                n = min(args.cpu_sample, B) if args.workload == "tiles256" else min(args.cpu_sample * 64, B)
                idx = np.linspace(0, B - 1, n).astype(int)
                jobs = cpu_jobs(w, kw, shared_psf, idx)
                _oracle_solve(jobs[0])
                t0 = time.perf_counter()
                for j in jobs:
                    _oracle_solve(j)
                dt = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": n / dt, "unit": "images/s", "cores": 1, "kind": "port",
                                        "sample": f"{n} of the {B} solves (evenly spaced indices), oracle port, 1 thread, after one warm-up solve"}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
