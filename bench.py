#!/usr/bin/env python
"""bench.py -- log_prob+grad evaluations/s of the ppcseq NB model on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3_60kx500] [--impl reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one log_prob + full gradient evaluation of one gene shard (the named workload) at one
of 8 theta points ~ U(-2,2)^D.  With N ranks every rank owns its own shard of an (N x G)-gene model
(weak scaling): a step is one evaluation of that model = local fused kernel + all-reduce(SUM) of 8
doubles + a tiny finalise kernel.  `value` = shard evaluations per second over all ranks.
Timing: CUDA events around every step on the launching stream, L2 flushed (256 MiB memset) between
steps outside the timed events, max over ranks.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "log_prob+grad evals/sec"
UNIT = "evals/s"


def _env_int(k, d):
    return int(os.environ.get(k, d))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(w, excl, seconds=12.0):
    """C oracle (map_rect-style threads) on a bounded sample of the same workload, host cores."""
    from oracle import c_oracle, model_np
    cores = os.cpu_count() or 1
    Gs = min(w.G, max(256, 2_000_000 // w.S))
    def run(G_sub, reps):
        d = model_np.ModelData(w.counts[:G_sub], w.X, w.exposure, min(w.K, G_sub),
                               exclude=None if excl is None else excl[:G_sub])
        th = np.random.default_rng(0).uniform(-2, 2, model_np.dim(G_sub, d.K, w.C))
        c_oracle.log_prob_grad(d, th, n_shards=cores)
        t0 = time.perf_counter()
        for _ in range(reps):
            c_oracle.log_prob_grad(d, th, n_shards=cores)
        return (time.perf_counter() - t0) / reps
    t = run(Gs, 1)
    per_elem = t / (Gs * w.S)
    # scale the sample so that the whole baseline takes about `seconds`
    G_sub = int(min(w.G, max(Gs, seconds / 4 / per_elem / w.S)))
    reps = max(1, int(seconds / max(per_elem * G_sub * w.S, 1e-9)))
    reps = min(reps, 50)
    t = run(G_sub, reps)
    evals_per_s = 1.0 / (t * w.G / G_sub)
    return {"value": evals_per_s, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} evals of the first {G_sub} of {w.G} genes x {w.S} samples, scaled to the full "
                      f"workload; C restatement of the Stan program, one thread per map_rect shard, no AD tape"}


def ppc_bench(w, device, n_post=1000, genes=4000, p=0.05):
    """Posterior-predictive draws/s (BASELINE metric, second half): exact analysis over n_post posterior draws of the
    first `genes` genes of the workload, through the C ABI (ppcseq_fit_from_draws + ppcseq_ppc_summary), host wall
    clock around the summary call (includes the device->host copy of the four [K,S] outputs)."""
    import ppcseq_b200
    from ppcseq_b200 import Fit
    Gp = min(genes, w.G)
    m = ppcseq_b200.NBModel(w.counts[:Gp], w.X, w.exposure, Gp, device=device)
    lay = m.layout
    rng = np.random.default_rng(5)
    full = ppcseq_b200.layout(w.G, w.K, w.C)
    th = np.zeros(lay.D)
    th[:3] = w.theta_true[:3]; th[-3:] = w.theta_true[-3:]
    th[lay.o_intercept:lay.o_intercept + Gp] = w.theta_true[full.o_intercept:full.o_intercept + Gp]
    th[lay.o_sigma_raw:lay.o_sigma_raw + Gp] = w.theta_true[full.o_sigma_raw:full.o_sigma_raw + Gp]
    if w.C >= 2:
        th[lay.o_alpha1:lay.o_alpha1 + Gp] = w.theta_true[full.o_alpha1:full.o_alpha1 + Gp]
    if w.C >= 3:
        th[lay.o_alpha2:lay.o_alpha2 + (w.C - 2) * Gp] = w.theta_true[full.o_alpha2:full.o_alpha2 + (w.C - 2) * Gp]
    draws = th[None, :] + 0.05 * rng.standard_normal((n_post, lay.D))
    fit = Fit.from_draws(m, draws)
    fit.ppc_summary(p, exact=True, seed=1)                       # warm-up
    l0 = ppcseq_b200.lib().ppcseq_launch_count()
    t0 = time.perf_counter()
    reps = 3
    for r in range(reps):
        fit.ppc_summary(p, exact=True, seed=2 + r)
    dt = (time.perf_counter() - t0) / reps
    launches = (ppcseq_b200.lib().ppcseq_launch_count() - l0) // reps
    n = float(n_post) * Gp * w.S
    fit.close(); m.close()
    return {"value": n / dt, "unit": "NB draws/s", "seconds_per_call": dt, "gpu_launches_per_call": int(launches),
            "config": {"genes": Gp, "samples": w.S, "posterior_draws": n_post, "p": p, "analysis": "exact (fit_to_counts_rng)",
                       "outputs": ".lower/.upper/mean/sd per (gene, sample)"}}


def identify_outliers_bench(device, G=515, K=15, S=21, seed=3):
    """identify_outliers wall-clock, both passes, on a synthetic table of the README configuration's shape
    (15 genes to check + 500 controls x 21 samples, ~Label, VB, approximate analysis in pass 2)."""
    import pandas as pd
    from ppcseq_b200 import synthetic
    from ppcseq_b200.api import identify_outliers
    w = synthetic.make(G=G, S=S, C=2, K=K, mask=False, seed=seed)
    df = pd.DataFrame({
        "symbol": np.repeat([f"g{i}" for i in range(G)], S), "sample": np.tile([f"s{j:03d}" for j in range(S)], G),
        "value": w.counts.reshape(-1), "Label": np.tile(np.where(w.X[:, 1] > 0, "B", "A"), G),
        "PValue": np.repeat(np.where(np.arange(G) < K, 1e-6, 0.9), S), "do_check": np.repeat(np.arange(G) < K, S)})
    out = {}
    for name, vb in (("vb", True), ("nuts", False)):
        t0 = time.perf_counter()
        res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                                significance="PValue", do_check="do_check", percent_false_positive_genes=5,
                                approximate_posterior_inference=vb, cores=4, seed=11, device=device)
        out[name] = {"wall_s": time.perf_counter() - t0, "genes_flagged": int((res["ppc_samples_failed"] > 0).sum()),
                     "fit2_evals": float(res.attrs["fit 2 info"][1])}
    out["config"] = {"G": G, "K": K, "S": S, "formula": "~ Label", "percent_false_positive_genes": 5,
                     "includes": "input preparation + TMM on the host, model upload, 2 inference passes, PPC, flags"}
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    from ppcseq_b200 import synthetic
    w = synthetic.make(args.workload)
    excl = None
    if len(w.exclude_pairs):
        excl = np.zeros((w.G, w.S), bool)
        excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
    from oracle import c_oracle, model_np
    cores = os.cpu_count() or 1
    # bounded sample per step: as many leading genes as ~2 s of CPU allow
    probe_G = min(w.G, max(256, 1_000_000 // w.S))
    d = model_np.ModelData(w.counts[:probe_G], w.X, w.exposure, min(w.K, probe_G),
                           exclude=None if excl is None else excl[:probe_G])
    th = np.random.default_rng(0).uniform(-2, 2, model_np.dim(probe_G, d.K, w.C))
    c_oracle.log_prob_grad(d, th, n_shards=cores)
    t0 = time.perf_counter(); c_oracle.log_prob_grad(d, th, n_shards=cores); tp = time.perf_counter() - t0
    G_sub = int(min(w.G, max(probe_G, 2.0 / (tp / probe_G))))
    d = model_np.ModelData(w.counts[:G_sub], w.X, w.exposure, min(w.K, G_sub),
                           exclude=None if excl is None else excl[:G_sub])
    ths = np.random.default_rng(1).uniform(-2, 2, (8, model_np.dim(G_sub, d.K, w.C)))
    for i in range(args.warmup):
        c_oracle.log_prob_grad(d, ths[i % 8], n_shards=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        c_oracle.log_prob_grad(d, ths[i % 8], n_shards=cores)
    T = time.perf_counter() - t0
    ms_step_full = T / args.steps * (w.G / G_sub) * 1e3
    value = 1e3 / ms_step_full
    sample = (f"each step = one evaluation of the first {G_sub} of {w.G} genes x {w.S} samples on {cores} host "
              f"threads (C restatement of the Stan program; rstan/Stan cannot be built here), scaled to the full workload")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step_full, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": args.workload, "G": w.G, "S": w.S, "C": w.C, "K": w.K,
                      "pass2_mask": bool(len(w.exclude_pairs))},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import ppcseq_b200
    from ppcseq_b200 import synthetic
    from ppcseq_b200._lib import check

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = ppcseq_b200.lib()

    cfg = dict(synthetic.CONFIGS[args.workload])
    w = synthetic.make(G=cfg["G"], S=cfg["S"], C=cfg["C"], mask=cfg["mask"], seed=cfg["seed"] + rank)
    shard = (w.G * world, rank * w.G) if world > 1 else None
    t0 = time.perf_counter()
    # K is "all genes checked": the shard constructor takes the global K
    model = ppcseq_b200.NBModel(w.counts, w.X, w.exposure, w.K * world if world > 1 else w.K,
                                device=local_rank, shard=shard)
    if args.no_mask:
        w.exclude_pairs = w.exclude_pairs[:0]
    if len(w.exclude_pairs):
        model.set_exclusion(w.exclude_pairs)
    model.set_design_path({"auto": 0, "general": 1, "element": 2, "moments": 3}[args.path])
    t_create = time.perf_counter() - t0
    D = model.D
    ths_host = synthetic.random_thetas(w, 8, seed=1)          # same hyper-parameters on every rank
    if world > 1:
        hyper = torch.from_numpy(np.concatenate([ths_host[:, :3], ths_host[:, -3:]], axis=1)).to(dev)
        dist.broadcast(hyper, 0)
        h = hyper.cpu().numpy()
        ths_host[:, :3] = h[:, :3]; ths_host[:, -3:] = h[:, 3:]
    ths = torch.from_numpy(ths_host).to(dev)
    lp = torch.zeros(1, dtype=torch.float64, device=dev)
    grad = torch.zeros(D, dtype=torch.float64, device=dev)
    partials = torch.zeros(8, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)             # a real (non-NULL) stream shared with the library
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    H = model.handle

    fused = world > 1 and args.collective == "fused"
    if fused:
        from ppcseq_b200 import dist as pdist
        pdist.connect(model, rank, world, channels=1, cap=1)      # all-reduce fused into the kernel (peer mailboxes)

    def step(i):
        th = ths[i % 8]
        if world == 1 or fused:
            check(L.ppcseq_log_prob_grad_device(H, 1, th.data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))
        else:
            check(L.ppcseq_log_prob_grad_partial_device(H, 1, th.data_ptr(), 1, partials.data_ptr(), grad.data_ptr(), sp))
            dist.all_reduce(partials)
            check(L.ppcseq_finalize_hyper_device(H, 1, th.data_ptr(), partials.data_ptr(), 1, 1, lp.data_ptr(),
                                                 grad.data_ptr(), sp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.ppcseq_launch_count()
    barrier()
    for i in range(args.steps):
        if not args.no_flush:
            flush.zero_()                               # L2 flush, outside the timed events
        ev[i][0].record(stream)
        step(i)
        ev[i][1].record(stream)
    barrier()
    launches = L.ppcseq_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    times = np.array([a.elapsed_time(b) for a, b in ev])          # ms
    total_ms = torch.tensor([times.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public host API: pinned host thetas in, lp + gradients out (pinned), every step ----
    # The call a user makes is model.log_prob_grad(thetas[B, D]) -> ppcseq_log_prob_grad (C ABI, host pointers).  One
    # call carries EB thetas (the chains of a sampler / the draws of an ELBO estimate); inside the library theta b+1
    # goes host->device and gradient b-1 device->host while evaluation b runs.  `single_call` is the same API with B = 1.
    EB = 8
    th_pin = torch.from_numpy(ths_host).pin_memory()
    g_pin = torch.empty((EB, D), dtype=torch.float64).pin_memory()
    lp_pin = torch.empty(EB, dtype=torch.float64).pin_memory()
    th_dev = torch.empty(D, dtype=torch.float64, device=dev)
    th_np, g_np, lp_np = th_pin.numpy(), g_pin.numpy(), lp_pin.numpy()

    def step_e2e_single(i):
        if world == 1 or fused:
            return model.log_prob_grad(th_np[i % 8], out=(lp_np[:1], g_np[:1]))
        th_dev.copy_(th_pin[i % 8], non_blocking=True)
        check(L.ppcseq_log_prob_grad_partial_device(H, 1, th_dev.data_ptr(), 1, partials.data_ptr(), grad.data_ptr(), sp))
        dist.all_reduce(partials)
        check(L.ppcseq_finalize_hyper_device(H, 1, th_dev.data_ptr(), partials.data_ptr(), 1, 1, lp.data_ptr(),
                                             grad.data_ptr(), sp))
        g_pin[0].copy_(grad, non_blocking=True); lp_pin[:1].copy_(lp, non_blocking=True)
        torch.cuda.synchronize()
        return float(lp_pin[0]), g_pin[0]

    def timed(fn, n_calls, evals_per_call):
        for i in range(3):
            fn(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_calls):
            fn(i)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n_calls * evals_per_call / float(t.item())

    e2e_single = timed(step_e2e_single, args.steps, 1)
    if world == 1 or fused:
        n_calls = max(1, (args.steps + EB - 1) // EB)
        e2e_value = timed(lambda i: model.log_prob_grad(th_np, out=(lp_np, g_np)), n_calls, EB)
        e2e_mode = f"{EB} thetas per call, 3-stage copy/compute pipeline inside ppcseq_log_prob_grad"
    else:
        e2e_value, e2e_mode = e2e_single, "one theta per call (nccl variant)"

    if rank == 0:
        peaks, which = measured_peaks()
        B_eval = w.algorithmic_bytes_per_eval()
        achieved = B_eval / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(args.workload)
        fp64 = ctypes.c_double()
        check(L.ppcseq_measure_fp64_peak(local_rank, ctypes.byref(fp64)))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "G": w.G, "S": w.S, "C": w.C, "K": w.K,
                       "pass2_mask": bool(len(w.exclude_pairs)), "D": int(D),
                       "thetas": "8 points ~ U(-2,2)^D, cycled", "l2": "NOT flushed (diagnostic run)" if args.no_flush else "flushed between steps (256 MiB memset)",
                       "per_rank": "each rank owns one such shard of an N x G gene model" if world > 1 else "single GPU",
                       "collective": ("fused in-kernel peer all-reduce (NVLink mailboxes)" if fused else
                                      "nccl all_reduce of 8 doubles + finalize kernel") if world > 1 else "none",
                       "likelihood_path": args.path},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": which,
                         "algorithmic_bytes_per_launch": B_eval,
                         "fp64_peak_tflops_measured": fp64.value,
                         "note": "achieved = SURVEY 8(d) algorithmic bytes (dense counts once + theta/grad + X) / kernel time; the "
                                 "kernel reads data-only sufficient statistics instead (`traffic` = measured DRAM bytes) and is "
                                 "bound by FP64 issue + dependent special-function chains, see DESIGN.md section 4"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * D),
                    "d2h_bytes_per_step": int(8 * D + 8), "mode": e2e_mode, "single_call": e2e_single,
                    "model_create_s": t_create, "model_create_h2d_bytes": int(w.counts.nbytes)},
            "gpu_launches": int(launches), "clocks": clocks,
            "step_ms": {"min": float(times.min()), "median": float(np.median(times)), "max": float(times.max())},
        }
        if world == 1 and not args.no_cpu_baseline:
            excl = None
            if len(w.exclude_pairs):
                excl = np.zeros((w.G, w.S), bool)
                excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
            out["cpu_baseline"] = cpu_baseline(w, excl)
        if world == 1 and not args.no_extras:
            try:
                out["ppc"] = ppc_bench(w, local_rank)
                out["identify_outliers"] = identify_outliers_bench(local_rank)
            except Exception as e:                      # the headline line must still be printed
                out["extras_error"] = repr(e)
        print(json.dumps(out), flush=True)
    if fused:
        from ppcseq_b200 import dist as pdist
        if pdist.comm_timed_out(model):
            raise RuntimeError("fused all-reduce timed out on rank %d (ranks out of step)" % rank)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3_60kx500")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mask", action="store_true", help="pass-1 variant of the workload (no exclusion list)")
    ap.add_argument("--no-extras", action="store_true", help="skip the PPC draws/s and identify_outliers wall-clock legs")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: leave L2 warm between steps (not a valid bench line)")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--path", default="auto", choices=["auto", "general", "element", "moments"],
                    help="likelihood path of the kernel (ppcseq_model_set_design_path)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    rank, world, local_rank = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1), _env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
