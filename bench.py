#!/usr/bin/env python
"""bench.py -- log_prob+grad evaluations/s of the ppcseq NB model on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3_60kx500] [--impl reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
              --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one log_prob + full gradient evaluation of the named workload at one of 9 theta points (8 ~ U(-2,2)^D
plus the generating truth).  With N ranks (--scaling strong, the default) the genes of that ONE fixed problem are split
into N contiguous blocks, one per GPU (BASELINE config 3: "60k x 500 on 1/2/4/8 B200"); a step is one fused kernel per
rank whose grid reduction ends in the in-kernel all-reduce of the 8 partial sums over NVLink mailboxes.  `value` =
evaluations of the whole problem per second.  Before anything is timed at N > 1, the `parity` block checks the
cross-rank sums (bitwise identical on all ranks, equal to the un-fused partial + all-gather + finalize formulation and to
the unsharded single-GPU evaluation) and the run exits non-zero if it fails.  `--scaling weak` (every rank its own
full-size shard) is round 1's curve; a strong run also reports it under `weak`.
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
    """Posterior-predictive draws/s (BASELINE metric, second half) through the C ABI (ppcseq_fit_from_draws +
    ppcseq_ppc_summary), host wall clock around the summary call (includes the device->host copy of the four [K,S]
    outputs).  `value`: the approximate analysis of pass 2 (fit_to_counts_rng_approximated: 50,000 NB draws per pair
    resampled from n_post posterior draws -- the mode that carries >99 % of the draws of a run at S >= 500);
    `exact`: fit_to_counts_rng over the n_post saved draws (pass 1)."""
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

    calls = []

    def rate(exact, nd, pp, genes_used, reps):
        fit.ppc_summary(pp, exact=exact, n_draws=nd, truncation_compensation=1.0 if exact else 0.7352941, seed=1)   # warm-up, full size
        l0 = ppcseq_b200.lib().ppcseq_launch_count()
        ts = []
        for r in range(reps):
            t0 = time.perf_counter()
            fit.ppc_summary(pp, exact=exact, n_draws=nd, truncation_compensation=1.0 if exact else 0.7352941, seed=2 + r)
            ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts))                       # per-call wall times are kept in `calls_s`
        calls.append([round(t, 4) for t in ts])
        launches = (ppcseq_b200.lib().ppcseq_launch_count() - l0) // reps
        return (float(n_post) if exact else float(nd)) * genes_used * w.S / dt, dt, int(launches)

    ex_rate, ex_dt, ex_l = rate(True, 0, p, Gp, 3)
    fit.close(); m.close()
    # approximate analysis on fewer genes (50,000 draws per pair: 400 genes x S pairs = 1e10 draws per call at S = 500)
    Ga = min(400, Gp)
    m = ppcseq_b200.NBModel(w.counts[:Ga], w.X, w.exposure, Ga, device=device)
    la = m.layout
    tha = np.concatenate([th[:3], th[lay.o_intercept:lay.o_intercept + Ga], th[lay.o_alpha1:lay.o_alpha1 + Ga] if w.C >= 2 else th[:0],
                          th[lay.o_alpha2:lay.o_alpha2 + (w.C - 2) * Ga] if w.C >= 3 else th[:0],
                          th[lay.o_sigma_raw:lay.o_sigma_raw + Ga], th[-3:]])
    assert tha.shape == (la.D,)
    fit = Fit.from_draws(m, tha[None, :] + 0.05 * rng.standard_normal((n_post, la.D)))
    ap_rate, ap_dt, ap_l = rate(False, 50000, 2e-4, Ga, 3)
    fit.close(); m.close()
    prof = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        prof = json.load(open(tp))
    return {"value": ap_rate, "unit": "NB draws/s", "seconds_per_call": ap_dt, "gpu_launches_per_call": ap_l,
            "dtype": "gamma / Poisson-rate arithmetic fp32, eta and the small-rate Poisson cdf fp64, sums exact integers",
            "config": {"genes": Ga, "samples": w.S, "posterior_draws": n_post, "nb_draws_per_pair": 50000, "p": 2e-4,
                       "analysis": "approximate (fit_to_counts_rng_approximated), truncation_compensation 0.7352941",
                       "outputs": ".lower/.upper/mean/sd per (gene, sample)"},
            "exact": {"value": ex_rate, "unit": "NB draws/s", "seconds_per_call": ex_dt, "gpu_launches_per_call": ex_l,
                      "config": {"genes": Gp, "samples": w.S, "posterior_draws": n_post, "p": p,
                                 "analysis": "exact (fit_to_counts_rng)"}},
            "calls_s": {"exact": calls[0], "approximate": calls[1]},
            "ncu": prof.get("ppc"),
            "bound": "ALU / RNG issue (memory traffic is 32 B per pair): see `ncu` (issue-slot and lane utilisation of the "
                     "last committed capture, profiles/)"}


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


def bench_config(args, w, masked):
    """The `config` object, identical in both arms (the driver compares them)."""
    from ppcseq_b200.model import Layout
    return {"workload": args.workload, "G": w.G, "S": w.S, "C": w.C, "K": w.K, "pass2_mask": bool(masked),
            "D": int(Layout(w.G, w.K, w.C).D), "thetas": "8 points ~ U(-2,2)^D + the generating truth, cycled",
            "l2": "NOT flushed (diagnostic run)" if args.no_flush else "flushed between steps (256 MiB memset)"}


def input_edge_bench(w):
    """The host side in front of the path at the bench workload's size: the tidy table of the
    workload (G x S rows, rows shuffled within genes so that nothing is in layout order) -> gene selection, G / S
    index, dense counts (ppcseq_prep_table) and TMM (ppcseq_tmm_factors), native host code of the library."""
    from ppcseq_b200 import prep
    G, S = w.G, w.S
    rng = np.random.default_rng(4)
    perm = rng.permutation(S)
    sym = np.repeat(np.arange(G, dtype=np.int64), S)
    sam = np.tile(perm.astype(np.int64), G)
    val = np.ascontiguousarray(w.counts[:, perm]).reshape(-1)
    sig = np.repeat(rng.uniform(0, 1, G), S)
    chk = np.repeat(rng.random(G) < 0.5, S)
    t0 = time.perf_counter()
    counts, genes, samples, K, first_row = prep.prepare_table(sam, sym, val, sig, chk, G)
    t1 = time.perf_counter()
    f, tot, ref = prep.tmm_factors(counts)
    t2 = time.perf_counter()
    ok = bool(np.array_equal(np.sort(counts, axis=1), np.sort(w.counts[np.asarray(genes)], axis=1)) and K == int(chk[::S].sum()))
    return {"rows": int(G * S), "table_to_dense_s": t1 - t0, "tmm_s": t2 - t1, "rows_per_s": G * S / (t1 - t0),
            "host_threads": os.cpu_count(), "consistent": ok,
            "note": "host only (no kernel): ppcseq_prep_table + ppcseq_tmm_factors, csrc/prep_host.cu"}


def identify_outliers_cfg2_bench(device):
    """identify_outliers(), both passes, at BASELINE config 2 (20,000 genes x 21 samples, ~Label, every gene checked):
    wall clock split into prep / upload / pass 1 / pass 2 / result, VB and NUTS (profiles/tools/e2e_identify_outliers.py
    runs the same at configs 3-5; those take minutes and are committed under profiles/)."""
    import pandas as pd
    from ppcseq_b200 import synthetic
    from ppcseq_b200.api import identify_outliers
    w = synthetic.make("cfg2_20kx21")
    G, S = w.G, w.S
    df = pd.DataFrame({"symbol": np.repeat(np.arange(G, dtype=np.int64), S), "sample": np.tile(np.arange(S, dtype=np.int64), G),
                       "value": w.counts.reshape(-1), "PValue": np.repeat(np.linspace(1e-9, 1e-3, G), S),
                       "do_check": np.ones(G * S, bool), "Label": np.tile(np.where(w.X[:, 1] > 0, "B", "A"), G)})
    out = {}
    for name, vb in (("vb", True), ("nuts", False)):
        tm = {}
        t0 = time.perf_counter()
        res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value", significance="PValue",
                                do_check="do_check", how_many_negative_controls=0, approximate_posterior_inference=vb,
                                cores=4, seed=11, device=device, return_format="failing", timings=tm)
        i1, i2 = tm.pop("pass1_info"), tm.pop("pass2_info")
        out[name] = {"wall_s": time.perf_counter() - t0, "split_s": tm, "evals": [float(i1[1]), float(i2[1])],
                     "sampler_s": [float(i1[2]), float(i2[2])], "failing_rows": int(len(res)),
                     "pass2_ppc_draws": float(res.attrs["total_draws"])}
    out["config"] = {"workload": "cfg2_20kx21", "G": G, "S": S, "K": G, "formula": "~ Label", "percent_false_positive_genes": 1}
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
    from ppcseq_b200 import dist as pdist
    ths = np.vstack([np.random.default_rng(1).uniform(-2, 2, (8, model_np.dim(G_sub, d.K, w.C))),
                     pdist.local_theta(w.theta_true, w.G, w.K, w.C, 0, G_sub)[None, :]])
    flush = np.zeros(256 << 20, np.uint8)               # the same cold-cache treatment as the GPU arm
    for i in range(args.warmup):
        c_oracle.log_prob_grad(d, ths[i % 9], n_shards=cores)
    T = 0.0
    for i in range(args.steps):
        flush[:] = i & 0xff
        t0 = time.perf_counter()
        c_oracle.log_prob_grad(d, ths[i % 9], n_shards=cores)
        T += time.perf_counter() - t0
    ms_step_full = T / args.steps * (w.G / G_sub) * 1e3
    value = 1e3 / ms_step_full
    sample = (f"each step = one evaluation of the first {G_sub} of {w.G} genes x {w.S} samples on {cores} host "
              f"threads (C restatement of the Stan program; rstan/Stan cannot be built here), scaled to the full workload")
    import ppcseq_b200
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step_full, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": bench_config(args, w, bool(len(w.exclude_pairs))),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def _rank_problem(args, rank, world):
    """The rank's gene shard.  strong: genes [g0, g1) of the ONE named workload (BASELINE config: a fixed problem on
    1/2/4/8 GPUs).  weak: every rank owns its own copy-sized shard of an (N x G)-gene model (round-1 behaviour)."""
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200 import synthetic
    cfg = dict(synthetic.CONFIGS[args.workload])
    if args.scaling == "strong" or world == 1:
        w = synthetic.make(args.workload)
        g0, g1 = pdist.shard_range(w.G, rank, world)
        pairs = w.exclude_pairs
        if args.no_mask:
            pairs = pairs[:0]
        sel = (pairs[:, 0] >= g0) & (pairs[:, 0] < g1)
        lp = pairs[sel].copy()
        lp[:, 0] -= g0
        return dict(w=w, counts=w.counts[g0:g1], K_total=w.K, G_total=w.G, g0=g0, g1=g1, pairs=lp, all_pairs=pairs)
    w = synthetic.make(G=cfg["G"], S=cfg["S"], C=cfg["C"], mask=cfg["mask"], seed=cfg["seed"] + rank)
    pairs = w.exclude_pairs[:0] if args.no_mask else w.exclude_pairs
    return dict(w=w, counts=w.counts, K_total=w.K * world, G_total=w.G * world, g0=rank * w.G, g1=(rank + 1) * w.G,
                pairs=pairs, all_pairs=None)


def _theta_points(pr, args, world, n_random=8):
    """8 points ~ U(-2,2)^D plus the generating truth, as LOCAL vectors of this rank's shard."""
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200 import synthetic
    w = pr["w"]
    if args.scaling == "strong" or world == 1:
        glob = np.vstack([synthetic.random_thetas(w, n_random, seed=1), w.theta_true])
        loc = np.stack([pdist.local_theta(t, w.G, w.K, w.C, pr["g0"], pr["g1"]) for t in glob])
        return glob, np.ascontiguousarray(loc)
    loc = np.vstack([synthetic.random_thetas(w, n_random, seed=1), w.theta_true])
    return None, np.ascontiguousarray(loc)          # hyper-parameters are made identical across ranks by the caller


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import ppcseq_b200
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200._lib import check

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = ppcseq_b200.lib()
    strong = args.scaling == "strong" or world == 1

    pr = _rank_problem(args, rank, world)
    w = pr["w"]
    t0 = time.perf_counter()
    model = ppcseq_b200.NBModel(pr["counts"], w.X, w.exposure, pr["K_total"], device=local_rank,
                                shard=(pr["G_total"], pr["g0"]) if world > 1 else None)
    if len(pr["pairs"]):
        model.set_exclusion(pr["pairs"])
    model.set_design_path({"auto": 0, "general": 1, "element": 2, "moments": 3}[args.path])
    t_create = time.perf_counter() - t0
    D = model.D
    ths_glob, ths_host = _theta_points(pr, args, world)
    NT = ths_host.shape[0]
    if world > 1 and not strong:                    # weak: same hyper-parameters on every rank
        hyper = torch.from_numpy(np.concatenate([ths_host[:, :3], ths_host[:, -3:]], axis=1)).to(dev)
        dist.broadcast(hyper, 0)
        h = hyper.cpu().numpy()
        ths_host[:, :3] = h[:, :3]; ths_host[:, -3:] = h[:, 3:]
    ths = torch.from_numpy(ths_host).to(dev)
    BB = max(1, args.batch)
    lp = torch.zeros(BB, dtype=torch.float64, device=dev)
    grad = torch.zeros((BB, D), dtype=torch.float64, device=dev)
    partials = torch.zeros((BB, 8), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(256 << 18, dtype=torch.float32, device=dev) if args.flush == "write+read" else None
    stream = torch.cuda.Stream(device=dev)             # a real (non-NULL) stream shared with the library
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    H = model.handle

    fused = world > 1 and args.collective == "fused"
    if fused:
        pdist.connect(model, rank, world, channels=1, cap=max(BB, 4))   # all-reduce fused into the kernel (peer mailboxes)

    def eval_partial_path(th_ptr, B):
        """the un-fused formulation: partial sums -> all-gather -> rank-ordered sum -> finalize kernel"""
        check(L.ppcseq_log_prob_grad_partial_device(H, B, th_ptr, 1, partials.data_ptr(), grad.data_ptr(), sp))
        if world > 1:
            parts = [torch.empty_like(partials[:B]) for _ in range(world)]
            dist.all_gather(parts, partials[:B].contiguous())
            tot = torch.zeros_like(parts[0])
            for q in range(world):                  # the kernel's order: ((0 + p_0) + p_1) + ...
                tot = tot + parts[q]
            partials[:B].copy_(tot)
        check(L.ppcseq_finalize_hyper_device(H, B, th_ptr, partials.data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))

    def step(i, B=1):
        th = ths[(i * B) % NT:(i * B) % NT + B] if (i * B) % NT + B <= NT else ths[:B]
        if world == 1 or fused:
            check(L.ppcseq_log_prob_grad_device(H, B, th.data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))
        else:
            check(L.ppcseq_log_prob_grad_partial_device(H, B, th.data_ptr(), 1, partials.data_ptr(), grad.data_ptr(), sp))
            dist.all_reduce(partials[:B])
            check(L.ppcseq_finalize_hyper_device(H, B, th.data_ptr(), partials.data_ptr(), 1, 1, lp.data_ptr(),
                                                 grad.data_ptr(), sp))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity across ranks, BEFORE anything is timed (N > 1) ------------------------------------------------------
    parity = None
    if world > 1:
        parity = multi_gpu_parity(args, rank, world, local_rank, dev, model, pr, ths_glob, ths, ths_host, lp, grad, sp,
                                  eval_partial_path, fused, strong)
        barrier()

    def timed_device(B, steps, warmup):
        for i in range(warmup):
            step(i, B)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            if not args.no_flush:
                flush.zero_()                           # L2 flush, outside the timed events
                if flush_rd is not None:
                    flush_rd.sum()                      # ... then 256 MiB read: L2 ends up full of CLEAN foreign lines
            ev[i][0].record(stream)
            step(i, B)
            ev[i][1].record(stream)
        barrier()
        times = np.array([a.elapsed_time(b) for a, b in ev])          # ms
        tot = torch.tensor([times.sum()], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), times

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.ppcseq_launch_count()
    total_ms, times = timed_device(1, args.steps, 0)
    launches = L.ppcseq_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    # strong: one step = one evaluation of the WHOLE fixed problem (all ranks together); weak: N shard evaluations
    value = (1 if strong else world) * args.steps / (total_ms * 1e-3)
    batched = None
    if BB > 1:                                          # the chains of a sampler in ONE launch (grid.y = B)
        tb, _ = timed_device(BB, max(4, args.steps // 2), 3)
        nb = max(4, args.steps // 2)
        batched = {"B": BB, "ms_per_launch": tb / nb, "value": (1 if strong else world) * BB * nb / (tb * 1e-3), "unit": UNIT,
                   "note": "B thetas (the chains of a sampler / the draws of an ELBO estimate) evaluated by one launch; "
                           "fixed launch + reduction + all-reduce cost amortised over B"}

    # ---- end to end through the public host API: pinned host thetas in, lp + gradients out (pinned), every step ----
    # The call a user makes is model.log_prob_grad(thetas[B, D]) -> ppcseq_log_prob_grad (C ABI, host pointers).  One
    # call carries EB thetas (the chains of a sampler / the draws of an ELBO estimate); inside the library theta b+1
    # goes host->device and gradient b-1 device->host while evaluation b runs.  `single_call` is the same API with B = 1.
    EB = 8
    th_pin = torch.from_numpy(ths_host[:EB].copy()).pin_memory()
    g_pin = torch.empty((EB, D), dtype=torch.float64).pin_memory()
    lp_pin = torch.empty(EB, dtype=torch.float64).pin_memory()
    th_dev = torch.empty(D, dtype=torch.float64, device=dev)
    th_np, g_np, lp_np = th_pin.numpy(), g_pin.numpy(), lp_pin.numpy()

    def step_e2e_single(i):
        if world == 1 or fused:
            return model.log_prob_grad(th_np[i % EB], out=(lp_np[:1], g_np[:1]))
        th_dev.copy_(th_pin[i % EB], non_blocking=True)
        check(L.ppcseq_log_prob_grad_partial_device(H, 1, th_dev.data_ptr(), 1, partials.data_ptr(), grad.data_ptr(), sp))
        dist.all_reduce(partials[:1])
        check(L.ppcseq_finalize_hyper_device(H, 1, th_dev.data_ptr(), partials.data_ptr(), 1, 1, lp.data_ptr(),
                                             grad.data_ptr(), sp))
        g_pin[0].copy_(grad[0], non_blocking=True); lp_pin[:1].copy_(lp[:1], non_blocking=True)
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
        return (1 if strong else world) * n_calls * evals_per_call / float(t.item())

    e2e_single = timed(step_e2e_single, args.steps, 1)
    if world == 1 or fused:
        n_calls = max(1, (args.steps + EB - 1) // EB)
        e2e_value = timed(lambda i: model.log_prob_grad(th_np, out=(lp_np, g_np)), n_calls, EB)
        e2e_mode = f"{EB} thetas per call, 3-stage copy/compute pipeline inside ppcseq_log_prob_grad"
    else:
        e2e_value, e2e_mode = e2e_single, "one theta per call (nccl variant)"

    # optional: the "exposure-gradient allreduce" of BASELINE config 5 -- every rank's S-vector d lp / d exposure_rate of its
    # gene shard, summed over the ranks with NCCL (exposure is data in the reference: never part of the parity gradient)
    xg_block = None
    if world > 1:
        xg = torch.zeros(w.S, dtype=torch.float64, device=dev)
        for _ in range(3):
            check(L.ppcseq_exposure_grad_device(H, ths[NT - 1].data_ptr(), xg.data_ptr(), sp))
            dist.all_reduce(xg)
        barrier()
        nx = 10
        evx = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(nx)]
        for i in range(nx):
            flush.zero_()
            evx[i][0].record(stream)
            check(L.ppcseq_exposure_grad_device(H, ths[NT - 1].data_ptr(), xg.data_ptr(), sp))
            evx[i][1].record(stream)
            dist.all_reduce(xg)
            evx[i][2].record(stream)
        barrier()
        tk = torch.tensor([sum(a.elapsed_time(b) for a, b, _ in evx) / nx, sum(b.elapsed_time(c) for _, b, c in evx) / nx],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        xg_block = {"kernel_ms": float(tk[0]), "nccl_allreduce_ms": float(tk[1]), "bytes_allreduced": int(8 * w.S),
                    "note": "optional S-vector d lp / d exposure_rate of the rank's gene shard + NCCL all_reduce over the ranks"}

    weak = None
    if world > 1 and strong and not args.no_weak:
        weak = weak_scaling_leg(args, rank, world, local_rank, dev, stream, sp, flush)

    if rank == 0:
        peaks, which = measured_peaks()
        B_eval = w.algorithmic_bytes_per_eval() if not args.no_mask else w.algorithmic_bytes_per_eval() - w.G * w.S // 8
        if not strong:
            B_eval = B_eval                              # per shard = per launch
        B_launch = B_eval / world if (strong and world > 1) else B_eval     # one launch covers G / N genes
        achieved = B_launch / (ms_per_step * 1e-3) / 1e9
        prof = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            prof = json.load(open(tp))
        traffic = prof.get(args.workload) if (world == 1 and args.path in ("auto", "moments")) else None
        fp64_busy = prof.get(args.workload + ":fp64_pipe_frac") if traffic else None
        ncu_us = prof.get(args.workload + ":ncu_kernel_us") if traffic else None
        fp64 = ctypes.c_double()
        check(L.ppcseq_measure_fp64_peak(local_rank, ctypes.byref(fp64)))
        frac = achieved / peaks["hbm_gbs"]
        frac_traffic = (traffic / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]) if traffic else None
        cands = {"hbm (algorithmic bytes)": frac, "hbm (measured traffic)": frac_traffic or 0.0, "fp64 pipe": fp64_busy or 0.0}
        binding = max(cands, key=cands.get)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(args, w, bool(len(pr["pairs"])) or bool(pr["all_pairs"] is not None and len(pr["all_pairs"]))),
            "setup": {"per_rank": (f"genes split in {world} contiguous blocks of the ONE fixed problem" if strong else
                                   "each rank owns one such shard of an N x G gene model") if world > 1 else "single GPU",
                      "collective": ("fused in-kernel peer all-reduce (NVLink mailboxes)" if fused else
                                     "nccl all_reduce of 8 doubles + finalize kernel") if world > 1 else "none",
                      "likelihood_path": args.path},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": frac, "traffic": traffic, "peak_source": which,
                         "algorithmic_bytes_per_launch": B_launch,
                         "frac_traffic": frac_traffic, "fp64_pipe_frac": fp64_busy, "binding": binding,
                         "ncu_kernel_us": ncu_us,
                         "fp64_peak_tflops_measured": fp64.value,
                         "note": "frac = SURVEY 8(d) algorithmic bytes (dense counts once + theta/grad + X [+ mask]) / event time / "
                                 "measured HBM peak.  The kernel reads data-only sufficient statistics instead of the counts: "
                                 "frac_traffic = measured DRAM bytes (ncu, `traffic`) / event time / peak; fp64_pipe_frac = "
                                 "sm__pipe_fp64_cycles_active (ncu, same capture); `binding` = the largest of the three"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * D) * (world if strong else 1),
                    "d2h_bytes_per_step": int(8 * D + 8) * (world if strong else 1), "mode": e2e_mode, "single_call": e2e_single,
                    "model_create_s": t_create, "model_create_h2d_bytes": int(pr["counts"].nbytes)},
            "gpu_launches": int(launches), "clocks": clocks,
            "step_ms": {"min": float(times.min()), "median": float(np.median(times)), "max": float(times.max())},
        }
        if batched:
            # algorithmic bytes per launch = the per-evaluation figure x the evaluations one launch processes
            bb = B_launch * batched["B"] / (batched["ms_per_launch"] * 1e-3) / 1e9
            batched["roofline"] = {"bound": "hbm", "achieved": bb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bb / peaks["hbm_gbs"],
                                   "algorithmic_bytes_per_launch": B_launch * batched["B"],
                                   "note": "B evaluations per launch, each counted with the full per-evaluation byte figure "
                                           "(every theta needs its own pass over the sufficient statistics); the headline "
                                           "`roofline` block is the single-evaluation launch"}
            out["batched"] = batched
        if parity is not None:
            out["parity"] = parity
        if weak is not None:
            out["weak"] = weak
        if xg_block is not None:
            out["exposure_gradient_allreduce"] = xg_block
        if world == 1 and not args.no_cpu_baseline:
            excl = None
            if len(pr["pairs"]):
                excl = np.zeros((w.G, w.S), bool)
                excl[pr["pairs"][:, 0], pr["pairs"][:, 1]] = True
            out["cpu_baseline"] = cpu_baseline(w, excl)
        if world == 1 and not args.no_extras:
            try:
                out["ppc"] = ppc_bench(w, local_rank)
                out["paths"] = other_paths_bench(args, w, pr, local_rank, peaks)
                out["identify_outliers"] = identify_outliers_bench(local_rank)
                out["identify_outliers_cfg2"] = identify_outliers_cfg2_bench(local_rank)
                out["input_edge"] = input_edge_bench(w)
            except Exception as e:                      # the headline line must still be printed
                out["extras_error"] = repr(e)
        print(json.dumps(out), flush=True)
    bad = False
    if fused and pdist.comm_timed_out(model):
        bad = True
        print("fused all-reduce timed out on rank %d (ranks out of step)" % rank, file=sys.stderr)
    if parity is not None and not parity["ok"]:
        bad = True
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        sys.exit(1)


def multi_gpu_parity(args, rank, world, local_rank, dev, model, pr, ths_glob, ths, ths_host, lp, grad, sp, eval_partial_path,
                     fused, strong):
    """Correctness evidence the driver can see at N > 1, before any timing (every rank takes part, rank 0 reports):
      (1) lp and the 6 hyper-gradients of the fused in-kernel all-reduce are BITWISE identical on every rank;
      (2) they agree (<= 1e-12) with the un-fused formulation: ppcseq_log_prob_grad_partial_device on every rank,
          all-gather, rank-ordered sum, ppcseq_finalize_hyper_device;
      (3) strong scaling only: rank 0 also evaluates the UNSHARDED problem on its own GPU (the path the single-GPU
          parity tests check against the oracle): lp / hyper-gradients agree to 1e-12 and every gene-block gradient
          entry gathered from the shards is bitwise the unsharded one."""
    import torch
    import torch.distributed as dist

    import ppcseq_b200
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200._lib import check
    L = ppcseq_b200.lib()
    H = model.handle
    D = model.D
    NT = ths.shape[0]
    pts = [NT - 1, 0, 1]                                  # the generating truth and two random points
    hyp_idx = torch.tensor([0, 1, 2, D - 3, D - 2, D - 1], device=dev)
    res = {"ok": True, "thetas": len(pts), "ranks_bitwise_identical": True, "vs_partial_allreduce_finalize_max_rel": 0.0}
    fused_out = []
    for i in pts:
        if fused:
            check(L.ppcseq_log_prob_grad_device(H, 1, ths[i].data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))
        else:
            eval_partial_path(ths[i].data_ptr(), 1)
        torch.cuda.synchronize()
        mine = torch.cat([lp[:1], grad[0][hyp_idx]]).clone()
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        if any(not torch.equal(allv[0].view(torch.int64), a.view(torch.int64)) for a in allv):
            res["ranks_bitwise_identical"] = False
        fused_out.append((mine.cpu().numpy(), grad[0].clone()))
        eval_partial_path(ths[i].data_ptr(), 1)
        torch.cuda.synchronize()
        ref = torch.cat([lp[:1], grad[0][hyp_idx]]).cpu().numpy()
        rel = float(np.max(np.abs(ref - fused_out[-1][0]) / np.maximum(np.abs(ref), 1e-300)))
        res["vs_partial_allreduce_finalize_max_rel"] = max(res["vs_partial_allreduce_finalize_max_rel"], rel)
    res["all_finite"] = bool(all(np.isfinite(f[0]).all() for f in fused_out))
    if strong:
        w = pr["w"]
        Dmax = torch.tensor([D], device=dev)
        dist.all_reduce(Dmax, op=dist.ReduceOp.MAX)
        Dmax = int(Dmax.item())
        gathered = []
        for k in range(len(pts)):
            pad = torch.zeros(Dmax, dtype=torch.float64, device=dev)
            pad[:D] = fused_out[k][1]
            outl = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(outl, pad)
            gathered.append([o.cpu().numpy() for o in outl])
        if rank == 0:
            single = ppcseq_b200.NBModel(w.counts, w.X, w.exposure, w.K, device=local_rank)
            if len(pr["all_pairs"]):
                single.set_exclusion(pr["all_pairs"])
            single.set_design_path({"auto": 0, "general": 1, "element": 2, "moments": 3}[args.path])
            lp_rel, hy_rel, bitwise = 0.0, 0.0, True
            lay = ppcseq_b200.layout(w.G, w.K, w.C)
            for k, i in enumerate(pts):
                lp1, g1 = single.log_prob_grad(ths_glob[i])
                g = np.zeros_like(g1)
                for q in range(world):
                    g0, g1_ = pdist.shard_range(w.G, q, world)
                    Dq = ppcseq_b200.layout(g1_ - g0, pdist.local_K(w.K, g0, g1_), w.C).D
                    pdist.scatter_local_grad(g, gathered[k][q][:Dq], w.G, w.K, w.C, g0, g1_, write_hyper=(q == 0))
                lp_rel = max(lp_rel, abs(fused_out[k][0][0] - lp1) / abs(lp1))
                hy = np.concatenate([g[:3], g[-3:]]); hy1 = np.concatenate([g1[:3], g1[-3:]])
                hy_rel = max(hy_rel, float(np.max(np.abs(hy - hy1) / np.maximum(np.abs(hy1), 1e-3 * np.abs(hy1).max()))))
                bitwise = bitwise and bool(np.array_equal(g[3:lay.o_tail], g1[3:lay.o_tail]))
            single.close()
            res["vs_single_gpu"] = {"lp_max_rel": lp_rel, "hyper_grad_max_rel": hy_rel, "gene_block_gradients_bitwise": bitwise}
            if not (lp_rel < 1e-12 and hy_rel < 1e-11 and bitwise):
                res["ok"] = False
    if not (res["ranks_bitwise_identical"] and res["all_finite"] and res["vs_partial_allreduce_finalize_max_rel"] <= 1e-12):
        res["ok"] = False
    okt = torch.tensor([1 if res["ok"] else 0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    res["ok"] = bool(okt.item())
    return res


def weak_scaling_leg(args, rank, world, local_rank, dev, stream, sp, flush):
    """Round-1's curve kept as an extra key: every rank owns its own full-size shard of an (N x G)-gene model."""
    import copy

    import torch
    import torch.distributed as dist

    import ppcseq_b200
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200._lib import check
    L = ppcseq_b200.lib()
    a2 = copy.copy(args)
    a2.scaling = "weak"
    pr = _rank_problem(a2, rank, world)
    w = pr["w"]
    m = ppcseq_b200.NBModel(pr["counts"], w.X, w.exposure, pr["K_total"], device=local_rank, shard=(pr["G_total"], pr["g0"]))
    if len(pr["pairs"]):
        m.set_exclusion(pr["pairs"])
    _, th = _theta_points(pr, a2, world)
    hyper = torch.from_numpy(np.concatenate([th[:, :3], th[:, -3:]], axis=1)).to(dev)
    dist.broadcast(hyper, 0)
    h = hyper.cpu().numpy()
    th[:, :3] = h[:, :3]; th[:, -3:] = h[:, 3:]
    ths = torch.from_numpy(th).to(dev)
    pdist.connect(m, rank, world, channels=1, cap=1)
    lp = torch.zeros(1, dtype=torch.float64, device=dev)
    grad = torch.zeros(m.D, dtype=torch.float64, device=dev)
    steps = max(8, args.steps // 2)
    for i in range(4):
        check(L.ppcseq_log_prob_grad_device(m.handle, 1, ths[i].data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))
    dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.zero_()
        ev[i][0].record(stream)
        check(L.ppcseq_log_prob_grad_device(m.handle, 1, ths[i % len(ths)].data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), sp))
        ev[i][1].record(stream)
    dist.barrier(); torch.cuda.synchronize()
    tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    out = {"value": world * steps / (float(tot.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(tot.item()) / steps,
           "steps": steps, "timed_out": bool(pdist.comm_timed_out(m)),
           "note": "weak scaling: every rank owns its own 1-GPU-sized shard of an N x G gene model (round-1 definition)"}
    m.close()
    return out


def other_paths_bench(args, w, pr, device, peaks, steps=12):
    """The likelihood paths the headline does not exercise, each with its own roofline block (N = 1 extras):
    `element` = per-element categorical kernel (taken when the exposure range is too wide for the moment series),
    `general` = any model.matrix (continuous covariate: per-element exp), `streaming_theta` = the moment kernel at a
    theta that sends ~10 % of the genes through its streaming fallback (phase B2)."""
    import torch

    import ppcseq_b200
    from ppcseq_b200 import synthetic
    from ppcseq_b200._lib import check
    L = ppcseq_b200.lib()
    dev = torch.device("cuda", device)
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B_eval = w.algorithmic_bytes_per_eval()

    def time_model(m, th_host):
        ths = torch.from_numpy(np.ascontiguousarray(th_host)).to(dev)
        lp = torch.zeros(1, dtype=torch.float64, device=dev)
        grad = torch.zeros(m.D, dtype=torch.float64, device=dev)
        ms = (ctypes.c_float * steps)()
        check(L.ppcseq_time_log_prob_grad_device(m.handle, 1, ths[0].data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(), None, 3, 1, ms))
        tot = 0.0
        for i in range(steps):
            one = (ctypes.c_float * 1)()
            check(L.ppcseq_time_log_prob_grad_device(m.handle, 1, ths[i % len(ths)].data_ptr(), 1, 1, lp.data_ptr(), grad.data_ptr(),
                                                     None, 1, 1, one))
            tot += one[0]
        t = tot / steps
        ach = B_eval / (t * 1e-3) / 1e9
        return {"ms_per_step": t, "value": 1e3 / t, "unit": UNIT,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                             "algorithmic_bytes_per_launch": B_eval}}

    m = ppcseq_b200.NBModel(w.counts, w.X, w.exposure, w.K, device=device)
    if len(pr["pairs"]):
        m.set_exclusion(pr["pairs"])
    ths = np.vstack([synthetic.random_thetas(w, 4, seed=1), w.theta_true])
    m.set_design_path(2)
    out["element"] = time_model(m, ths)
    out["element"]["binding"] = "FP64 pipe / issue slots (ncu, profiles/r1_lpgrad_cat_ncu_full_summary.txt: FP64 pipe 51 %, issue 63 %)"
    m.set_design_path(1)
    out["general_on_categorical_design"] = time_model(m, ths)
    m.set_design_path(0)
    lay = m.layout
    rng = np.random.default_rng(17)
    th_s = w.theta_true.copy()
    pick = rng.random(w.G) < 0.10
    th_s[lay.o_sigma_raw:lay.o_sigma_raw + w.G][pick] = rng.uniform(-7.0, -4.0, int(pick.sum()))
    out["streaming_theta"] = time_model(m, th_s[None, :])
    out["streaming_theta"]["genes_streamed_frac"] = float(pick.mean())
    m.close()
    # continuous covariate: the same counts with a numeric second column => general path by necessity
    Xc = w.X.copy()
    Xc[:, 1] = np.random.default_rng(3).normal(0.0, 1.0, w.S)
    mc = ppcseq_b200.NBModel(w.counts, Xc, w.exposure, w.K, device=device)
    if len(pr["pairs"]):
        mc.set_exclusion(pr["pairs"])
    out["general_continuous_covariate"] = time_model(mc, ths)
    mc.close()
    # optional S-vector output d lp / d exposure_rate (exposure is data in the reference; BASELINE config 5's label)
    mx = ppcseq_b200.NBModel(w.counts, w.X, w.exposure, w.K, device=device)
    if len(pr["pairs"]):
        mx.set_exclusion(pr["pairs"])
    th_d = torch.from_numpy(np.ascontiguousarray(w.theta_true)).to(dev)
    xg = torch.zeros(w.S, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream(dev)
    spx = ctypes.c_void_p(st.cuda_stream)
    for _ in range(3):
        check(L.ppcseq_exposure_grad_device(mx.handle, th_d.data_ptr(), xg.data_ptr(), spx))
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(L.ppcseq_exposure_grad_device(mx.handle, th_d.data_ptr(), xg.data_ptr(), spx))
        e1.record(st)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = tot / steps
    bytes_x = 4 * w.G * w.S
    out["exposure_gradient_optional"] = {
        "ms_per_call": t, "unit": "ms", "launches_per_call": 2,
        "roofline": {"bound": "hbm", "achieved": bytes_x / (t * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_x / (t * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": bytes_x},
        "note": "d lp / d exposure_rate[s], S-vector; NOT part of the parity gradient (exposure is data in the reference)"}
    mx.close()
    del flush
    return out


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
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between steps: 256 MiB memset (default; leaves L2 full of dirty lines that the timed "
                         "kernel's fills must write back) or memset followed by a 256 MiB read (clean foreign lines)")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the ONE named problem split over the ranks (BASELINE configs; default), "
                         "weak = every rank owns its own full-size shard (round-1 curve; also reported as `weak` in a strong run)")
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling leg of a strong multi-GPU run")
    ap.add_argument("--batch", type=int, default=4, help="also time B thetas per launch (`batched` block); 1 = skip")
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
