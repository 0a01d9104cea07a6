#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.  Run in the BUILD container only
(`python tests/golden/make_golden.py`): it reads the reference's bundled dataset
/root/reference/data/counts.rda (documented in man/counts.Rd), which does not exist on the GPU box.

  bundled_test53.npz     the reference's own test configuration (tests/testthat/test-ppcSeq.R:7-32):
                         SLC16A12 / CYP1A1 / ART3 to check + 50 negative controls, ~Label
  bundled_readme515.npz  the README run (README.md:47-92): 15 genes with FDR < 0.01 + 500 controls
       each holds the model inputs derived by ppcseq_b200.prep (counts [G,S] int32, X, exposure_rate, K,
       gene / sample names) plus the raw selected columns, so the tests can re-derive them without R data.
  lp_grad_golden.npz     40-digit mpmath log_prob / gradient (oracle/model_mp.py) of the 53-gene problem at
                         seeded thetas, in the three (propto, jacobian) modes, pass 1 and with an exclusion.
  quantile_golden.npz    type-7 quantile / mean / sd / flag vectors of seeded draws (oracle/quantile.py).
  nuts_golden.npz        posterior mean / sd of a 24-gene problem from the NumPy oracle NUTS (oracle/nuts_np.py,
                         4 chains x 1000 draws) -- the CPU reference the GPU samplers are compared with.

PARITY STATUS: the reference holds no golden vectors for log_prob/grad/quantiles; these files
pin the repo's own oracle so that regressions in it are caught.  The reference-pinned facts carried here are
the dataset itself and the expected discrete outcomes (EXPECTED_* below).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import model_mp, model_np, quantile, rda  # noqa: E402
from ppcseq_b200 import prep  # noqa: E402

# tests/testthat/test-ppcSeq.R:26-30, :51-55  (column 4 = tot_deleterious_outliers)
EXPECTED_TEST = {"SLC16A12": 0, "CYP1A1": 1, "ART3": 0}
# README.md:75-92 (ppc_samples_failed, tot_deleterious_outliers)
EXPECTED_README = {"SLC16A12": (0, 0), "CYP1A1": (1, 1), "ART3": (0, 0), "DIO2": (0, 0), "OR51E2": (0, 0),
                   "MUC16": (0, 0), "CCNA1": (0, 0), "LYZ": (1, 1), "PPM1H": (0, 0), "SUSD5": (0, 0),
                   "TPRG1": (0, 0), "EPB42": (0, 0), "LRRC38": (0, 0), "SUSD4": (0, 0), "MMP8": (0, 0)}


def save_bundled(path, p, expected_names, expected_vals):
    np.savez_compressed(
        path, counts=p.counts, X=p.X, exposure_rate=p.exposure_rate, multiplier=p.multiplier, K=np.int32(p.K),
        genes=np.array(p.genes), samples=np.array(p.samples), design_columns=np.array(p.design_columns),
        tmm=p.tmm, reference_sample=np.array(p.reference_sample),
        expected_genes=np.array(expected_names), expected=np.array(expected_vals, dtype=np.int32))


def main():
    raw = rda.load_rda("/root/reference/data/counts.rda")["counts"]
    sym = raw["symbol"]
    cols = dict(sample=raw["sample"], transcript=sym, abundance=raw["value"], significance=raw["PValue"])
    cov = {"Label": raw["Label"]}

    chk = np.array([s in EXPECTED_TEST for s in sym])
    p53 = prep.prepare(**cols, do_check=chk, covariates=cov, formula="~ Label", how_many_negative_controls=50)
    assert p53.counts.shape == (53, 21) and p53.K == 3 and p53.genes[:3] == list(EXPECTED_TEST)
    save_bundled(os.path.join(HERE, "bundled_test53.npz"), p53, list(EXPECTED_TEST), list(EXPECTED_TEST.values()))

    chk = raw["FDR"] < 0.01
    p515 = prep.prepare(**cols, do_check=chk, covariates=cov, formula="~ Label", how_many_negative_controls=500)
    assert p515.counts.shape == (515, 21) and p515.K == 15 and p515.genes[:15] == list(EXPECTED_README)
    save_bundled(os.path.join(HERE, "bundled_readme515.npz"), p515, list(EXPECTED_README),
                 list(EXPECTED_README.values()))

    # library sizes of the whole dataset (for the exposure proxy comparison in DESIGN.md)
    samples = list(dict.fromkeys(raw["sample"]))                  # first appearance
    tot = {s: 0 for s in samples}
    for s, v in zip(raw["sample"], raw["value"]):
        tot[s] += int(v)
    np.savez_compressed(os.path.join(HERE, "bundled_library_sizes.npz"), samples=np.array(samples),
                        library_size=np.array([tot[s] for s in samples], dtype=np.int64))

    # ---- mpmath golden vectors ---------------------------------------------------------------------
    rng = np.random.default_rng(20240)
    D = model_np.dim(53, 3, 2)
    thetas = rng.uniform(-2, 2, (3, D))
    # one point near a plausible posterior mode (intercept ~ log mean count)
    lay = model_np.Layout(53, 3, 2)
    th = np.zeros(D)
    th[lay.o_intercept:lay.o_intercept + 53] = np.log(p53.counts.mean(axis=1) + 1.0)
    th[lay.o_sigma_raw:lay.o_sigma_raw + 53] = -1.0
    th[1] = np.log(1.5); th[2] = -1.0; th[lay.o_tail] = np.log(0.3); th[lay.o_tail + 1] = 1.0; th[lay.o_tail + 2] = np.log(0.6)
    thetas = np.vstack([thetas, th])
    excl = np.zeros((53, 21), bool)
    excl[1, 8] = True                      # CYP1A1's gross outlier sample
    excl[[5, 17, 40], [0, 20, 7]] = True
    modes = [(1, 1), (0, 1), (1, 0)]
    lp = np.empty((2, len(modes), len(thetas)))
    gr = np.empty((2, len(modes), len(thetas), D))
    for e, ex in enumerate([None, excl]):
        d = model_np.ModelData(p53.counts, p53.X, p53.exposure_rate, p53.K, exclude=ex)
        for mi, (pr, ja) in enumerate(modes):
            for ti, t in enumerate(thetas):
                l, g = model_mp.to_float(*model_mp.log_prob_grad(d, t, bool(pr), bool(ja)))
                lp[e, mi, ti] = l
                gr[e, mi, ti] = g
    np.savez_compressed(os.path.join(HERE, "lp_grad_golden.npz"), thetas=thetas, exclude=excl,
                        modes=np.array(modes), lp=lp, grad=gr)

    # ---- quantile / flags golden vectors --------------------------------------------------------------
    rng = np.random.default_rng(7)
    K, S, n = 3, 21, 1050
    mu = np.maximum(p53.counts[:K].astype(float), 1.0)
    draws = rng.negative_binomial(3.0, 3.0 / (3.0 + mu.reshape(-1)), size=(n, K * S)).astype(np.float64)
    pq = 1.0 / 100 / 21 * 2
    lo, up, mean, sd = quantile.summarise_draws(draws, pq)
    slope = np.array([0.7, -1.2, 0.0])
    fl = quantile.flags(p53.counts[:K], lo.reshape(K, S), up.reshape(K, S), mean.reshape(K, S), slope, p53.X)
    np.savez_compressed(os.path.join(HERE, "quantile_golden.npz"), draws=draws.astype(np.int32), p=pq, lower=lo,
                        upper=up, mean=mean, sd=sd, slope=slope, ppc=fl["ppc"], deleterious=fl["deleterious"],
                        ppc_samples_failed=fl["ppc_samples_failed"],
                        tot_deleterious_outliers=fl["tot_deleterious_outliers"])
    # ---- NUTS posterior of a small problem from the NumPy oracle sampler (oracle/nuts_np.py) ------------------
    from oracle import c_oracle, nuts_np
    from tests.helpers import small_problem
    G, S, C, K = 24, 12, 2, 12
    d = small_problem(G, S, C, K, seed=42)
    Dn = model_np.dim(G, K, C)
    chains = []
    ndiv = 0
    for ch in range(4):
        dr, st = nuts_np.sample(lambda q: c_oracle.log_prob_grad(d, q), Dn, 150 + 1000, 150, seed=100 + ch)
        chains.append(dr)
        ndiv += st["divergent"]
    allc = np.stack(chains)                                   # [4, 1000, D]
    np.savez_compressed(os.path.join(HERE, "nuts_golden.npz"), counts=d.counts, X=d.X, exposure=d.exposure, K=np.int32(K),
                        mean=allc.reshape(-1, Dn).mean(axis=0), sd=allc.reshape(-1, Dn).std(axis=0, ddof=1),
                        chain_means=allc.mean(axis=1), divergent=np.int32(ndiv))
    print("fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:28s} {os.path.getsize(os.path.join(HERE, f)):9d} B")


if __name__ == "__main__":
    main()
