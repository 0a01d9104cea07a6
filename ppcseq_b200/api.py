"""identify_outliers(): host-side mirror of the reference's exported entry point (R/methods.R:74-367) and of
do_inference() (R/utilities.R:1321-1547), driving the GPU-resident model, samplers and PPC kernels.

The reference is an R package; R is not available in this image, so the host side is Python with the same
argument names, defaults, thresholds and output columns.  The `.Call` shim a maintainer would add to the R
package instead is shown in INTEGRATION.md.  Everything numeric below runs in libppcseq_b200.so.
"""
from __future__ import annotations

import math
import os
import random
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import ppc as _ppc
from . import prep as _prep
from .inference import WARMUP, find_optimal_number_of_chains, sample_nuts, vb_iterative
from .model import NBModel

TRUNCATION_COMPENSATION = 0.7352941      # R/methods.R:339
LAMBDA_MU_MU = 5.612671                  # R/methods.R:218


def _col(data, name):
    v = data[name]
    return v.to_numpy() if hasattr(v, "to_numpy") else np.asarray(v)


@dataclass
class PassResult:
    """What do_inference returns for the checked genes (rows `.variable == "counts_rng"`), as [K, S] arrays."""
    lower: np.ndarray
    upper: np.ndarray
    mean: np.ndarray
    sd: np.ndarray
    ppc: np.ndarray
    deleterious: np.ndarray | None
    slope: np.ndarray
    total_draws: float
    fit: object = None
    info: np.ndarray = field(default_factory=lambda: np.zeros(8))


def do_inference(model: NBModel, *, approximate_posterior_inference: bool, approximate_posterior_analysis: bool,
                 adj_prob_theshold: float, how_many_posterior_draws: float, cores: int, seed: int,
                 to_exclude=None, truncation_compensation: float = 1.0, pass_fit: bool = False) -> PassResult:
    """R/utilities.R:1321-1547 with the map_rect packing removed (dense arrays are already on the device)."""
    model.set_exclusion(np.empty((0, 2), np.int32) if to_exclude is None else to_exclude)
    draws_practical = 1000 if approximate_posterior_analysis else int(how_many_posterior_draws)      # :1372
    chains = max(3, min(int(cores), find_optimal_number_of_chains(draws_practical)))                  # :1377-1380
    if approximate_posterior_inference:                                                               # :1487-1494
        fit = vb_iterative(model, output_samples=draws_practical, iter=50000, tol_rel_obj=0.005, seed=seed)
    else:                                                                                             # :1497-1512
        fit = sample_nuts(model, chains=chains, iter=int(math.ceil(draws_practical / chains)) + WARMUP, warmup=WARMUP,
                          seed=seed)
    if approximate_posterior_analysis:        # fit_to_counts_rng_approximated, :733-784
        lo, up, mean, sd = fit.ppc_summary(adj_prob_theshold, exact=False, n_draws=int(how_many_posterior_draws),
                                           truncation_compensation=truncation_compensation, seed=seed)
    else:                                     # fit_to_counts_rng, :685-703
        lo, up, mean, sd = fit.ppc_summary(adj_prob_theshold, exact=True,
                                           truncation_compensation=truncation_compensation, seed=seed)
    slope = fit.slope() if model.K else np.empty(0)                                                  # :1531
    fl = _ppc.flags(model, lo, up, mean, slope if model.C > 1 else None)                              # :1528, :1534
    res = PassResult(lo, up, mean, sd, fl["ppc"], fl["deleterious"], slope,
                     float(model.S) * model.K * how_many_posterior_draws, fit if pass_fit else None, fit.info())
    if not pass_fit:
        fit.close()
    return res


def identify_outliers(data, formula: str = "~ 1", *, sample: str, transcript: str, abundance: str, significance: str,
                      do_check: str, scaling_factor: str | None = None, percent_false_positive_genes: float = 1,
                      how_many_negative_controls: int = 500, approximate_posterior_inference: bool = True,
                      approximate_posterior_analysis: bool | None = True, draws_after_tail: float = 10,
                      cores: int | None = None, pass_fit: bool = False, do_check_only_on_detrimental: bool | None = None,
                      tol_rel_obj: float = 0.01, just_discovery: bool = False, seed: int | None = None,
                      adj_prob_theshold_2: float | None = None, device: int = 0, devices=None):
    """Same arguments and defaults as the reference (R/methods.R:74-98); `data` is a pandas DataFrame or a dict of
    row-aligned columns.  Returns a pandas DataFrame with the reference's columns: <transcript>, sample_wise_data
    (nested frame), ppc_samples_failed and -- when the formula has a covariate -- tot_deleterious_outliers.
    `devices=[0, 1, ...]` splits the genes over several GPUs inside this one process (both passes, PPC included)."""
    import pandas as pd
    covs = _prep.parse_formula(formula)
    if do_check_only_on_detrimental is None:
        do_check_only_on_detrimental = len(covs) > 0                                       # R/methods.R:93
    if cores is None:
        cores = os.cpu_count() or 1
    if seed is None:
        seed = random.randint(1, 999999)                                                   # :96
    del tol_rel_obj        # accepted and ignored, exactly as the reference does (0.005 is hard-coded, R/utilities.R:1492)
    smp, trn, abn = _col(data, sample), _col(data, transcript), _col(data, abundance)
    sig, chk = _col(data, significance), _col(data, do_check).astype(bool)
    for name, v in ((sample, smp), (transcript, trn), (abundance, abn), (significance, sig)):
        if pd.isna(v).any():
            raise ValueError(f"column {name} contains NA")                                 # check_if_any_NA
    if not chk.any():                                                                      # :117-127
        warnings.warn("ppcseq says: There are not transcripts with the category .to_check. NULL is returned.")
        return pd.DataFrame({transcript: [], "sample_wise_data": [], "ppc samples failed": [],
                             "tot deleterious_outliers": []})
    if not (0 <= percent_false_positive_genes <= 100):
        raise ValueError("percent_false_positive_genes must be a string from > 0% to < 100%")
    n_samples = len(set(smp.tolist()))
    if adj_prob_theshold_2 is None:                                                        # :156-160
        adj_prob_theshold_2 = percent_false_positive_genes / 100 / n_samples * (2 if do_check_only_on_detrimental else 1)
    adj_prob_theshold_1 = max(0.05, adj_prob_theshold_2 * 2)                               # :163
    draws_1 = max(draws_after_tail / adj_prob_theshold_1, 1000)                            # :166-167
    draws_2 = max(draws_after_tail / adj_prob_theshold_2, 1000)
    if approximate_posterior_analysis is None:                                             # :170-176
        approximate_posterior_analysis = draws_2 > 20000
    p = _prep.prepare(smp.tolist(), trn.tolist(), abn, sig, chk, {c: _col(data, c) if _col(data, c).dtype.kind in "fiu"
                                                                   else _col(data, c).tolist() for c in covs},
                      formula, how_many_negative_controls,
                      scaling_factor=None if scaling_factor is None else _col(data, scaling_factor))
    model = NBModel(p.counts, p.X, p.exposure_rate, p.K, lambda_mu_mu=LAMBDA_MU_MU, device=device, devices=devices)
    try:
        res1 = do_inference(model, approximate_posterior_inference=approximate_posterior_inference,
                            approximate_posterior_analysis=False, adj_prob_theshold=adj_prob_theshold_1,
                            how_many_posterior_draws=draws_1, cores=cores, seed=seed, pass_fit=pass_fit)     # :268-286
        K, S = p.K, len(p.samples)
        if just_discovery:
            res2 = res1
        else:
            bad = res1.deleterious if do_check_only_on_detrimental else ~res1.ppc                           # :292-300
            to_exclude = np.argwhere(bad).astype(np.int32)                                                  # (g, s) pairs
            res2 = do_inference(model, approximate_posterior_inference=approximate_posterior_inference,
                                approximate_posterior_analysis=approximate_posterior_analysis,
                                adj_prob_theshold=adj_prob_theshold_2, how_many_posterior_draws=draws_2, cores=cores,
                                seed=seed, to_exclude=to_exclude, truncation_compensation=TRUNCATION_COMPENSATION,
                                pass_fit=pass_fit)                                                          # :320-342
    finally:
        if not pass_fit:
            model.close()
    # ---- merge_results + format_results (R/utilities.R:539-608) ---------------------------------------------
    first_row = {}
    for i, s in enumerate(smp.tolist()):
        first_row.setdefault(s, i)
    rows = []
    for g in range(K):
        d = {"S": np.arange(1, S + 1), "G": np.full(S, g + 1), abundance: p.counts[g], sample: p.samples,
             "slope_before_outlier_filtering": np.full(S, res1.slope[g] if len(res1.slope) else np.nan)}
        for c in covs:
            cv = _col(data, c)
            d[c] = [cv[first_row[s]] for s in p.samples]
        d.update({"exposure_rate": p.exposure_rate, "multiplier": p.multiplier, ".lower": res2.lower[g],
                  ".upper": res2.upper[g],
                  "slope_after_outlier_filtering": np.full(S, res2.slope[g] if len(res2.slope) else np.nan),
                  "posterior_predictive_check_succeded": res2.ppc[g]})
        if res2.deleterious is not None:
            d["deleterious_outliers"] = res2.deleterious[g]
        rows.append(pd.DataFrame(d))
    out = pd.DataFrame({transcript: p.genes[:K], "sample_wise_data": rows,
                        "ppc_samples_failed": [int((~r["posterior_predictive_check_succeded"]).sum()) for r in rows]})
    if do_check_only_on_detrimental:
        out["tot_deleterious_outliers"] = [int(r["deleterious_outliers"].sum()) for r in rows]
    out.attrs.update({"total_draws": res2.total_draws, "transcript_column": transcript, "abundance_column": abundance,
                      "sample_column": sample, "formula": formula, "fit 1 info": res1.info, "fit 2 info": res2.info,
                      "seed": seed})
    if pass_fit:
        out.attrs.update({"fit 1": res1.fit, "fit 2": res2.fit, "model": model})
    return out
