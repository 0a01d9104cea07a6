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


def _has_na(v) -> bool:
    if v.dtype.kind in "iub":
        return False
    if v.dtype.kind == "f":
        return bool(np.isnan(v).any())
    import pandas as pd
    return bool(pd.isna(v).any())


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
    ppc_samples_failed: np.ndarray = None          # int32 [K], from the flags kernel (R/utilities.R:597)
    tot_deleterious_outliers: np.ndarray = None    # int32 [K] or None (R/utilities.R:604)
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
                     float(model.S) * model.K * how_many_posterior_draws, fl["ppc_samples_failed"],
                     fl["tot_deleterious_outliers"], fit if pass_fit else None, fit.info())
    if res.info[8] > 0:
        warnings.warn(f"{int(res.info[8])} posterior-predictive gamma draws were clamped at 2^30 "
                      "(Stan's neg_binomial_2_log_rng raises there): the posterior is not usable")
    if not pass_fit:
        fit.close()
    return res


def identify_outliers(data, formula: str = "~ 1", *, sample: str, transcript: str, abundance: str, significance: str,
                      do_check: str, scaling_factor: str | None = None, percent_false_positive_genes: float = 1,
                      how_many_negative_controls: int = 500, approximate_posterior_inference: bool = True,
                      approximate_posterior_analysis: bool | None = True, draws_after_tail: float = 10,
                      cores: int | None = None, pass_fit: bool = False, do_check_only_on_detrimental: bool | None = None,
                      tol_rel_obj: float = 0.01, just_discovery: bool = False, seed: int | None = None,
                      adj_prob_theshold_2: float | None = None, device: int = 0, devices=None,
                      return_format: str = "nested", timings: dict | None = None):
    """Same arguments and defaults as the reference (R/methods.R:74-98); `data` is a pandas DataFrame or a dict of
    row-aligned columns.  Returns a pandas DataFrame with the reference's columns: <transcript>, sample_wise_data
    (nested frame), ppc_samples_failed and -- when the formula has a covariate -- tot_deleterious_outliers.
    `devices=[0, 1, ...]` splits the genes over several GPUs inside this one process (both passes, PPC included).
    `return_format`: "nested" = the reference's tibble (one nested frame per checked gene, R/utilities.R:539-608);
    "long" = the same columns as ONE frame with a row per (gene, sample) -- no per-gene objects, for 10^7..10^8 pairs;
    "failing" = only the rows that fail the posterior-predictive check (what a user inspects at scale).  The per-gene
    totals (ppc_samples_failed, tot_deleterious_outliers) come from the flags kernel in every format and are attached
    as `attrs["gene_totals"]` for "long" / "failing".  `timings`: optional dict filled with the wall-clock split
    (prep / upload / pass 1 / pass 2 / result)."""
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
        if _has_na(v):
            raise ValueError(f"column {name} contains NA")                                 # check_if_any_NA
    if not chk.any():                                                                      # :117-127
        warnings.warn("ppcseq says: There are not transcripts with the category .to_check. NULL is returned.")
        return pd.DataFrame({transcript: [], "sample_wise_data": [], "ppc samples failed": [],
                             "tot deleterious_outliers": []})
    if not (0 <= percent_false_positive_genes <= 100):
        raise ValueError("percent_false_positive_genes must be a string from > 0% to < 100%")
    import time as _time
    t0 = _time.perf_counter()
    p = _prep.prepare(smp, trn, abn, sig, chk, {c: _col(data, c) for c in covs}, formula, how_many_negative_controls,
                      scaling_factor=None if scaling_factor is None else _col(data, scaling_factor))
    t1 = _time.perf_counter()
    n_samples = len(p.samples)            # distinct samples of the (rectangular) table
    if adj_prob_theshold_2 is None:                                                        # :156-160
        adj_prob_theshold_2 = percent_false_positive_genes / 100 / n_samples * (2 if do_check_only_on_detrimental else 1)
    adj_prob_theshold_1 = max(0.05, adj_prob_theshold_2 * 2)                               # :163
    draws_1 = max(draws_after_tail / adj_prob_theshold_1, 1000)                            # :166-167
    draws_2 = max(draws_after_tail / adj_prob_theshold_2, 1000)
    if approximate_posterior_analysis is None:                                             # :170-176
        approximate_posterior_analysis = draws_2 > 20000
    t1b = _time.perf_counter()
    model = NBModel(p.counts, p.X, p.exposure_rate, p.K, lambda_mu_mu=LAMBDA_MU_MU, device=device, devices=devices)
    t2 = _time.perf_counter()
    try:
        res1 = do_inference(model, approximate_posterior_inference=approximate_posterior_inference,
                            approximate_posterior_analysis=False, adj_prob_theshold=adj_prob_theshold_1,
                            how_many_posterior_draws=draws_1, cores=cores, seed=seed, pass_fit=pass_fit)     # :268-286
        K, S = p.K, len(p.samples)
        t3 = _time.perf_counter()
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
    t4 = _time.perf_counter()
    # ---- merge_results + format_results (R/utilities.R:539-608), columnar: one long frame, then views of it --------
    if return_format not in ("nested", "long", "failing"):
        raise ValueError("return_format must be 'nested', 'long' or 'failing'")
    # a row of the long table is a (gene, sample) pair; `sel` = the pairs that are materialised (all, or the failing ones)
    sel = np.flatnonzero(~res2.ppc.reshape(-1)) if return_format == "failing" else np.arange(K * S)
    gi, si = np.divmod(sel, S)
    slope1 = res1.slope if len(res1.slope) else np.full(K, np.nan)
    slope2 = res2.slope if len(res2.slope) else np.full(K, np.nan)
    cols = {"S": si + 1, "G": gi + 1, abundance: p.counts[:K].reshape(-1)[sel],
            sample: np.asarray(p.samples, dtype=object)[si], "slope_before_outlier_filtering": slope1[gi]}
    for c in covs:
        cols[c] = _col(data, c)[p.first_row][si]
    cols.update({"exposure_rate": p.exposure_rate[si], "multiplier": p.multiplier[si],
                 ".lower": res2.lower.reshape(-1)[sel], ".upper": res2.upper.reshape(-1)[sel],
                 "slope_after_outlier_filtering": slope2[gi],
                 "posterior_predictive_check_succeded": res2.ppc.reshape(-1)[sel]})
    if res2.deleterious is not None:
        cols["deleterious_outliers"] = res2.deleterious.reshape(-1)[sel]
    totals = pd.DataFrame({transcript: p.genes[:K], "ppc_samples_failed": res2.ppc_samples_failed.astype(np.int64)})
    if do_check_only_on_detrimental:
        totals["tot_deleterious_outliers"] = res2.tot_deleterious_outliers.astype(np.int64)
    if return_format == "nested":
        long = pd.DataFrame(cols)
        out = pd.DataFrame({transcript: p.genes[:K],
                            "sample_wise_data": [long.iloc[g * S:(g + 1) * S].reset_index(drop=True) for g in range(K)],
                            "ppc_samples_failed": totals["ppc_samples_failed"].to_numpy()})
        if do_check_only_on_detrimental:
            out["tot_deleterious_outliers"] = totals["tot_deleterious_outliers"].to_numpy()
    else:
        out = pd.DataFrame(cols)
        out.insert(0, transcript, np.asarray(p.genes[:K], dtype=object)[out["G"].to_numpy() - 1])
        out.attrs["gene_totals"] = totals
    if timings is not None:
        timings.update({"prep_s": t1 - t0, "upload_s": t2 - t1b, "pass1_s": t3 - t2, "pass2_s": t4 - t3,
                        "result_s": _time.perf_counter() - t4, "pass1_info": res1.info, "pass2_info": res2.info})
    out.attrs.update({"total_draws": res2.total_draws, "transcript_column": transcript, "abundance_column": abundance,
                      "sample_column": sample, "formula": formula, "fit 1 info": res1.info, "fit 2 info": res2.info,
                      "seed": seed})
    if pass_fit:
        out.attrs.update({"fit 1": res1.fit, "fit 2": res2.fit, "model": model})
    return out
