"""TEST INFRASTRUCTURE (oracle): NumPy statement of the path's input edge, the checker of the native
ppcseq_prep_table / ppcseq_tmm_factors (ppcseq_b200/csrc/prep_host.cu).  Never imported by the product.

  select_to_check_and_house_keeping   R/utilities.R:628-649
  format_input                        R/utilities.R:924-959   (G = first appearance, checked genes first; S likewise)
  get_scaled_counts_bulk + calcNormFactor  R/tidybulk.R:150-241, :262-323 (edgeR TMM on the SELECTED genes)

edgeR is not in the reference tree (Bioconductor dependency, `edgeR::calcNormFactors`, called at R/tidybulk.R:294-304):
`tmm_norm_factors` restates its published algorithm (Robinson & Oshlack 2010; edgeR 3.x `.calcFactorTMM`:
logratioTrim = 0.3, sumTrim = 0.05, doWeighting, Acutoff = -1e10, factors scaled to geometric mean 1) with explicit
average ranks.  Pinned only by the row-by-row restatement in tests/test_prep.py and the bundled fixture: edgeR itself
cannot run here ("parity unpinned" for the TMM factors, see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np


def factorize(values):
    """(codes, uniques), uniques in order of first appearance -- dplyr's distinct() / `factor(...) %>% as.integer`
    on a table arranged by first appearance (R/utilities.R:949-958)."""
    import pandas as pd
    codes, uniques = pd.factorize(np.asarray(values, dtype=object) if isinstance(values, list) else np.asarray(values))
    return codes.astype(np.int64), np.asarray(uniques)


def rank_average(x: np.ndarray) -> np.ndarray:
    """R's rank(ties.method = "average"), 1-based."""
    from scipy.stats import rankdata
    return rankdata(x, method="average")


def calc_factor_tmm(obs, ref, logratio_trim=0.3, sum_trim=0.05, a_cutoff=-1e10):
    obs = obs.astype(np.float64)
    ref = ref.astype(np.float64)
    nO, nR = obs.sum(), ref.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        logR = np.log2((obs / nO) / (ref / nR))
        absE = (np.log2(obs / nO) + np.log2(ref / nR)) / 2.0
        v = (nO - obs) / nO / obs + (nR - ref) / nR / ref
    fin = np.isfinite(logR) & np.isfinite(absE) & (absE > a_cutoff)
    logR, absE, v = logR[fin], absE[fin], v[fin]
    if len(logR) == 0 or np.max(np.abs(logR)) < 1e-6:
        return 1.0
    n = len(logR)
    loL = np.floor(n * logratio_trim) + 1
    hiL = n + 1 - loL
    loS = np.floor(n * sum_trim) + 1
    hiS = n + 1 - loS
    rL, rS = rank_average(logR), rank_average(absE)
    keep = (rL >= loL) & (rL <= hiL) & (rS >= loS) & (rS <= hiS)
    with np.errstate(divide="ignore", invalid="ignore"):
        den = np.sum(1.0 / v[keep])
        f = np.sum(logR[keep] / v[keep]) / den if den > 0 else np.nan
    if np.isnan(f):
        f = 0.0
    return float(2.0 ** f)


def tmm_norm_factors(mat: np.ndarray, ref_column: int) -> np.ndarray:
    """edgeR::calcNormFactors(method = "TMM") on a genes x samples matrix; lib.size = column sums."""
    x = np.asarray(mat, dtype=np.float64)
    x = x[(x > 0).sum(axis=1) > 0]                      # drop all-zero rows
    ref = np.ascontiguousarray(x[:, ref_column])
    f = np.array([calc_factor_tmm(x[:, j], ref) for j in range(x.shape[1])])
    return f / np.exp(np.mean(np.log(f)))


def reference_column(mat: np.ndarray) -> int:
    """first sample (factor level order) whose median count is the maximum (R/tidybulk.R:262-291)."""
    med = np.median(np.asarray(mat, dtype=np.float64), axis=0)
    return int(np.argmin(np.abs(med - med.max())))


def prepare_table(sample, transcript, abundance, significance, do_check, how_many_negative_controls: int = 500):
    """-> dict(counts int32 [G, S], genes [G], samples [S], K, first_row [S]); all arguments row-aligned columns."""
    import pandas as pd
    abundance = np.asarray(abundance)
    significance = np.asarray(significance, dtype=np.float64)
    do_check = np.asarray(do_check, dtype=bool)
    if not do_check.any():
        raise ValueError("no transcripts with the category .do_check")
    t_code, t_names = factorize(transcript)
    s_code, s_names = factorize(sample)
    # --- select_to_check_and_house_keeping (R/utilities.R:628-649)
    order = np.argsort(significance, kind="stable")                         # arrange(significance)
    distinct_sorted = pd.unique(t_code[order])                              # distinct(transcript): first appearance
    in_tail = np.zeros(len(t_names), bool)
    if how_many_negative_controls > 0:
        in_tail[distinct_sorted[-how_many_negative_controls:]] = True
    rows = np.concatenate([np.flatnonzero(do_check), np.flatnonzero(~do_check & in_tail[t_code])])
    # --- format_input: G and S by first appearance (R/utilities.R:924-959)
    gidx, g_first = factorize(t_code[rows])
    sidx, s_first = factorize(s_code[rows])
    G, S = len(g_first), len(s_first)
    K = int(len(pd.unique(t_code[do_check])))
    if len(rows) > G * S:
        raise ValueError("the input has duplicated (transcript, sample) rows")
    counts = np.full((G, S), -1, dtype=np.int64)
    counts[gidx, sidx] = abundance[rows]
    if (counts < 0).any():
        raise ValueError("the input is not rectangular (every gene needs every sample)")   # R/utilities.R:1360
    run_max = np.maximum.accumulate(sidx)
    first_in_rows = np.flatnonzero(np.concatenate([[True], sidx[1:] > run_max[:-1]]))
    return dict(counts=counts.astype(np.int32), genes=[t_names[i] for i in g_first], samples=[s_names[i] for i in s_first],
                K=K, first_row=rows[first_in_rows])
