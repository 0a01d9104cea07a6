"""Input side of the hot path: tidy table -> (gene selection, G/S indexing, dense int32 counts, design
matrix, TMM exposure).  Host-side mirror of the reference's one-off preprocessing:

  select_to_check_and_house_keeping   R/utilities.R:628-649
  format_input                        R/utilities.R:924-959   (G = first appearance, checked genes first; S likewise)
  create_design_matrix / parse_formula  R/utilities.R:887-900, :220-225   (model.matrix, treatment contrasts)
  get_scaled_counts_bulk + calcNormFactor  R/tidybulk.R:150-241, :262-323 (edgeR TMM on the SELECTED genes)
  exposure_rate = -log(multiplier)    R/methods.R:222-238

edgeR is not in the reference tree (a Bioconductor dependency, `edgeR::calcNormFactors`, called at
R/tidybulk.R:294-304); `tmm_norm_factors` restates its published TMM algorithm (Robinson & Oshlack 2010;
edgeR 3.x `.calcFactorTMM`: logratioTrim = 0.3, sumTrim = 0.05, doWeighting, Acutoff = -1e10, factors scaled
to geometric mean 1).  This is S-vector / one-off work: it stays on the host by design.
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np


@dataclass
class Prepared:
    counts: np.ndarray          # int32 [G, S] gene-major, G / S in the reference's index order
    X: np.ndarray               # float64 [S, C]
    exposure_rate: np.ndarray   # float64 [S]
    multiplier: np.ndarray      # float64 [S]
    K: int                      # how_many_to_check: genes 0..K-1 are the checked ones
    genes: list                 # [G] names in G order
    samples: list               # [S] names in S order
    design_columns: list        # [C] model.matrix column names
    tmm: np.ndarray             # float64 [S] TMM factors (S order)
    reference_sample: str


def _first_appearance(values):
    seen, order = {}, []
    for v in values:
        if v not in seen:
            seen[v] = len(order)
            order.append(v)
    return order, seen


def _rank_average(x: np.ndarray) -> np.ndarray:
    """R's rank(ties.method = "average"), 1-based."""
    order = np.argsort(x, kind="mergesort")
    xs = x[order]
    n = len(x)
    ranks = np.empty(n)
    i = 0
    while i < n:
        j = i
        while j + 1 < n and xs[j + 1] == xs[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def _calc_factor_tmm(obs, ref, logratio_trim=0.3, sum_trim=0.05, a_cutoff=-1e10):
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
    rL, rS = _rank_average(logR), _rank_average(absE)
    keep = (rL >= loL) & (rL <= hiL) & (rS >= loS) & (rS <= hiS)
    den = np.sum(1.0 / v[keep])
    f = np.sum(logR[keep] / v[keep]) / den if den > 0 else np.nan
    if np.isnan(f):
        f = 0.0
    return float(2.0 ** f)


def tmm_norm_factors(mat: np.ndarray, ref_column: int) -> np.ndarray:
    """edgeR::calcNormFactors(method = "TMM") on a genes x samples matrix; lib.size = column sums."""
    x = np.asarray(mat, dtype=np.float64)
    x = x[(x > 0).sum(axis=1) > 0]                      # drop all-zero rows
    f = np.array([_calc_factor_tmm(x[:, j], x[:, ref_column]) for j in range(x.shape[1])])
    return f / np.exp(np.mean(np.log(f)))


def parse_formula(formula: str) -> list:
    """Covariate names of a one-sided additive formula ('~ Label + batch'); '~ 1' -> []."""
    rhs = formula.split("~", 1)[1]
    terms = [t.strip() for t in rhs.split("+")]
    for t in terms:
        if not re.fullmatch(r"[A-Za-z_.][A-Za-z0-9_.]*|1|0", t):
            raise ValueError(f"unsupported formula term {t!r}: only additive main effects are mirrored")
    return [t for t in terms if t not in ("1", "0")]


def model_matrix(formula: str, columns: dict, n: int):
    """model.matrix for additive main effects: strings are factors (levels sorted, treatment contrasts),
    numbers enter as they are.  Returns (X [n, C], column names)."""
    cols, names = [np.ones(n)], ["(Intercept)"]
    for name in parse_formula(formula):
        v = columns[name]
        if isinstance(v, np.ndarray) and v.dtype.kind in "fiu":
            cols.append(v.astype(np.float64))
            names.append(name)
        else:
            levels = sorted(set(v))
            for lv in levels[1:]:
                cols.append(np.array([1.0 if x == lv else 0.0 for x in v]))
                names.append(f"{name}{lv}")
    return np.stack(cols, axis=1), names


def prepare(sample, transcript, abundance, significance, do_check, covariates: dict, formula: str,
            how_many_negative_controls: int = 500, scaling_factor=None) -> Prepared:
    """All arguments are row-aligned columns of the tidy input table (R/methods.R:74-98)."""
    sample = list(sample)
    transcript = list(transcript)
    abundance = np.asarray(abundance)
    if abundance.dtype.kind not in "iu":
        raise ValueError("the abundance column must be of class integer")          # R/methods.R:139-148
    significance = np.asarray(significance, dtype=np.float64)
    do_check = np.asarray(do_check, dtype=bool)
    n = len(sample)
    if not do_check.any():
        raise ValueError("no transcripts with the category .do_check")
    # --- select_to_check_and_house_keeping -------------------------------------------------------
    order = np.argsort(significance, kind="mergesort")                      # arrange(significance), stable
    distinct_sorted, _ = _first_appearance([transcript[i] for i in order])
    tail = set(distinct_sorted[-how_many_negative_controls:]) if how_many_negative_controls > 0 else set()
    rows_check = [i for i in range(n) if do_check[i]]
    rows_ctrl = [i for i in range(n) if not do_check[i] and transcript[i] in tail]
    rows = rows_check + rows_ctrl
    # --- format_input: G and S by first appearance -----------------------------------------------
    genes, gidx = _first_appearance([transcript[i] for i in rows])
    samples, sidx = _first_appearance([sample[i] for i in rows])
    G, S = len(genes), len(samples)
    K = len({transcript[i] for i in rows_check})
    counts = np.full((G, S), -1, dtype=np.int64)
    for i in rows:
        counts[gidx[transcript[i]], sidx[sample[i]]] = abundance[i]
    if (counts < 0).any():
        raise ValueError("the input is not rectangular (every gene needs every sample)")   # R/utilities.R:1360
    counts = counts.astype(np.int32)
    # --- create_design_matrix: distinct(sample, covariates) arranged by sample --------------------
    cov_names = parse_formula(formula)
    first_row = {}
    for i in rows:
        first_row.setdefault(sample[i], i)
    sorted_samples = sorted(samples)
    cov_cols = {}
    for name in cov_names:
        v = covariates[name]
        vals = [v[first_row[s]] for s in sorted_samples]
        cov_cols[name] = np.asarray(vals) if isinstance(v, np.ndarray) and v.dtype.kind in "fiu" else vals
    X_sorted, colnames = model_matrix(formula, cov_cols, S)
    # The reference indexes X rows by the S index although model.matrix is in sorted-sample order
    # (R/utilities.R:887-900 vs :955-958); the two orders coincide whenever samples first appear sorted.
    X = X_sorted
    # --- exposure: TMM on the selected genes (R/methods.R:222-238) --------------------------------
    if scaling_factor is None:
        pos = [sidx[s] for s in sorted_samples]                   # factor(sample): sorted levels
        mat = counts[:, pos].astype(np.float64)                   # genes x samples(sorted)
        med = np.median(mat, axis=0)
        ref = int(np.argmin(np.abs(med - med.max())))             # first sample whose median is the maximum
        nf = tmm_norm_factors(mat, ref)
        tot = mat.sum(axis=0)
        mult_sorted = 1.0 / (tot * nf) * tot[ref]
        multiplier = np.empty(S)
        tmm = np.empty(S)
        for j, s in enumerate(sorted_samples):
            multiplier[sidx[s]] = mult_sorted[j]
            tmm[sidx[s]] = nf[j]
        ref_name = sorted_samples[ref]
    else:
        sf = np.asarray(scaling_factor, dtype=np.float64)
        multiplier = np.array([sf[first_row[s]] for s in samples])
        tmm = np.ones(S)
        ref_name = ""
    exposure_rate = -np.log(multiplier)
    return Prepared(counts, X, exposure_rate, multiplier, K, genes, samples, colnames, tmm, ref_name)
