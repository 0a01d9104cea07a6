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
    first_row: np.ndarray = None  # int64 [S]: a row of the input table that belongs to sample S (its covariate values)


def _factorize(values):
    """(codes, uniques) with uniques in order of first appearance -- dplyr's distinct() / the reference's
    `mutate(G = factor(...) %>% as.integer)` on a table already arranged by first appearance (R/utilities.R:949-958).
    Vectorised (hash based): no per-row Python."""
    import pandas as pd
    codes, uniques = pd.factorize(np.asarray(values, dtype=object) if isinstance(values, list) else np.asarray(values))
    return codes.astype(np.int64), np.asarray(uniques)


def _rank_average(x: np.ndarray) -> np.ndarray:
    """R's rank(ties.method = "average"), 1-based."""
    from scipy.stats import rankdata
    return rankdata(x, method="average")


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
    """edgeR::calcNormFactors(method = "TMM") on a genes x samples matrix; lib.size = column sums.  The samples are
    independent given the reference column: they are spread over the host cores (the two rank computations per sample
    are sorts, which release the GIL)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    x = np.asarray(mat, dtype=np.float64)
    x = x[(x > 0).sum(axis=1) > 0]                      # drop all-zero rows
    ref = np.ascontiguousarray(x[:, ref_column])
    cols = range(x.shape[1])
    if x.shape[1] >= 8 and x.shape[0] >= 2000:
        with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
            f = np.array(list(ex.map(lambda j: _calc_factor_tmm(x[:, j], ref), cols)))
    else:
        f = np.array([_calc_factor_tmm(x[:, j], ref) for j in cols])
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
    """All arguments are row-aligned columns of the tidy input table (R/methods.R:74-98).  Everything that touches
    the rows of the table is vectorised (factorisation, stable sorts, fancy indexing): 3e8 rows (config 5) take
    seconds, not hours; only S-length work (design matrix levels) is plain Python."""
    abundance = np.asarray(abundance)
    if abundance.dtype.kind not in "iu":
        raise ValueError("the abundance column must be of class integer")          # R/methods.R:139-148
    significance = np.asarray(significance, dtype=np.float64)
    do_check = np.asarray(do_check, dtype=bool)
    n = len(abundance)
    if not do_check.any():
        raise ValueError("no transcripts with the category .do_check")
    t_code, t_names = _factorize(transcript)
    s_code, s_names = _factorize(sample)
    # --- select_to_check_and_house_keeping (R/utilities.R:628-649) ---------------------------------
    order = np.argsort(significance, kind="stable")                         # arrange(significance)
    import pandas as pd
    distinct_sorted = pd.unique(t_code[order])                              # distinct(transcript): first appearance (hash)
    in_tail = np.zeros(len(t_names), bool)
    if how_many_negative_controls > 0:
        in_tail[distinct_sorted[-how_many_negative_controls:]] = True
    rows = np.concatenate([np.flatnonzero(do_check), np.flatnonzero(~do_check & in_tail[t_code])])
    # --- format_input: G and S by first appearance (R/utilities.R:924-959) -------------------------
    gidx, g_first = _factorize(t_code[rows])
    sidx, s_first = _factorize(s_code[rows])
    genes = [t_names[i] for i in g_first]
    samples = [s_names[i] for i in s_first]
    G, S = len(genes), len(samples)
    K = int(len(pd.unique(t_code[do_check])))
    counts = np.full((G, S), -1, dtype=np.int64)
    counts[gidx, sidx] = abundance[rows]
    if (counts < 0).any():
        raise ValueError("the input is not rectangular (every gene needs every sample)")   # R/utilities.R:1360
    counts = counts.astype(np.int32)
    # --- create_design_matrix: distinct(sample, covariates) arranged by sample (R/utilities.R:887-900) --------
    cov_names = parse_formula(formula)
    # first row (in `rows` order) of S index 0..S-1: the codes are numbered by first appearance, so a first occurrence is
    # where the code exceeds everything before it (O(n), no sort)
    run_max = np.maximum.accumulate(sidx)
    first_in_rows = np.flatnonzero(np.concatenate([[True], sidx[1:] > run_max[:-1]]))
    first_row = rows[first_in_rows]                                         # [S], indexed by S index
    sorted_pos = sorted(range(S), key=lambda j: samples[j])                 # S index of the j-th sample in sorted order
    sorted_samples = [samples[j] for j in sorted_pos]
    if sorted_samples != samples:
        import warnings
        warnings.warn("samples do not first appear in sorted order: the reference pairs the rows of model.matrix "
                      "(sorted by sample, R/utilities.R:887-900) with the S index (first appearance, :955-958), so "
                      "sample j is modelled with the covariates of the j-th SORTED sample; this behaviour is reproduced")
    cov_cols = {}
    for name in cov_names:
        v = covariates[name]
        if isinstance(v, np.ndarray) and v.dtype.kind in "fiu":
            cov_cols[name] = v[first_row[sorted_pos]]
        else:
            va = np.asarray(v, dtype=object)
            cov_cols[name] = list(va[first_row[sorted_pos]])
    X_sorted, colnames = model_matrix(formula, cov_cols, S)
    # The reference indexes X rows by the S index although model.matrix is in sorted-sample order
    # (R/utilities.R:887-900 vs :955-958); the two orders coincide whenever samples first appear sorted.
    X = X_sorted
    # --- exposure: TMM on the selected genes (R/methods.R:222-238) --------------------------------
    if scaling_factor is None:
        mat = counts[:, sorted_pos].astype(np.float64)            # genes x samples(sorted): factor(sample) levels
        med = np.median(mat, axis=0)
        ref = int(np.argmin(np.abs(med - med.max())))             # first sample whose median is the maximum
        nf = tmm_norm_factors(mat, ref)
        tot = mat.sum(axis=0)
        mult_sorted = 1.0 / (tot * nf) * tot[ref]
        multiplier = np.empty(S)
        tmm = np.empty(S)
        multiplier[sorted_pos] = mult_sorted
        tmm[sorted_pos] = nf
        ref_name = sorted_samples[ref]
    else:
        sf = np.asarray(scaling_factor, dtype=np.float64)
        multiplier = sf[first_row]
        tmm = np.ones(S)
        ref_name = ""
    exposure_rate = -np.log(multiplier)
    return Prepared(counts, X, exposure_rate, multiplier, K, genes, samples, colnames, tmm, ref_name, first_row)
