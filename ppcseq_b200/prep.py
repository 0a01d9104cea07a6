"""Input side of the hot path: tidy table -> (gene selection, G/S indexing, dense int32 counts, design
matrix, TMM exposure).  Host-side mirror of the reference's one-off preprocessing:

  select_to_check_and_house_keeping   R/utilities.R:628-649
  format_input                        R/utilities.R:924-959   (G = first appearance, checked genes first; S likewise)
  create_design_matrix / parse_formula  R/utilities.R:887-900, :220-225   (model.matrix, treatment contrasts)
  get_scaled_counts_bulk + calcNormFactor  R/tidybulk.R:150-241, :262-323 (edgeR TMM on the SELECTED genes)
  exposure_rate = -log(multiplier)    R/methods.R:222-238

The O(rows) part (selection, indexing, dense scatter) and TMM are native host code in the library
(csrc/prep_host.cu: ppcseq_prep_table, ppcseq_tmm_factors -- an R caller binds the same two entry points); this module
passes the columns, and does the S-length work (design matrix).  edgeR is not in the reference tree (a Bioconductor
dependency, `edgeR::calcNormFactors`, called at R/tidybulk.R:294-304): its published TMM algorithm is restated there.
One-off work: it stays on the host by design.
"""
from __future__ import annotations

import ctypes
import re
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int32_p, c_uint8_p


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


def _ids(values):
    """A column of the table as int64 ids for the native pass: integer columns go as they are, anything else through
    pandas' hash factorisation (what an R caller holds already: factor codes).  -> (ids, names or None)."""
    a = values if isinstance(values, np.ndarray) else np.asarray(values)
    if a.dtype.kind in "iu" and not (a.dtype.kind == "u" and a.dtype.itemsize == 8):
        return np.ascontiguousarray(a, dtype=np.int64), None
    import pandas as pd
    codes, uniques = pd.factorize(a)
    if len(codes) and codes.min() < 0:
        raise ValueError("missing values in an id column")
    return np.ascontiguousarray(codes, dtype=np.int64), np.asarray(uniques)


def _i64p(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def prepare_table(sample, transcript, abundance, significance, do_check, how_many_negative_controls: int = 500,
                  threads: int = 0):
    """Native ppcseq_prep_table: -> (counts int32 [G, S], genes [G], samples [S], K, first_row int64 [S])."""
    L = _lib.lib()
    abundance = abundance if isinstance(abundance, np.ndarray) else np.asarray(abundance)
    if abundance.dtype.kind not in "iu":
        raise ValueError("the abundance column must be of class integer")          # R/methods.R:139-148
    if abundance.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        abundance = abundance.astype(np.int64)
    abundance = np.ascontiguousarray(abundance)
    significance = np.ascontiguousarray(significance, dtype=np.float64)
    chk = np.ascontiguousarray(do_check, dtype=bool)
    n = len(abundance)
    if not chk.any():
        raise ValueError("no transcripts with the category .do_check")
    t_ids, t_names = _ids(transcript)
    s_ids, s_names = _ids(sample)
    if not (len(t_ids) == len(s_ids) == len(significance) == len(chk) == n):
        raise ValueError("the columns of the table have different lengths")
    h = ctypes.c_void_p()
    try:
        _lib.check(L.ppcseq_prep_table(n, _i64p(t_ids), _i64p(s_ids), abundance.ctypes.data_as(ctypes.c_void_p),
                                       abundance.dtype.itemsize, significance.ctypes.data_as(c_double_p),
                                       chk.view(np.uint8).ctypes.data_as(c_uint8_p), int(how_many_negative_controls),
                                       int(threads), ctypes.byref(h)))
    except _lib.PpcseqError as e:
        if e.rc == 1:                                                               # PPCSEQ_EINVAL: a property of the table
            raise ValueError(str(e)) from None
        raise
    try:
        G, S, K = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.check(L.ppcseq_prep_dims(h, ctypes.byref(G), ctypes.byref(S), ctypes.byref(K)))
        G, S, K = G.value, S.value, K.value
        g_ids, s_idv, first_row = np.empty(G, np.int64), np.empty(S, np.int64), np.empty(S, np.int64)
        counts = np.empty((G, S), np.int32)
        _lib.check(L.ppcseq_prep_fetch(h, _i64p(g_ids), _i64p(s_idv), _i64p(first_row), counts.ctypes.data_as(c_int32_p)))
    finally:
        L.ppcseq_prep_free(h)
    genes = list(g_ids if t_names is None else t_names[g_ids])
    samples = list(s_idv if s_names is None else s_names[s_idv])
    return counts, genes, samples, K, first_row


def tmm_factors(counts: np.ndarray, order=None, ref_column: int = -1, threads: int = 0):
    """Native ppcseq_tmm_factors on dense int32 counts [G, S]: (factors [S], lib_size [S], ref), all in `order`
    (factor(sample) level order; None = column order).  ref_column < 0: first level with the largest median."""
    c = np.ascontiguousarray(counts, dtype=np.int32)
    G, S = c.shape
    o = None if order is None else np.ascontiguousarray(order, dtype=np.int32)
    f, tot, ref = np.empty(S), np.empty(S), ctypes.c_int32()
    _lib.check(_lib.lib().ppcseq_tmm_factors(G, S, c.ctypes.data_as(c_int32_p), None if o is None else o.ctypes.data_as(c_int32_p),
                                             int(ref_column), int(threads), f.ctypes.data_as(c_double_p),
                                             tot.ctypes.data_as(c_double_p), ctypes.byref(ref)))
    return f, tot, ref.value


def tmm_norm_factors(mat: np.ndarray, ref_column: int) -> np.ndarray:
    """edgeR::calcNormFactors(method = "TMM") on a genes x samples matrix of counts; lib.size = column sums."""
    m = np.asarray(mat)
    ci = m.astype(np.int32)
    if not np.array_equal(ci, m):
        raise ValueError("tmm_norm_factors takes integer counts below 2^31")
    return tmm_factors(ci, None, int(ref_column))[0]


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
            how_many_negative_controls: int = 500, scaling_factor=None, threads: int = 0) -> Prepared:
    """All arguments are row-aligned columns of the tidy input table (R/methods.R:74-98).  Everything that touches the
    rows of the table runs in the native library (ppcseq_prep_table, ppcseq_tmm_factors: threaded passes over the
    columns, 3e8 rows in seconds); only S-length work (design matrix levels) is plain Python."""
    counts, genes, samples, K, first_row = prepare_table(sample, transcript, abundance, significance, do_check,
                                                         how_many_negative_controls, threads)
    G, S = counts.shape
    # --- create_design_matrix: distinct(sample, covariates) arranged by sample (R/utilities.R:887-900) --------
    cov_names = parse_formula(formula)
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
        pick = first_row[sorted_pos]                                        # S rows of the column, never the column itself
        if isinstance(v, np.ndarray):
            cov_cols[name] = v[pick] if v.dtype.kind in "fiu" else list(v[pick])
        elif hasattr(v, "iloc"):
            cov_cols[name] = list(v.iloc[pick])
        else:
            cov_cols[name] = [v[i] for i in pick]
    X_sorted, colnames = model_matrix(formula, cov_cols, S)
    # The reference indexes X rows by the S index although model.matrix is in sorted-sample order
    # (R/utilities.R:887-900 vs :955-958); the two orders coincide whenever samples first appear sorted.
    X = X_sorted
    # --- exposure: TMM on the selected genes (R/methods.R:222-238) --------------------------------
    if scaling_factor is None:
        # levels of factor(sample) = sorted samples; reference = first level whose median is the maximum
        nf, tot, ref = tmm_factors(counts, sorted_pos, -1, threads)
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
