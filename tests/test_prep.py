"""CPU: the input preparation (ppcseq_b200/prep.py over the native ppcseq_prep_table / ppcseq_tmm_factors of the library,
csrc/prep_host.cu: tidy table -> gene selection, G / S indexing, dense counts, design matrix, TMM exposure; reference
R/utilities.R:628-649, :924-959, :887-900, R/tidybulk.R:150-323) against (a) a row-by-row restatement with plain Python
loops (round 1's implementation, kept here as the checker) on shuffled, unsorted, duplicated-significance tables and
(b) the NumPy statement oracle/prep_np.py at 10^6 rows, where per-row Python would take minutes; TMM by selection against
TMM by explicit average ranks on tie-heavy counts; thread-count independence; the error cases."""
import time
import warnings

import numpy as np
import pytest

from ppcseq_b200 import prep
from oracle.prep_np import calc_factor_tmm as _calc_factor_tmm
from ppcseq_b200.prep import Prepared, model_matrix, parse_formula


def _first_appearance(values):
    seen, order = {}, []
    for v in values:
        if v not in seen:
            seen[v] = len(order)
            order.append(v)
    return order, seen


def _rank_average_loops(x: np.ndarray) -> np.ndarray:
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



def _tmm_loops(mat, ref_column):
    x = np.asarray(mat, dtype=np.float64)
    x = x[(x > 0).sum(axis=1) > 0]
    f = np.array([_calc_factor_tmm(x[:, j], x[:, ref_column]) for j in range(x.shape[1])])
    return f / np.exp(np.mean(np.log(f)))


def _prepare_loops(sample, transcript, abundance, significance, do_check, covariates: dict, formula: str,
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
        nf = _tmm_loops(mat, ref)
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


def _table(G, S, n_check, seed, shuffle, sorted_samples=True, numeric_cov=False):
    rng = np.random.default_rng(seed)
    genes = [f"gene{int(i):05d}" for i in rng.permutation(G)]
    samples = [f"s{int(i):04d}" for i in (np.arange(S) if sorted_samples else rng.permutation(S))]
    mean = np.exp(rng.uniform(1, 8, G))[:, None] * np.exp(rng.normal(0, 0.3, S))[None, :]
    counts = rng.poisson(mean).astype(np.int64)
    pval = np.round(rng.uniform(0, 1, G), 2)                         # many ties in the significance column
    check = np.zeros(G, bool)
    check[rng.choice(G, n_check, replace=False)] = True
    label = {s: ("A", "B", "C")[j % 3] for j, s in enumerate(samples)}
    age = {s: float(rng.normal(50, 10)) for s in samples}
    rows = [(g, s) for g in range(G) for s in range(S)]
    if shuffle:
        rows = [rows[i] for i in rng.permutation(len(rows))]
    return dict(
        sample=[samples[s] for _, s in rows], transcript=[genes[g] for g, _ in rows],
        abundance=np.array([counts[g, s] for g, s in rows], dtype=np.int64),
        significance=np.array([pval[g] for g, _ in rows]), do_check=np.array([check[g] for g, _ in rows]),
        covariates={"Label": [label[samples[s]] for _, s in rows],
                    "age": np.array([age[samples[s]] for _, s in rows])},
        formula="~ Label + age" if numeric_cov else "~ Label")


@pytest.mark.parametrize("shuffle,sorted_samples,numeric_cov,nctrl", [(False, True, False, 20), (True, True, True, 20),
                                                                      (True, False, False, 7), (False, False, True, 0)])
def test_vectorised_prepare_equals_the_row_loops(shuffle, sorted_samples, numeric_cov, nctrl):
    t = _table(60, 9, 11, seed=3, shuffle=shuffle, sorted_samples=sorted_samples, numeric_cov=numeric_cov)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = prep.prepare(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], t["covariates"],
                         t["formula"], nctrl)
    b = _prepare_loops(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], t["covariates"],
                       t["formula"], nctrl)
    assert a.K == b.K and list(a.genes) == list(b.genes) and list(a.samples) == list(b.samples)
    assert a.design_columns == b.design_columns and a.reference_sample == b.reference_sample
    assert np.array_equal(a.counts, b.counts) and a.counts.dtype == np.int32
    assert np.array_equal(a.X, b.X)
    for x, y in ((a.exposure_rate, b.exposure_rate), (a.multiplier, b.multiplier), (a.tmm, b.tmm)):
        assert np.allclose(x, y, rtol=1e-13, atol=0)


def test_scaling_factor_bypass_and_unsorted_warning():
    t = _table(30, 6, 5, seed=8, shuffle=True, sorted_samples=False)
    sf_by_sample = {s: 0.5 + 0.1 * j for j, s in enumerate(sorted(set(t["sample"])))}
    sf = np.array([sf_by_sample[s] for s in t["sample"]])
    with pytest.warns(UserWarning, match="sorted order"):
        a = prep.prepare(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], t["covariates"],
                         t["formula"], 10, scaling_factor=sf)
    b = _prepare_loops(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], t["covariates"],
                       t["formula"], 10, scaling_factor=sf)
    assert np.array_equal(a.multiplier, b.multiplier) and np.array_equal(a.X, b.X) and np.array_equal(a.counts, b.counts)
    assert np.allclose(a.multiplier, [sf_by_sample[s] for s in a.samples])


def test_non_rectangular_and_non_integer_inputs_are_rejected():
    t = _table(12, 4, 3, seed=1, shuffle=False)
    keep = np.ones(len(t["sample"]), bool)
    keep[5] = False
    with pytest.raises(ValueError, match="rectangular"):
        prep.prepare([s for s, k in zip(t["sample"], keep) if k], [g for g, k in zip(t["transcript"], keep) if k],
                     t["abundance"][keep], t["significance"][keep], t["do_check"][keep],
                     {"Label": [v for v, k in zip(t["covariates"]["Label"], keep) if k]}, "~ Label", 5)
    with pytest.raises(ValueError, match="integer"):
        prep.prepare(t["sample"], t["transcript"], t["abundance"].astype(float), t["significance"], t["do_check"],
                     t["covariates"], "~ Label", 5)


def test_prepare_at_scale_is_vectorised():
    """2,000 genes x 500 samples = 1e6 rows, every gene checked, TMM bypassed by a scaling-factor column (TMM is
    S-vector work timed separately): the row-loop implementation needs ~10 s here, the vectorised one well under 2."""
    G, S = 2000, 500
    rng = np.random.default_rng(0)
    genes = np.repeat(np.array([f"g{i:05d}" for i in range(G)]), S)
    samples = np.tile(np.array([f"s{j:04d}" for j in range(S)]), G)
    ab = rng.poisson(200.0, G * S).astype(np.int32)
    t0 = time.perf_counter()
    p = prep.prepare(samples, genes, ab, np.repeat(rng.uniform(0, 1, G), S), np.ones(G * S, bool),
                     {"Label": np.tile(np.where(np.arange(S) % 2 == 0, "A", "B"), G)}, "~ Label", 0,
                     scaling_factor=np.ones(G * S))
    dt = time.perf_counter() - t0
    assert p.counts.shape == (G, S) and p.K == G and np.array_equal(p.counts.reshape(-1), ab)
    assert p.X.shape == (S, 2) and dt < 5.0, dt


# ---- native against the NumPy statement (oracle/prep_np.py) -----------------------------------------------------------
def _big_table(G, S, n_check, seed, n_mixed=0):
    """integer ids, rows shuffled, significance with ties, do_check per gene (+ n_mixed genes whose rows disagree)."""
    rng = np.random.default_rng(seed)
    gid = rng.permutation(10 * G)[:G].astype(np.int64) - 3 * G            # arbitrary, also negative, ids
    sid = rng.permutation(5 * S)[:S].astype(np.int64)
    counts = rng.poisson(np.exp(rng.uniform(0, 7, G))[:, None] * np.exp(rng.normal(0, 0.4, S))[None, :]).astype(np.int64)
    pval = np.round(rng.uniform(0, 1, G), 3)
    check = np.zeros(G, bool)
    check[rng.choice(G, n_check, replace=False)] = True
    g, s = np.divmod(rng.permutation(G * S), S)
    chk_rows = check[g].copy()
    if n_mixed:
        for gg in rng.choice(np.flatnonzero(check), n_mixed, replace=False):
            r = np.flatnonzero(g == gg)
            chk_rows[r[::2]] = False
    return dict(sample=sid[s], transcript=gid[g], abundance=counts[g, s], significance=pval[g], do_check=chk_rows)


@pytest.mark.parametrize("G,S,n_check,nctrl,n_mixed,threads", [(2000, 500, 300, 400, 0, 0), (2000, 500, 2000, 0, 0, 3),
                                                               (700, 41, 50, 10000, 5, 0), (300, 7, 1, 37, 0, 1)])
def test_native_table_pass_equals_the_numpy_statement(G, S, n_check, nctrl, n_mixed, threads):
    from oracle import prep_np
    t = _big_table(G, S, n_check, seed=G + S, n_mixed=n_mixed)
    ref = prep_np.prepare_table(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], nctrl)
    counts, genes, samples, K, first_row = prep.prepare_table(t["sample"], t["transcript"], t["abundance"],
                                                              t["significance"], t["do_check"], nctrl, threads=threads)
    assert K == ref["K"] and list(genes) == list(ref["genes"]) and list(samples) == list(ref["samples"])
    assert counts.dtype == np.int32 and np.array_equal(counts, ref["counts"])
    assert np.array_equal(first_row, ref["first_row"])
    # int32 abundance and string ids take the same route
    c2, g2, s2, K2, _ = prep.prepare_table(np.array([f"s{v}" for v in t["sample"]]), np.array([f"g{v}" for v in t["transcript"]]),
                                           t["abundance"].astype(np.int32), t["significance"], t["do_check"], nctrl, threads=2)
    assert np.array_equal(c2, counts) and K2 == K and g2 == [f"g{v}" for v in genes] and s2 == [f"s{v}" for v in samples]


def test_native_table_pass_thread_count_does_not_matter():
    t = _big_table(1500, 300, 200, seed=5)
    outs = [prep.prepare_table(t["sample"], t["transcript"], t["abundance"], t["significance"], t["do_check"], 250, threads=k)
            for k in (1, 2, 7)]
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and o[1] == outs[0][1] and o[2] == outs[0][2] and o[3] == outs[0][3]
        assert np.array_equal(o[4], outs[0][4])


def test_native_table_pass_rejects_bad_tables():
    t = _big_table(40, 6, 5, seed=2)
    args = lambda **kw: [kw.get(k, t[k]) for k in ("sample", "transcript", "abundance", "significance", "do_check")]
    r = int(np.flatnonzero(t["do_check"])[0])                          # a selected row, twice
    dup = {k: np.concatenate([v, v[r:r + 1]]) for k, v in t.items()}
    with pytest.raises(ValueError, match="duplicated"):
        prep.prepare_table(*[dup[k] for k in ("sample", "transcript", "abundance", "significance", "do_check")], 10)
    neg = t["abundance"].copy(); neg[3] = -1
    with pytest.raises(ValueError, match="non-negative"):
        prep.prepare_table(*args(abundance=neg), 10)
    big = t["abundance"].copy(); big[3] = 2**31
    with pytest.raises(ValueError, match="non-negative"):
        prep.prepare_table(*args(abundance=big), 10)
    sig = t["significance"].copy(); sig[0] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        prep.prepare_table(*args(significance=sig), 10)
    with pytest.raises(ValueError, match="do_check"):
        prep.prepare_table(*args(do_check=np.zeros(len(sig), bool)), 10)
    keep = np.ones(len(sig), bool); keep[r] = False
    with pytest.raises(ValueError, match="rectangular"):
        prep.prepare_table(*[t[k][keep] for k in ("sample", "transcript", "abundance", "significance", "do_check")], 10)


@pytest.mark.parametrize("kind", ["nb", "ties", "zeros", "tiny"])
def test_native_tmm_by_selection_equals_tmm_by_average_ranks(kind):
    """The native TMM finds the trimmed set by order-statistic selection + the tie groups' average ranks; the checker
    ranks explicitly (scipy rankdata).  Tie-heavy counts (values 0..6) put whole tie groups on the trim boundaries."""
    from oracle import prep_np
    rng = np.random.default_rng(11)
    if kind == "nb":
        c = rng.negative_binomial(3, 0.01, (5000, 24))
    elif kind == "ties":
        c = rng.integers(0, 7, (4000, 16))
    elif kind == "zeros":
        c = rng.negative_binomial(2, 0.05, (3000, 12))
        c[:, 5] = 0                                            # an empty library: factor 1 before the rescaling
        c[rng.random(3000) < 0.2] = 0                          # all-zero genes
    else:
        c = rng.integers(0, 50, (3, 5))
    c = c.astype(np.int32)
    order = rng.permutation(c.shape[1]).astype(np.int32)
    f, tot, ref = prep.tmm_factors(c, order)
    mat = c[:, order].astype(np.float64)
    ref_np = prep_np.reference_column(mat)
    assert ref == ref_np and np.array_equal(tot, mat.sum(axis=0))
    want = prep_np.tmm_norm_factors(mat, ref_np)
    assert np.allclose(f, want, rtol=1e-12, atol=0), np.abs(f / want - 1).max()
    f1 = prep.tmm_factors(c, order, ref_column=2, threads=1)[0]
    assert np.allclose(f1, prep_np.tmm_norm_factors(mat, 2), rtol=1e-12, atol=0)


def test_prepare_at_config_scale_timing():
    """3e6 rows (6,000 genes x 500 samples) through the whole prepare incl. TMM: seconds with per-row NumPy passes,
    well under one here."""
    G, S = 6000, 500
    rng = np.random.default_rng(1)
    ab = rng.negative_binomial(5, 0.02, G * S).astype(np.int64)
    sym, sam = np.repeat(np.arange(G, dtype=np.int64), S), np.tile(np.arange(S, dtype=np.int64), G)
    lab = np.tile(np.where(np.arange(S) % 2 > 0, "B", "A"), G)
    prep.prepare(sam[:S * 10], sym[:S * 10], ab[:S * 10], np.zeros(S * 10), np.ones(S * 10, bool), {"Label": lab[:S * 10]}, "~ Label", 0)
    t0 = time.perf_counter()
    p = prep.prepare(sam, sym, ab, np.repeat(np.linspace(0, 1, G), S), np.ones(G * S, bool), {"Label": lab}, "~ Label", 0)
    dt = time.perf_counter() - t0
    assert p.counts.shape == (G, S) and np.array_equal(p.counts.reshape(-1), ab) and abs(np.exp(np.mean(np.log(p.tmm))) - 1) < 1e-12
    assert dt < 5.0, dt


# ---- property test: random small tables, every selection / ordering corner ---------------------------------------------
def test_native_table_pass_random_tables_property():
    """Random rectangular tables (2-12 genes x 1-6 samples), rows in random order, random significance ties, do_check
    decided per ROW (so genes can be half checked), any number of negative controls, 1-5 forced row chunks (threads < 0):
    the native pass and the NumPy statement agree on everything."""
    from hypothesis import given, settings, strategies as st
    from oracle import prep_np

    @settings(max_examples=150, deadline=None)
    @given(st.integers(2, 12), st.integers(1, 6), st.integers(0, 15), st.integers(1, 5), st.integers(0, 2**31 - 1))
    def run(G, S, nctrl, threads, seed):
        rng = np.random.default_rng(seed)
        gid = rng.permutation(100)[:G].astype(np.int64)
        sid = rng.permutation(100)[:S].astype(np.int64)
        g, s = np.divmod(rng.permutation(G * S), S)
        sig = rng.integers(0, 4, G * S) / 4.0                       # per-row significance, many ties
        chk = rng.random(G * S) < rng.choice([0.1, 0.5, 1.0])
        ab = rng.integers(0, 1000, G * S).astype(np.int64)
        args = (sid[s], gid[g], ab, sig, chk, nctrl)
        if not chk.any():
            with pytest.raises(ValueError):
                prep.prepare_table(*args, threads=-threads)
            return
        try:
            ref = prep_np.prepare_table(*args)
        except ValueError as e:                                      # a half-selected gene leaves holes: both must refuse
            with pytest.raises(ValueError, match="rectangular|duplicated"):
                prep.prepare_table(*args, threads=-threads)
            assert "rectangular" in str(e) or "duplicated" in str(e)
            return
        counts, genes, samples, K, first_row = prep.prepare_table(*args, threads=-threads)
        assert K == ref["K"] and list(genes) == list(ref["genes"]) and list(samples) == list(ref["samples"])
        assert np.array_equal(counts, ref["counts"]) and np.array_equal(first_row, ref["first_row"])

    run()
