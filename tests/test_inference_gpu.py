"""GPU: the device-resident samplers (NUTS, mean-field ADVI) and the two-pass identify_outliers() pipeline.

Parity levels (north_star): posterior means concordant within Monte Carlo error with the CPU oracle sampler
(tests/golden/nuts_golden.npz, made by oracle/nuts_np.py), and the outlier calls on the bundled dataset
identical to what the reference's own tests pin (tests/testthat/test-ppcSeq.R:26-30, :51-55; README.md:75-92).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tidy(npz):
    """Rebuild the tidy table the reference's API takes from a bundled fixture (gene-major, fixture order)."""
    import pandas as pd
    z = np.load(os.path.join(GOLD, npz))
    G, S = z["counts"].shape
    K = int(z["K"])
    label = np.where(z["X"][:, 1] > 0, "Neoadjuvant", "High")
    df = pd.DataFrame({
        "symbol": np.repeat(z["genes"], S), "sample": np.tile(z["samples"], G),
        "value": z["counts"].reshape(-1).astype(np.int32), "Label": np.tile(label, G),
        # any significance that keeps the controls in the tail works: checked genes small, controls large
        "PValue": np.repeat(np.where(np.arange(G) < K, 1e-6, 0.99), S),
        "is_significant": np.repeat(np.arange(G) < K, S)})
    return z, df


def test_nuts_concordant_with_oracle_sampler(built_lib):
    from ppcseq_b200 import NBModel, inference
    g = np.load(os.path.join(GOLD, "nuts_golden.npz"))
    m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
    fit = inference.sample_nuts(m, chains=4, iter=150 + 1000, warmup=150, seed=11)
    assert fit.n_draws == 4000
    info = fit.info(8)
    assert info[0] == 1 and info[1] > 4000 and 0.6 < info[5] <= 1.0 and info[6] > 0
    dr = fit.draws(0, m.D)
    mean, sd = dr.mean(axis=0), dr.std(axis=0, ddof=1)
    z = np.abs(mean - g["mean"]) / g["sd"]
    # MC error of either sampler is ~0.05-0.1 posterior sd per parameter (4000 autocorrelated draws each).  The last
    # parameter, log sigma_sigma, is the neck of the hierarchical funnel: chains linger there (divergent transitions,
    # the oracle's own four chains spread by 0.2 sd), so it gets the bound of a single poorly mixing chain.
    funnel = m.D - 1
    assert np.delete(z, funnel).max() < 0.3, (int(z.argmax()), float(z.max()))
    assert z[funnel] < 0.75 and np.percentile(z, 90) < 0.2
    sd_ok = np.delete((sd / g["sd"] > 0.75) & (sd / g["sd"] < 1.35), funnel)
    assert np.all(sd_ok)
    # chains are separate streams: their means must agree with each other too
    per_chain = dr.reshape(4, 1000, -1).mean(axis=1)
    assert np.delete(np.abs(per_chain - mean) / g["sd"], funnel, axis=1).max() < 0.4


def test_nuts_reproducible_and_options(built_lib):
    from ppcseq_b200 import NBModel, PpcseqError, inference
    g = np.load(os.path.join(GOLD, "nuts_golden.npz"))
    m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
    a = inference.sample_nuts(m, chains=2, iter=60, warmup=30, seed=3).draws(0, m.D)
    b = inference.sample_nuts(m, chains=2, iter=60, warmup=30, seed=3, threads=1).draws(0, m.D)
    assert np.array_equal(a, b)                                      # fixed seed => identical draws, any threading
    c = inference.sample_nuts(m, chains=2, iter=60, warmup=30, seed=4).draws(0, m.D)
    assert not np.array_equal(a, c)
    with pytest.raises(PpcseqError):
        inference.sample_nuts(m, chains=0, iter=10, warmup=5)
    with pytest.raises(PpcseqError):
        inference.sample_nuts(m, chains=1, iter=10, warmup=10)
    init = np.tile(g["mean"], (2, 1))
    d = inference.sample_nuts(m, chains=2, iter=40, warmup=20, seed=5, init=init).draws(0, m.D)
    assert np.isfinite(d).all()


def test_advi_concordant_with_oracle_sampler(built_lib):
    from ppcseq_b200 import NBModel, inference
    g = np.load(os.path.join(GOLD, "nuts_golden.npz"))
    m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
    fit = inference.advi(m, output_samples=2000, iter=50000, tol_rel_obj=0.005, seed=2)
    info = fit.info(8)
    assert info[0] == 2 and info[3] >= 100 and info[4] in (1, 2, 3) and np.isfinite(info[5])
    lay = m.layout
    dr = fit.draws(lay.o_intercept, m.G)
    ref_m, ref_s = g["mean"][lay.o_intercept:lay.o_intercept + m.G], g["sd"][lay.o_intercept:lay.o_intercept + m.G]
    # mean-field ADVI is an approximation: the gene intercepts (well identified) must land within ~1 posterior sd
    assert (np.abs(dr.mean(axis=0) - ref_m) / ref_s).max() < 1.0
    # reproducible for a fixed seed
    again = inference.advi(m, output_samples=2000, iter=50000, tol_rel_obj=0.005, seed=2).draws(lay.o_intercept, m.G)
    assert np.array_equal(dr, again)


@pytest.mark.parametrize("approx_analysis", [True, False])
def test_reference_testthat_outcome_vb(approx_analysis, built_lib):
    """tests/testthat/test-ppcSeq.R: VB, ~Label, 3 genes + 50 controls, pfp = 1 -> tot_deleterious_outliers == c(0,1,0)."""
    from ppcseq_b200.api import identify_outliers
    z, df = _tidy("bundled_test53.npz")
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=1,
                            tol_rel_obj=0.01, approximate_posterior_inference=True,
                            approximate_posterior_analysis=approx_analysis, how_many_negative_controls=50, cores=1, seed=7)
    assert list(res["symbol"]) == ["SLC16A12", "CYP1A1", "ART3"]
    assert list(res.iloc[:, 3].astype(int)) == [0, 1, 0]
    assert list(res.columns) == ["symbol", "sample_wise_data", "ppc_samples_failed", "tot_deleterious_outliers"]
    sw = res["sample_wise_data"].iloc[1]
    assert list(sw.columns) == ["S", "G", "value", "sample", "slope_before_outlier_filtering", "Label", "exposure_rate",
                                "multiplier", ".lower", ".upper", "slope_after_outlier_filtering",
                                "posterior_predictive_check_succeded", "deleterious_outliers"]
    # the flagged sample is the 5835-count one
    assert int(sw.loc[sw["deleterious_outliers"], "value"].iloc[0]) == 5835


def test_reference_outcome_nuts(built_lib):
    """The same configuration through the NUTS path (vignette path, R/utilities.R:1497-1512)."""
    from ppcseq_b200.api import identify_outliers
    z, df = _tidy("bundled_test53.npz")
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=1,
                            approximate_posterior_inference=False, approximate_posterior_analysis=True,
                            how_many_negative_controls=50, cores=4, seed=9)
    assert list(res["tot_deleterious_outliers"].astype(int)) == [0, 1, 0]
    assert res.attrs["fit 2 info"][0] == 1


def test_readme_table_vb(built_lib):
    """README.md:47-92: 15 genes with FDR < 0.01 + 500 controls, pfp = 5: CYP1A1 and LYZ fail one sample each."""
    from ppcseq_b200.api import identify_outliers
    z, df = _tidy("bundled_readme515.npz")
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=5, seed=22)
    assert list(res["symbol"]) == list(z["expected_genes"])
    failed, dele = res["ppc_samples_failed"].to_numpy(), res["tot_deleterious_outliers"].to_numpy()
    # The README run is an UNSEEDED VB run whose upper bound is a 99.5 % quantile of 2,100 draws (10 draws in the
    # tail): MMP8's count of 219 and CCNA1's largest count sit inside the Monte Carlo band of that bound (observed
    # 177-240 across seeds here), so those two calls flip between runs of the reference as well (measured over 40
    # seeds, scratch/dbg_readme2.py: MMP8 differs from the README in ~50 % of the runs, CCNA1 in ~40 %, SUSD4 in
    # 1 of 40, every other gene never -- the same frequencies before and after the sampler's random-word layout
    # changed).  Every other gene must match the README table exactly.
    robust = np.array([gname not in ("MMP8", "CCNA1") for gname in z["expected_genes"]])
    assert np.array_equal(failed[robust], z["expected"][robust, 0])
    assert np.array_equal(dele[robust], z["expected"][robust, 1])
    assert np.all(failed[~robust] <= 1) and np.all(dele[~robust] <= 1)


def test_intercept_only_formula_and_empty_check(built_lib):
    import pandas as pd
    from ppcseq_b200.api import identify_outliers
    z, df = _tidy("bundled_test53.npz")
    res = identify_outliers(df, "~ 1", sample="sample", transcript="symbol", abundance="value", significance="PValue",
                            do_check="is_significant", how_many_negative_controls=50, seed=3)
    assert list(res.columns) == ["symbol", "sample_wise_data", "ppc_samples_failed"]      # no deleterious column when C = 1
    assert res["ppc_samples_failed"].iloc[1] >= 1
    df2 = df.assign(is_significant=False)
    with pytest.warns(UserWarning):
        empty = identify_outliers(df2, "~ Label", sample="sample", transcript="symbol", abundance="value",
                                  significance="PValue", do_check="is_significant")
    assert isinstance(empty, pd.DataFrame) and len(empty) == 0
    assert list(empty.columns)[2:] == ["ppc samples failed", "tot deleterious_outliers"]   # names with spaces, R/methods.R:126


def test_result_formats_long_and_failing(built_lib):
    """f4 (R/utilities.R:539-608 at scale): the columnar formats carry exactly the nested frames' content."""
    from ppcseq_b200.api import identify_outliers
    z, df = _tidy("bundled_test53.npz")
    kw = dict(sample="sample", transcript="symbol", abundance="value", significance="PValue", do_check="is_significant",
              percent_false_positive_genes=1, how_many_negative_controls=50, cores=1, seed=7)
    tm = {}
    nested = identify_outliers(df, "~ Label", **kw, timings=tm)
    long = identify_outliers(df, "~ Label", **kw, return_format="long")
    failing = identify_outliers(df, "~ Label", **kw, return_format="failing")
    assert set(tm) >= {"prep_s", "upload_s", "pass1_s", "pass2_s", "result_s"} and all(tm[k] >= 0 for k in list(tm)[:5])
    import pandas as pd
    cat = pd.concat(list(nested["sample_wise_data"]), ignore_index=True)
    assert list(long.columns) == ["symbol"] + list(cat.columns)
    for c in cat.columns:
        assert np.array_equal(long[c].to_numpy(), cat[c].to_numpy()), c        # same seed => the same numbers
    assert list(long["symbol"].iloc[::21]) == list(nested["symbol"])
    tot = long.attrs["gene_totals"]
    assert list(tot["ppc_samples_failed"]) == list(nested["ppc_samples_failed"])
    assert list(tot["tot_deleterious_outliers"]) == list(nested["tot_deleterious_outliers"]) == [0, 1, 0]
    assert len(failing) == int(nested["ppc_samples_failed"].sum()) and not failing["posterior_predictive_check_succeded"].any()
    # the per-gene totals of the flags kernel agree with a recount from the rows
    recount = [int((~f["posterior_predictive_check_succeded"]).sum()) for f in nested["sample_wise_data"]]
    assert recount == list(nested["ppc_samples_failed"])


def test_batched_chain_driver_reproduces_the_threaded_driver(built_lib):
    """csrc/nuts_batched.cu (one host thread, every launch covers all chains: the driver of gene-sharded runs) against
    csrc/nuts.cu (a host thread and a stream per chain): same kernels' arithmetic, same Philox keys, same device-side
    tree decisions => the same draws, bit for bit, and the same diagnostics."""
    from ppcseq_b200 import NBModel, inference
    g = np.load(os.path.join(GOLD, "nuts_golden.npz"))
    m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
    a = inference.sample_nuts(m, chains=4, iter=150 + 120, warmup=150, seed=13)
    b = inference.sample_nuts(m, chains=4, iter=150 + 120, warmup=150, seed=13, threads=-1)
    da, db = a.draws(0, m.D), b.draws(0, m.D)
    assert np.isfinite(db).all() and np.array_equal(da, db)
    ia, ib = a.info(8), b.info(8)
    assert np.array_equal(ia[[3, 4, 5, 6, 7]], ib[[3, 4, 5, 6, 7]])       # divergences, tree hits, accept, step size, leapfrogs
