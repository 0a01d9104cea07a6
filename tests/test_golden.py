"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py in the build container):
CPU -- the fp64 oracles reproduce the 40-digit mpmath golden vectors and the quantile/flag vectors;
       the input preparation (gene selection, indexing, TMM exposure) reproduces the bundled fixtures
       when the reference dataset is present (build container only).
GPU -- the CUDA path, through the C ABI, against the same golden vectors.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle, model_np
from oracle import quantile as Q
from tests.helpers import grad_err, rel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def _data53(exclude=None):
    z = load("bundled_test53.npz")
    return model_np.ModelData(z["counts"], z["X"], z["exposure_rate"], int(z["K"]), exclude=exclude)


def test_fixture_shapes_and_reference_facts():
    z = load("bundled_test53.npz")
    assert z["counts"].shape == (53, 21) and int(z["K"]) == 3
    assert list(z["genes"][:3]) == ["SLC16A12", "CYP1A1", "ART3"]          # tests/testthat/test-ppcSeq.R:11
    assert list(z["expected"]) == [0, 1, 0]                                  # :26-30
    # CYP1A1 carries 820 and 5835 against a typical 0-50
    assert sorted(z["counts"][1])[-2:] == [820, 5835]
    r = load("bundled_readme515.npz")
    assert r["counts"].shape == (515, 21) and int(r["K"]) == 15
    assert r["counts"].max() == 262664
    assert list(r["genes"][:15]) == list(r["expected_genes"])
    assert np.array_equal(z["X"][:, 1], r["X"][:, 1]) and z["X"][:, 1].sum() == 11   # Neoadjuvant x 11


@pytest.mark.parametrize("which", ["np", "c"])
def test_oracles_reproduce_mpmath_golden(which):
    g = load("lp_grad_golden.npz")
    for e, ex in enumerate([None, g["exclude"]]):
        d = _data53(ex)
        for mi, (pr, ja) in enumerate(g["modes"]):
            for ti, th in enumerate(g["thetas"]):
                if which == "np":
                    lp, gr = model_np.log_prob_grad(d, th, bool(pr), bool(ja))
                else:
                    lp, gr = c_oracle.log_prob_grad(d, th, bool(pr), bool(ja), n_shards=2)
                assert rel(lp, g["lp"][e, mi, ti]) < 1e-13
                assert grad_err(gr, g["grad"][e, mi, ti]) < 1e-11      # fp64 cancellation near the mode


def test_quantile_oracle_reproduces_golden():
    g = load("quantile_golden.npz")
    lo, up, mean, sd = Q.summarise_draws(g["draws"].astype(np.float64), float(g["p"]))
    assert np.array_equal(lo, g["lower"]) and np.array_equal(up, g["upper"])
    assert np.array_equal(mean, g["mean"]) and np.array_equal(sd, g["sd"])
    z = load("bundled_test53.npz")
    fl = Q.flags(z["counts"][:3], lo.reshape(3, 21), up.reshape(3, 21), mean.reshape(3, 21), g["slope"], z["X"])
    assert np.array_equal(fl["ppc"], g["ppc"]) and np.array_equal(fl["deleterious"], g["deleterious"])
    assert np.array_equal(fl["tot_deleterious_outliers"], g["tot_deleterious_outliers"])


@pytest.mark.skipif(not os.path.exists("/root/reference/data/counts.rda"), reason="reference dataset not on this box")
def test_prep_reproduces_bundled_fixture():
    from oracle import rda
    from ppcseq_b200 import prep
    raw = rda.load_rda("/root/reference/data/counts.rda")["counts"]
    chk = np.array([s in ("SLC16A12", "CYP1A1", "ART3") for s in raw["symbol"]])
    p = prep.prepare(raw["sample"], raw["symbol"], raw["value"], raw["PValue"], chk, {"Label": raw["Label"]},
                     "~ Label", how_many_negative_controls=50)
    z = load("bundled_test53.npz")
    assert np.array_equal(p.counts, z["counts"]) and np.array_equal(p.X, z["X"])
    # the fixture's exposure was written by the NumPy statement of TMM (explicit ranks); the native one (selection) agrees to rounding
    assert np.allclose(p.exposure_rate, z["exposure_rate"], rtol=0, atol=1e-14) and p.genes == list(z["genes"])


def test_tmm_properties():
    """TMM factors have geometric mean 1; identical libraries give 1; a scaled copy gives 1 (scale-free)."""
    from ppcseq_b200 import prep
    rng = np.random.default_rng(0)
    base = rng.negative_binomial(5, 0.01, size=(400, 1)).astype(float)
    mat = np.hstack([base, base * 3.0, rng.negative_binomial(5, 0.01, size=(400, 2))])
    f = prep.tmm_norm_factors(mat, 0)
    assert abs(np.exp(np.mean(np.log(f))) - 1.0) < 1e-12
    f2 = prep.tmm_norm_factors(np.hstack([base, base * 3.0]), 0)
    assert np.allclose(f2, 1.0, atol=1e-12)


def test_model_matrix_treatment_contrasts():
    from ppcseq_b200 import prep
    X, names = prep.model_matrix("~ Label + batch + age", {"Label": ["b", "a", "b", "c"], "batch": ["x", "y", "x", "y"],
                                                            "age": np.array([1.0, 2.0, 3.0, 4.0])}, 4)
    assert names == ["(Intercept)", "Labelb", "Labelc", "batchy", "age"]
    assert np.array_equal(X[:, 1], [1, 0, 1, 0]) and np.array_equal(X[:, 2], [0, 0, 0, 1])
    assert np.array_equal(X[:, 3], [0, 1, 0, 1]) and np.array_equal(X[:, 4], [1, 2, 3, 4])
    X1, n1 = prep.model_matrix("~ 1", {}, 3)
    assert X1.shape == (3, 1) and n1 == ["(Intercept)"]


# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_reproduces_mpmath_golden(built_lib):
    from ppcseq_b200 import NBModel
    g = load("lp_grad_golden.npz")
    z = load("bundled_test53.npz")
    m = NBModel(z["counts"], z["X"], z["exposure_rate"], int(z["K"]))
    for e, ex in enumerate([None, g["exclude"]]):
        m.set_exclusion(np.argwhere(ex) if ex is not None else np.empty((0, 2), np.int32))
        for path in (1, 2, 3):
            m.set_design_path(path)
            for mi, (pr, ja) in enumerate(g["modes"]):
                lps, grs = m.log_prob_grad(g["thetas"], bool(pr), bool(ja))
                for ti in range(len(lps)):
                    assert rel(lps[ti], g["lp"][e, mi, ti]) < 1e-10, (e, path, mi, ti)
                    assert grad_err(grs[ti], g["grad"][e, mi, ti]) < 1e-10, (e, path, mi, ti)


@pytest.mark.gpu
def test_gpu_reproduces_quantile_golden(built_lib):
    from ppcseq_b200 import NBModel, ppc
    g = load("quantile_golden.npz")
    z = load("bundled_test53.npz")
    lo, up, mean, sd = ppc.summarise_draws(g["draws"].astype(np.float64), float(g["p"]))
    assert np.array_equal(lo, g["lower"]) and np.array_equal(up, g["upper"])
    assert np.array_equal(mean, g["mean"]) and np.array_equal(sd, g["sd"])
    m = NBModel(z["counts"], z["X"], z["exposure_rate"], int(z["K"]))
    fl = ppc.flags(m, lo.reshape(3, 21), up.reshape(3, 21), mean.reshape(3, 21), g["slope"])
    assert np.array_equal(fl["ppc"], g["ppc"]) and np.array_equal(fl["deleterious"], g["deleterious"])
    assert np.array_equal(fl["ppc_samples_failed"], g["ppc_samples_failed"])
    assert np.array_equal(fl["tot_deleterious_outliers"], g["tot_deleterious_outliers"])
