"""CPU: pin the oracles (mpmath golden, NumPy, C) with something they did not write -- oracle/stan_literal.py, a
literal transcription of negBinomial_MPI.stan INCLUDING the reference's map_rect packing, whose gradient comes from
torch.autograd (the mechanism Stan uses) and whose densities are cross-checked against scipy.stats.

Tolerances: the literal transcription evaluates lgamma(n + phi) - lgamma(n + 1) in plain fp64, which loses digits at
large counts, so counts here are moderate and the bars are 1e-9 on lp (relative) and 1e-7 on the
gradient (the 1e-9-rule scale of tests/helpers.grad_err); the 1e-10 claims of the other tests rest on mpmath."""
import os

import numpy as np
import pytest

from oracle import c_oracle, model_np, stan_literal
from tests.helpers import grad_err, rel

HERE = os.path.dirname(os.path.abspath(__file__))


def _problem(G, S, C, K, seed, exclude_frac, continuous=False):
    rng = np.random.default_rng(seed)
    X = np.ones((S, C))
    for c in range(1, C):
        X[:, c] = rng.normal(size=S) if continuous else rng.integers(0, 2, S)
    mean = np.exp(rng.uniform(0.0, 7.0, G))[:, None]
    counts = rng.negative_binomial(3.0, 3.0 / (3.0 + mean), size=(G, S)).astype(np.int32)
    counts[0, 0] = 0
    ex = rng.normal(0, 0.3, S)
    excl = (rng.random((G, S)) < exclude_frac) if exclude_frac > 0 else None
    return model_np.ModelData(counts, X, ex, K, exclude=excl)


CASES = [  # G, S, C, K, exclude_frac, continuous, shards
    (7, 5, 1, 3, 0.0, False, 1),
    (11, 6, 2, 11, 0.1, False, 3),
    (10, 7, 3, 4, 0.15, False, 4),       # K < G, shards that do not divide G, R = 1
    (9, 5, 4, 9, 0.1, True, 2),          # continuous covariates, R = 2
    (5, 4, 2, 5, 0.3, False, 8),         # more shards than genes: empty shards vanish (n_shards = min(...))
]


@pytest.mark.parametrize("G,S,C,K,ef,cont,shards", CASES)
@pytest.mark.parametrize("jac", [True, False])
def test_oracles_match_the_literal_transcription(G, S, C, K, ef, cont, shards, jac):
    d = _problem(G, S, C, K, seed=31 * G + S, exclude_frac=ef, continuous=cont)
    for th_seed in (1, 2):
        th = np.random.default_rng(th_seed).uniform(-2, 2, model_np.dim(G, K, C))
        lp_t, g_t = stan_literal.log_prob_grad(d.counts, d.X, d.exposure, K, th, d.lambda_mu_mu, d.exclude, jac, shards)
        lp_s = stan_literal.log_prob_scipy(d.counts, d.X, d.exposure, K, th, d.lambda_mu_mu, d.exclude, jac)
        lp_n, g_n = model_np.log_prob_grad(d, th, False, jac)
        lp_c, g_c = c_oracle.log_prob_grad(d, th, False, jac, n_shards=2)
        assert rel(lp_t, lp_s) < 1e-9                       # autograd graph vs scipy densities
        assert rel(lp_n, lp_t) < 1e-9 and rel(lp_c, lp_t) < 1e-9 and rel(lp_n, lp_s) < 1e-9
        assert grad_err(g_n, g_t) < 1e-7 and grad_err(g_c, g_t) < 1e-7
        # propto = true drops data-only constants: the gradient must be the same vector
        _, g_p = c_oracle.log_prob_grad(d, th, True, jac, n_shards=1)
        assert grad_err(g_p, g_t) < 1e-7


def test_packing_is_shard_count_invariant():
    d = _problem(13, 6, 3, 7, seed=5, exclude_frac=0.2)
    th = np.random.default_rng(3).uniform(-2, 2, model_np.dim(13, 7, 3))
    ref = stan_literal.log_prob_grad(d.counts, d.X, d.exposure, 7, th, d.lambda_mu_mu, d.exclude, True, 1)
    for shards in (2, 5, 13):
        lp, g = stan_literal.log_prob_grad(d.counts, d.X, d.exposure, 7, th, d.lambda_mu_mu, d.exclude, True, shards)
        assert rel(lp, ref[0]) < 1e-13 and np.allclose(g, ref[1], rtol=1e-11, atol=1e-11)


def test_mpmath_golden_matches_the_literal_transcription():
    """The committed mpmath golden values (bundled 53-gene problem, tests/golden/lp_grad_golden.npz) against the
    literal transcription: the bundled counts reach 24,912, where plain fp64 lgamma differences keep ~1e-9."""
    g = np.load(os.path.join(HERE, "golden", "lp_grad_golden.npz"))
    b = np.load(os.path.join(HERE, "golden", "bundled_test53.npz"), allow_pickle=True)
    counts, X, ex, K = b["counts"], b["X"], b["exposure_rate"], int(b["K"])
    modes = [tuple(int(x) for x in m) for m in g["modes"]]
    for e, excl in enumerate((None, g["exclude"].astype(bool))):
        for mi, (propto, jac) in enumerate(modes):
            if propto:
                continue                                    # the transcription keeps every constant
            for ti, th in enumerate(g["thetas"]):
                lp, gr = stan_literal.log_prob_grad(counts, X, ex, K, th, 5.612671, excl, bool(jac), shards=4)
                assert rel(lp, g["lp"][e, mi, ti]) < 1e-9
                assert grad_err(gr, g["grad"][e, mi, ti]) < 1e-7
    # gradients of the propto = true modes equal those of propto = false
    assert any(not m[0] for m in modes)


# ---- the summary statistics of the PPC (R quantile type 7, mean, sd) against third-party implementations ---------------
@pytest.mark.parametrize("n,p", [(1000, 0.05), (1000, 0.025), (8000, 0.00125), (20, 0.05), (2, 0.3), (1, 0.5), (999, 0.5)])
def test_quantile_oracle_against_numpy_scipy_pandas_type7(n, p):
    """oracle/quantile.py restates R's quantile.default(type = 7) by hand; NumPy's method="linear", SciPy's
    mquantiles(alphap = 1, betap = 1) and pandas' Series.quantile are independent implementations of the same
    Hyndman-Fan definition 7.  Integer-valued draws with many ties, as the PPC produces."""
    import pandas as pd
    from scipy.stats.mstats import mquantiles
    from oracle import quantile as Q
    rng = np.random.default_rng(n)
    draws = rng.negative_binomial(3, 0.02, (n, 40)).astype(np.float64)
    draws[:, 0] = 7.0                                   # a constant pair
    lo, up, mean, sd = Q.summarise_draws(draws, p)
    for probs, got in ((p, lo), (1.0 - p, up)):
        want_np = np.quantile(draws, probs, axis=0, method="linear")
        want_sp = np.asarray(mquantiles(draws, prob=[probs], alphap=1, betap=1, axis=0))[0]
        want_pd = pd.DataFrame(draws).quantile(probs).to_numpy()
        for want in (want_np, want_sp, want_pd):
            assert np.allclose(got, want, rtol=1e-13, atol=0)
    assert np.allclose(mean, draws.mean(axis=0), rtol=1e-14, atol=0)
    if n > 1:
        assert np.allclose(sd, draws.std(axis=0, ddof=1), rtol=1e-11, atol=1e-12)
    else:
        assert np.isnan(sd).all()
