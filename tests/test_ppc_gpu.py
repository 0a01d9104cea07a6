"""GPU parity for the posterior-predictive summaries: quantiles/flags from an identical draws matrix must
be BIT-EXACT against the oracle (north_star); mean/sd follow the exact-integer definition."""
import numpy as np
import pytest

from oracle import quantile as Q
from tests.helpers import small_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,p", [(1000, 37, 0.05), (2100, 21, 4.761904761904762e-3), (10500, 9, 9.523809523809524e-4),
                                   (7, 5, 0.25), (1, 3, 0.1), (2, 4, 0.5), (1000, 8, 0.0), (333, 6, 1.0 / 3.0)])
def test_summarise_draws_bit_exact(n, m, p, built_lib):
    from ppcseq_b200 import ppc
    rng = np.random.default_rng(n + m)
    mu = np.exp(rng.uniform(0, 12, m))
    draws = rng.negative_binomial(2.0, 2.0 / (2.0 + mu), size=(n, m)).astype(np.float64)
    draws[:, 0] = 7.0                                   # all ties
    if m > 1:
        draws[:, 1] = rng.integers(0, 3, n)             # heavy ties
    lo, up, mean, sd = ppc.summarise_draws(draws, p)
    lo_r, up_r, mean_r, sd_r = Q.summarise_draws(draws, p)
    assert np.array_equal(lo, lo_r) and np.array_equal(up, up_r)
    assert np.array_equal(mean, mean_r)
    assert np.array_equal(sd, sd_r, equal_nan=True)


def test_summarise_rejects_non_integer(built_lib):
    from ppcseq_b200 import PpcseqError, ppc
    with pytest.raises(PpcseqError):
        ppc.summarise_draws(np.array([[0.5, 1.0], [2.0, 3.0]]), 0.1)


@pytest.mark.parametrize("C", [1, 2, 3])
def test_flags_bit_exact(C, built_lib):
    from ppcseq_b200 import NBModel, ppc
    G, S, K = 40, 21, 17
    d = small_problem(G, S, C, K, seed=4)
    rng = np.random.default_rng(8)
    c = d.counts[:K].astype(np.float64)
    lower = np.floor(c * rng.uniform(0.3, 1.2, (K, S)))
    upper = lower + np.floor(c * rng.uniform(0.0, 1.5, (K, S)))
    lower[0, :3] = c[0, :3]                              # boundary: count == lower  -> inside
    upper[1, :3] = c[1, :3]                              # boundary: count == upper  -> inside
    mean = (lower + upper) / 2 + rng.normal(0, 1, (K, S))
    mean[2, :4] = c[2, :4]                               # count == mean -> not "higher than mean"
    slope = rng.normal(0, 1, K)
    slope[3] = 0.0                                       # slope == 0: group_high is False for both groups
    m = NBModel(d.counts, d.X, d.exposure, K)
    out = ppc.flags(m, lower, upper, mean, slope)
    ref = Q.flags(d.counts[:K], lower, upper, mean, slope, d.X)
    assert np.array_equal(out["ppc"], ref["ppc"])
    assert np.array_equal(out["ppc_samples_failed"], ref["ppc_samples_failed"])
    if C > 1:
        assert np.array_equal(out["deleterious"], ref["deleterious"])
        assert np.array_equal(out["tot_deleterious_outliers"], ref["tot_deleterious_outliers"])
    else:
        assert out["deleterious"] is None and ref["deleterious"] is None


def _fit_problem(G=12, S=21, C=2, K=7, n_post=1000, seed=3, spread=0.05):
    """A small model + synthetic 'posterior' draws around a plausible point."""
    from ppcseq_b200 import Fit, NBModel
    from oracle import model_np
    d = small_problem(G, S, C, K, seed=seed)
    lay = model_np.Layout(G, K, C)
    rng = np.random.default_rng(seed)
    th0 = np.zeros(lay.D)
    th0[lay.o_intercept:lay.o_intercept + G] = rng.uniform(0.5, 9.0, G)
    th0[lay.o_alpha1:lay.o_alpha1 + K] = rng.normal(0, 0.5, K) if C >= 2 else 0
    th0[lay.o_sigma_raw:lay.o_sigma_raw + G] = rng.uniform(-2.5, 1.0, G)      # phi in [0.37, 12]
    draws = th0[None, :] + spread * rng.standard_normal((n_post, lay.D))
    m = NBModel(d.counts, d.X, d.exposure, K)
    return d, lay, th0, draws, m, Fit.from_draws(m, draws)


def test_fit_queries(built_lib):
    d, lay, th0, draws, m, fit = _fit_problem(n_post=257)
    assert fit.n_draws == 257
    got = fit.draws(lay.o_alpha1, m.K)
    assert np.array_equal(got, draws[:, lay.o_alpha1:lay.o_alpha1 + m.K])
    assert np.allclose(fit.slope(), draws[:, lay.o_alpha1:lay.o_alpha1 + m.K].mean(axis=0), rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("n_post,p", [(1000, 0.05), (300, 0.1), (2100, 4.761904761904762e-3)])
def test_streaming_summary_equals_summary_of_its_own_draws(n_post, p, built_lib):
    """The fused path never materialises draws; its tail selection must give exactly what the explicit
    type-7 summary of the same Philox stream gives (bit for bit)."""
    from ppcseq_b200 import ppc
    d, lay, th0, draws, m, fit = _fit_problem(n_post=n_post)
    for tc in (1.0, 0.7352941):
        raw = fit.ppc_draws(truncation_compensation=tc, seed=11)
        lo, up, mean, sd = fit.ppc_summary(p, exact=True, truncation_compensation=tc, seed=11)
        lo_r, up_r, mean_r, sd_r = Q.summarise_draws(raw.reshape(n_post, -1), p)
        assert np.array_equal(lo.ravel(), lo_r) and np.array_equal(up.ravel(), up_r)
        assert np.array_equal(mean.ravel(), mean_r) and np.array_equal(sd.ravel(), sd_r)
        lo_g, up_g, _, _ = ppc.summarise_draws(raw.reshape(n_post, -1), p)
        assert np.array_equal(lo.ravel(), lo_g) and np.array_equal(up.ravel(), up_g)


def test_nb_sampler_distribution(built_lib):
    """Philox gamma-Poisson draws against the NB2 law: moments (z-test) and a KS test per pair."""
    from scipy import stats
    n_post = 20000
    d, lay, th0, draws, m, fit = _fit_problem(G=6, S=5, C=2, K=6, n_post=n_post, spread=0.0)
    tc = 0.7352941
    raw = fit.ppc_draws(truncation_compensation=tc, seed=5)            # [n, K, S]
    alpha = np.zeros((2, 6)); alpha[0] = th0[lay.o_intercept:lay.o_intercept + 6]; alpha[1] = th0[lay.o_alpha1:lay.o_alpha1 + 6]
    eta = (d.X @ alpha).T + d.exposure[None, :]
    mu = np.exp(eta)
    phi = np.exp(-th0[lay.o_sigma_raw:lay.o_sigma_raw + 6])[:, None] * tc
    var = mu + mu * mu / phi
    z = (raw.mean(axis=0) - mu) / np.sqrt(var / n_post)
    assert np.abs(z).max() < 4.5, z
    # variance: compare on the log scale with a generous MC band (heavy-tailed fourth moment)
    assert np.all(np.abs(np.log(raw.var(axis=0) / var)) < 0.25)
    pmin = 1.0
    for g in range(6):
        for s in range(5):
            r = phi[g, 0]; pr = r / (r + mu[g, s])
            # randomised PIT makes the discrete KS test exact
            x = raw[:, g, s]
            u = np.random.default_rng(g * 5 + s).uniform(size=n_post)
            pit = stats.nbinom.cdf(x - 1, r, pr) + u * stats.nbinom.pmf(x, r, pr)
            pmin = min(pmin, stats.kstest(pit, "uniform").pvalue)
    assert pmin > 1e-4, pmin                                            # 30 tests: Bonferroni-safe


def test_supersampled_summary_close_to_exact(built_lib):
    """Approximate analysis (resample with replacement) agrees with the exact one within MC error."""
    d, lay, th0, draws, m, fit = _fit_problem(n_post=1000)
    lo_e, up_e, mean_e, _ = fit.ppc_summary(0.05, exact=True, seed=3)
    lo_a, up_a, mean_a, _ = fit.ppc_summary(0.05, exact=False, n_draws=2000, seed=4)     # p*n = 100 per tail
    assert np.all(np.abs(mean_a - mean_e) <= 0.2 * mean_e + 1.0)
    assert np.all(np.abs(up_a - up_e) <= 0.3 * up_e + 3.0)
    # reproducible for a fixed seed
    lo_a2, up_a2, mean_a2, _ = fit.ppc_summary(0.05, exact=False, n_draws=2000, seed=4)
    assert np.array_equal(up_a, up_a2) and np.array_equal(mean_a, mean_a2)


def test_wide_tail_falls_back_to_the_explicit_matrix(built_lib):
    """p * n beyond the 128 order statistics the streaming selection keeps: the draws are materialised and
    summarised explicitly -- same stream, so the result equals the oracle summary of ppc_draws bit for bit."""
    d, lay, th0, draws, m, fit = _fit_problem(n_post=1000)
    raw = fit.ppc_draws(seed=6)
    lo, up, mean, sd = fit.ppc_summary(0.4, exact=True, seed=6)
    lo_r, up_r, mean_r, sd_r = Q.summarise_draws(raw.reshape(1000, -1), 0.4)
    assert np.array_equal(lo.ravel(), lo_r) and np.array_equal(up.ravel(), up_r)
    assert np.array_equal(mean.ravel(), mean_r) and np.array_equal(sd.ravel(), sd_r)


def test_nb_sampler_tail_fidelity(built_lib):
    """Pass 2 at S = 500 .. 5,000 samples reads quantiles at p = 4e-5 .. 4e-6 from 2.5e5 .. 2.5e6 draws
    (R/methods.R:156-167), and the sampler's gamma / Poisson-rate arithmetic is fp32 (Stan's RNG is fp64): check the
    extreme order statistics of 2.5e6 draws per pair against EXACT negative-binomial tail masses (scipy.stats.nbinom).
    For the k-th smallest draw q of n: #{draws <= x} ~ Binomial(n, F(x)), so F(q - 1) n must not exceed k by more than
    its binomial noise and F(q) n must not fall short of it (and symmetrically in the upper tail); 4.5 sd, 40 tests.
    Cases cover every branch: shape < 1 (boost), rates below 10 (fp64 inversion), around 10 (both Poisson paths in
    one warp), PTRS, and rates near 1e6."""
    from scipy import stats
    from ppcseq_b200 import Fit, NBModel
    cases = [(3.0, 0.9), (8.0, 2.0), (12.0, 1.2), (50.0, 5.0), (5000.0, 20.0), (2.0e5, 0.6), (0.3, 0.4), (9.5, 30.0)]   # (mu, phi)
    G, S, tc = len(cases), 2, 0.7352941
    counts = np.ones((G, S), np.int32)
    m = NBModel(counts, np.ones((S, 1)), np.zeros(S), G)
    lay = m.layout
    th = np.zeros(lay.D)
    th[lay.o_intercept:lay.o_intercept + G] = np.log([c[0] for c in cases])
    th[lay.o_sigma_raw:lay.o_sigma_raw + G] = -np.log([c[1] for c in cases])
    fit = Fit.from_draws(m, np.tile(th, (64, 1)))
    n = 2_500_000
    for k in (11, 101):                                   # p ~ 4e-6 and 4e-5: 1 + (n - 1) p = k exactly => an order statistic
        p = (k - 1) / (n - 1)
        lo, up, mean, sd = fit.ppc_summary(p, exact=False, n_draws=n, truncation_compensation=tc, seed=17 + k)
        for g, (mu, phi) in enumerate(cases):
            r = phi * tc
            law = stats.nbinom(r, r / (r + mu))
            assert np.all(np.abs(mean[g] - mu) < 6 * np.sqrt(law.var() / n) + 1e-9), (mu, phi, mean[g])
            for s in range(S):
                q = lo[g, s]
                assert q == np.floor(q)
                a, b = law.cdf(q - 1) * n, law.cdf(q) * n                # expected #draws <= q - 1, <= q
                assert a - 4.5 * np.sqrt(a) < k <= b + 4.5 * np.sqrt(b) + 1, ("lower", mu, phi, k, q, a, b)
                q = up[g, s]
                a, b = law.sf(q) * n, law.sf(q - 1) * n                  # expected #draws > q, >= q
                assert a - 4.5 * np.sqrt(a) < k <= b + 4.5 * np.sqrt(b) + 1, ("upper", mu, phi, k, q, a, b)
