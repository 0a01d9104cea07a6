"""CPU: the fp64 oracles (NumPy, C) against the 40-digit mpmath truth, and the mp partials against
numeric differentiation of the direct transcription.  No GPU, no product code."""
import numpy as np
import pytest

from oracle import c_oracle, model_mp, model_np
from tests.helpers import grad_err, rel, small_problem

CASES = [  # G, S, C, K, exclude_frac, continuous
    (5, 6, 3, 3, 0.1, False),
    (7, 4, 1, 2, 0.0, False),
    (9, 5, 2, 9, 0.05, True),
    (6, 8, 4, 4, 0.1, True),
]


@pytest.mark.parametrize("G,S,C,K,ef,cont", CASES)
@pytest.mark.parametrize("propto,jac", [(True, True), (False, True), (True, False)])
def test_np_and_c_match_mpmath(G, S, C, K, ef, cont, propto, jac):
    d = small_problem(G, S, C, K, seed=G * 100 + S, exclude_frac=ef, big=True, continuous=cont)
    rng = np.random.default_rng(7)
    th = rng.uniform(-2, 2, model_np.dim(G, K, C))
    lp_mp, g_mp = model_mp.to_float(*model_mp.log_prob_grad(d, th, propto, jac))
    lp_np, g_np = model_np.log_prob_grad(d, th, propto, jac)
    lp_c, g_c = c_oracle.log_prob_grad(d, th, propto, jac, n_shards=3)
    assert rel(lp_np, lp_mp) < 1e-13 and grad_err(g_np, g_mp) < 1e-12
    assert rel(lp_c, lp_mp) < 1e-13 and grad_err(g_c, g_mp) < 1e-12


def test_mp_partials_match_numeric_differentiation():
    d = small_problem(4, 5, 3, 2, seed=3, exclude_frac=0.1, big=True, continuous=True)
    th = np.random.default_rng(5).uniform(-2, 2, model_np.dim(4, 2, 3))
    _, g = model_mp.log_prob_grad(d, th)
    ng = model_mp.numeric_grad(d, th, range(len(th)))
    assert max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(ng, g)) < 1e-25


def test_adversarial_values():
    """n = 0, n = 2.58e6, phi in {1e-3, 1e5}, eta in {-20, +20}."""
    G, S, C, K = 6, 4, 2, 6
    d = small_problem(G, S, C, K, seed=11, big=True)
    lay = model_np.Layout(G, K, C)
    th = np.zeros(lay.D)
    th[lay.o_intercept:lay.o_intercept + G] = [-20, 20, 0, 5, -20, 20]
    th[lay.o_sigma_raw:lay.o_sigma_raw + G] = [np.log(1e3), np.log(1e3), -np.log(1e5), -np.log(1e5), 0, 0]
    lp_mp, g_mp = model_mp.to_float(*model_mp.log_prob_grad(d, th))
    lp_np, g_np = model_np.log_prob_grad(d, th)
    lp_c, g_c = c_oracle.log_prob_grad(d, th)
    assert rel(lp_np, lp_mp) < 1e-12 and grad_err(g_np, g_mp) < 1e-10
    assert rel(lp_c, lp_mp) < 1e-12 and grad_err(g_c, g_mp) < 1e-10


def test_c_oracle_thread_count_invariance():
    d = small_problem(40, 9, 3, 20, seed=2, exclude_frac=0.05)
    th = np.random.default_rng(1).uniform(-2, 2, model_np.dim(40, 20, 3))
    lp1, g1 = c_oracle.log_prob_grad(d, th, n_shards=1)
    lp4, g4 = c_oracle.log_prob_grad(d, th, n_shards=4)
    assert rel(lp4, lp1) < 1e-14 and np.allclose(g1, g4, rtol=1e-13, atol=0)
