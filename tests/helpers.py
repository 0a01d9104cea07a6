"""Shared builders for the parity tests (small seeded problems + tolerance rule)."""
import numpy as np

from oracle import model_np as M


def small_problem(G, S, C, K, seed, exclude_frac=0.0, big=False, continuous=False):
    rng = np.random.default_rng(seed)
    X = np.ones((S, C))
    for c in range(1, C):
        X[:, c] = rng.normal(size=S) if continuous else rng.integers(0, 2, S)
    mean = np.exp(rng.uniform(0.0, 9.0, G))[:, None]
    counts = rng.negative_binomial(3.0, 3.0 / (3.0 + mean), size=(G, S)).astype(np.int32)
    counts[0, 0] = 0
    if big:
        counts[1 % G, 1 % S] = 2580228          # largest count in the bundled dataset
        counts[2 % G, :] = 0                    # an all-zero gene
    ex = rng.normal(0, 0.15, S)
    excl = None
    if exclude_frac > 0:
        excl = rng.random((G, S)) < exclude_frac
    return M.ModelData(counts, X, ex, K, exclude=excl)


def grad_err(g, g_ref):
    """max_i |g_i - ref_i| / max(|ref_i|, 1e-3 * ||ref||_inf): the 1e-9 rule of SURVEY.md 7.2."""
    scale = np.maximum(np.abs(g_ref), 1e-3 * np.abs(g_ref).max())
    return float(np.max(np.abs(g - g_ref) / scale))


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
