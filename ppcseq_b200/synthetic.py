"""Synthetic NB workloads of the shapes BASELINE.json names.  Seeded, no network."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .model import Layout

# name -> (G, S, C, K, pass2 mask, seed)    K = G: "all genes checked"
CONFIGS = {
    "cfg2_20kx21": dict(G=20_000, S=21, C=2, mask=False, seed=20241),
    "cfg3_60kx500": dict(G=60_000, S=500, C=3, mask=True, seed=20242),
    "cfg4_20kx2000": dict(G=20_000, S=2_000, C=2, mask=False, seed=20243),
    "cfg5_60kx5000": dict(G=60_000, S=5_000, C=2, mask=False, seed=20244),
}


@dataclass
class Workload:
    name: str
    counts: np.ndarray        # int32 [G,S]
    X: np.ndarray             # float64 [S,C]
    exposure: np.ndarray      # float64 [S]
    K: int
    exclude_pairs: np.ndarray  # int32 [n,2] (g,s); empty when the config has no pass-2 mask
    theta_true: np.ndarray    # unconstrained vector that generated the data
    lambda_mu_mu: float = 5.612671

    @property
    def G(self): return self.counts.shape[0]
    @property
    def S(self): return self.counts.shape[1]
    @property
    def C(self): return self.X.shape[1]
    @property
    def D(self): return Layout(self.G, self.K, self.C).D

    def algorithmic_bytes_per_eval(self) -> int:
        """B_eval of BASELINE.md section 3."""
        G, S, C, K = self.G, self.S, self.C, self.K
        b = 4 * G * S + 16 * (2 * G + K * (C - 1)) + 8 * S * (C + 1)
        if len(self.exclude_pairs):
            b += G * S // 8
        return b


def _skew_normal(rng, xi, omega, a, size):
    d = a / np.sqrt(1 + a * a)
    u0 = rng.standard_normal(size)
    v = rng.standard_normal(size)
    return xi + omega * (d * np.abs(u0) + np.sqrt(1 - d * d) * v)


def design(S: int, C: int) -> np.ndarray:
    """~Label (C=2) or ~Label+batch (C=3): balanced 2-level Label, 2-level batch crossed with Label."""
    X = np.ones((S, C))
    if C >= 2:
        X[:, 1] = (np.arange(S) % 2).astype(float)
    if C >= 3:
        X[:, 2] = ((np.arange(S) // 2) % 2).astype(float)
    for c in range(3, C):
        X[:, c] = ((np.arange(S) >> (c - 1)) % 2).astype(float)
    return X


def make(name: str = None, *, G=None, S=None, C=None, K=None, mask=False, seed=0, outlier_frac=1e-3) -> Workload:
    if name is not None:
        cfg = CONFIGS[name]
        G, S, C, mask, seed = cfg["G"], cfg["S"], cfg["C"], cfg["mask"], cfg["seed"]
    K = G if K is None else K
    rng = np.random.Generator(np.random.PCG64(seed))
    intercept = np.clip(_skew_normal(rng, 5.6, 2.0, -1.0, G), 0.0, 12.0)
    sigma_raw = rng.normal(-0.35 * intercept + 1.5, 0.5)
    phi = np.exp(-sigma_raw)
    X = design(S, C)
    alpha = np.zeros((C, G))
    alpha[0] = intercept
    if C >= 2:
        de = rng.random(G) < 0.05
        alpha[1, :K] = np.where(de[:K], rng.laplace(0.0, 1.0, K), 0.0)
    if C >= 3:
        alpha[2:, :K] = rng.normal(0.0, 0.3, (C - 2, K))
    exposure = rng.normal(0.0, 0.15, S)
    exposure -= exposure.mean()
    counts = np.empty((G, S), dtype=np.int32)
    step = max(1, (1 << 22) // S)
    for g0 in range(0, G, step):
        g1 = min(G, g0 + step)
        eta = (X @ alpha[:, g0:g1]).T + exposure[None, :]
        lam = rng.gamma(phi[g0:g1, None], np.exp(eta) / phi[g0:g1, None])
        counts[g0:g1] = np.minimum(rng.poisson(lam), 2**31 - 1).astype(np.int32)
    # gross outliers: 0.1 % of the elements multiplied by 10..100
    n_out = int(round(outlier_frac * G * S))
    flat = rng.choice(G * S, size=n_out, replace=False) if n_out else np.empty(0, dtype=np.int64)
    mult = rng.uniform(10.0, 100.0, n_out)
    cf = counts.reshape(-1)
    cf[flat] = np.minimum((cf[flat].astype(np.float64) + 1.0) * mult, 2**31 - 1).astype(np.int32)
    pairs = np.empty((0, 2), dtype=np.int32)
    if mask:
        pairs = np.stack([flat // S, flat % S], axis=1).astype(np.int32)
    lay = Layout(G, K, C)
    theta = np.zeros(lay.D)
    theta[0] = 0.0                      # lambda_mu - lambda_mu_mu
    theta[1] = np.log(2.0)
    theta[2] = -1.0
    theta[lay.o_intercept:lay.o_intercept + G] = intercept
    if C >= 2:
        theta[lay.o_alpha1:lay.o_alpha1 + K] = alpha[1, :K]
    if C >= 3:
        theta[lay.o_alpha2:lay.o_alpha2 + (C - 2) * K] = alpha[2:, :K].T.reshape(-1)
    theta[lay.o_sigma_raw:lay.o_sigma_raw + G] = sigma_raw
    theta[lay.o_tail] = np.log(0.35)
    theta[lay.o_tail + 1] = 1.5
    theta[lay.o_tail + 2] = np.log(0.5)
    return Workload(name or f"synthetic_{G}x{S}", counts, X, exposure, K, pairs, theta)


def random_thetas(w: Workload, n: int, seed: int = 1) -> np.ndarray:
    """n points ~ U(-2,2)^D (Stan's init range) -- the timing points of SURVEY.md 8d."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-2.0, 2.0, (n, w.D))
