"""NumPy fp64 restatement of the ppcseq Stan model's log-density and gradient.  ORACLE ONLY.

Follows /root/reference/inst/stan/negBinomial_MPI.stan:
  * parameters block  :180-199  -> `unpack` / `dim`
  * transformed params :200-206 -> phi = exp(-sigma_raw), alpha (merge_coefficients :122-139),
                                   eta = X * alpha  (+ exposure, added in lp_reduce :100-101)
  * priors            :210-223  -> `_priors`
  * likelihood        :58-120, :226-240 -> `_likelihood` (exclusion term :105-115 as a 0/1 weight)
Gradients are hand-derived (Stan obtains them by reverse-mode AD); both are validated against
the 40-digit mpmath evaluation in oracle/model_mp.py by tests/test_oracle.py.

PARITY STATUS: parity unpinned (see oracle/__init__.py) -- no reference golden vectors exist.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy import special as sp

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
SQRT_2_OVER_PI = math.sqrt(2.0 / math.pi)


@dataclass
class ModelData:
    """Dense restatement of the Stan data block (negBinomial_MPI.stan:142-173).

    counts is gene-major [G, S]; the map_rect shard packing (counts_package, symbol_end, ...)
    is an artefact of the CPU sharding and is not represented.
    """
    counts: np.ndarray            # int32 [G, S]
    X: np.ndarray                 # float64 [S, C]
    exposure: np.ndarray          # float64 [S]      (exposure_rate, :168)
    K: int                        # how_many_to_check (:165)
    lambda_mu_mu: float = 5.612671  # R/methods.R:218
    exclude: np.ndarray | None = None  # bool [G, S], True = dropped from the likelihood (:105-115)
    truncation_compensation: float = 1.0
    G: int = field(init=False)
    S: int = field(init=False)
    C: int = field(init=False)

    def __post_init__(self):
        self.counts = np.ascontiguousarray(self.counts, dtype=np.int32)
        self.X = np.ascontiguousarray(self.X, dtype=np.float64)
        self.exposure = np.ascontiguousarray(self.exposure, dtype=np.float64)
        self.G, self.S = self.counts.shape
        self.C = self.X.shape[1]
        assert self.X.shape[0] == self.S and self.exposure.shape == (self.S,)
        assert 0 <= self.K <= self.G
        if self.exclude is not None:
            self.exclude = np.ascontiguousarray(self.exclude, dtype=bool)
            assert self.exclude.shape == self.counts.shape


def dim(G: int, K: int, C: int) -> int:
    """Length of the unconstrained vector (declaration order, :180-199)."""
    return 6 + 2 * G + K + max(0, C - 2) * K


@dataclass
class Layout:
    G: int
    K: int
    C: int

    @property
    def o_intercept(self): return 3
    @property
    def o_alpha1(self): return 3 + self.G
    @property
    def o_alpha2(self): return 3 + self.G + self.K
    @property
    def o_sigma_raw(self): return 3 + self.G + self.K + max(0, self.C - 2) * self.K
    @property
    def o_tail(self): return self.o_sigma_raw + self.G
    @property
    def D(self): return self.o_tail + 3


def unpack(theta: np.ndarray, G: int, K: int, C: int) -> dict:
    L = Layout(G, K, C)
    assert theta.shape == (L.D,), (theta.shape, L.D)
    R = max(0, C - 2)
    return dict(
        u_lm=theta[0], u_ls=theta[1], lambda_skew=theta[2],
        intercept=theta[L.o_intercept:L.o_intercept + G],
        alpha_sub_1=theta[L.o_alpha1:L.o_alpha1 + K],
        # Stan stores matrix[R, K] column-major: element (r, k) at k*R + r
        alpha_2=theta[L.o_alpha2:L.o_alpha2 + R * K].reshape(K, R).T,
        sigma_raw=theta[L.o_sigma_raw:L.o_sigma_raw + G],
        u_ss=theta[L.o_tail], sigma_intercept=theta[L.o_tail + 1], u_sg=theta[L.o_tail + 2],
    )


def alpha_matrix(p: dict, G: int, K: int, C: int) -> np.ndarray:
    """merge_coefficients (:122-139): C x G, rows >= 2 are zero for genes > K."""
    a = np.zeros((C, G))
    a[0] = p["intercept"]
    if C >= 2:
        a[1, :K] = p["alpha_sub_1"]
    if C >= 3:
        a[2:, :K] = p["alpha_2"]
    return a


def _stirling_tail(x):
    """lgamma(x) - [(x-1/2)log x - x + 1/2 log 2pi] for x >= 16."""
    w = 1.0 / (x * x)
    return (1.0 / x) * (1.0 / 12 + w * (-1.0 / 360 + w * (1.0 / 1260 + w * (-1.0 / 1680 + w * (1.0 / 1188 + w * (-691.0 / 360360))))))


def lgamma_ratio(n, phi):
    """lgamma(n+phi) - lgamma(n+1), free of the catastrophic cancellation at large n.

    For n+1 >= 32 it is written as (y-1/2) log1p(d/y) + d log(x) - d + tail(x) - tail(y) with
    y = n+1, d = phi-1, x = n+phi.
    """
    n = np.asarray(n, dtype=np.float64)
    phi = np.broadcast_to(np.asarray(phi, dtype=np.float64), n.shape)
    out = np.empty_like(n)
    small = (n < 32) | (n + phi < 32)
    out[small] = sp.gammaln(n[small] + phi[small]) - sp.gammaln(n[small] + 1.0)
    b = ~small
    y = n[b] + 1.0
    d = phi[b] - 1.0
    x = n[b] + phi[b]
    out[b] = (y - 0.5) * np.log1p(d / y) + d * np.log(x) - d + _stirling_tail(x) - _stirling_tail(y)
    return out


def _likelihood(d: ModelData, eta, phi):
    """Sum of neg_binomial_2_log_lpmf over the non-excluded elements and its partials.

    eta [G,S] includes exposure; phi [G].  Returns (ll, dll/deta [G,S], dll/dphi [G]).
    Closed forms, mu = exp(eta), a = mu + phi:
      ell      = lgamma(n+phi) - lgamma(n+1) - lgamma(phi) + n eta + phi log phi - (n+phi) log a
      d/d eta  = n - (n+phi) mu / a
      d/d phi  = (mu - n)/a + log phi - log a - psi(phi) + psi(n+phi)
    """
    n = d.counts.astype(np.float64)
    ph = phi[:, None]
    mu = np.exp(eta)
    a = mu + ph
    log_a = np.log(a)
    ell = lgamma_ratio(n, ph) - sp.gammaln(ph) + n * eta + ph * np.log(ph) - (n + ph) * log_a
    d_eta = n - (n + ph) * (mu / a)
    d_phi = (mu - n) / a + np.log(ph) - log_a - sp.digamma(ph) + sp.digamma(n + ph)
    if d.exclude is not None:
        w = ~d.exclude
        ell = ell * w
        d_eta = d_eta * w
        d_phi = d_phi * w
    return math.fsum(ell.ravel()), d_eta, d_phi.sum(axis=1)


def log_prob_grad(d: ModelData, theta: np.ndarray, propto: bool = True, jacobian: bool = True):
    """log_prob<propto, jacobian>(theta) and its gradient w.r.t. the unconstrained vector."""
    theta = np.asarray(theta, dtype=np.float64)
    G, S, C, K = d.G, d.S, d.C, d.K
    Lo = Layout(G, K, C)
    p = unpack(theta, G, K, C)
    L = d.lambda_mu_mu
    lambda_mu = p["u_lm"] + L                      # :183 offset
    lambda_sigma = math.exp(p["u_ls"])              # :184 lower=0
    lambda_skew = p["lambda_skew"]
    sigma_slope = -math.exp(p["u_ss"])              # :195 upper=0
    sigma_intercept = p["sigma_intercept"]
    sigma_sigma = math.exp(p["u_sg"])               # :197 lower=0
    intercept, sigma_raw = p["intercept"], p["sigma_raw"]

    phi = np.exp(-sigma_raw)                        # :203
    alpha = alpha_matrix(p, G, K, C)                # :204
    eta = (d.X @ alpha).T + d.exposure[None, :]     # :205 and :100-101  -> [G,S]

    ll, d_eta, d_phi = _likelihood(d, eta, phi)
    g = np.zeros(Lo.D)
    d_alpha = d_eta @ d.X                           # [G,C] adjoint of X*alpha
    g_intercept = d_alpha[:, 0].copy()
    g[Lo.o_alpha1:Lo.o_alpha1 + K] = d_alpha[:K, 1] if C >= 2 else 0.0
    if C >= 3:
        g[Lo.o_alpha2:Lo.o_alpha2 + (C - 2) * K] = d_alpha[:K, 2:].reshape(-1)   # (k, r) -> k*R + r
    g_sigma_raw = -phi * d_phi

    lp = ll
    # ---- priors (:210-223) -------------------------------------------------------------
    g_lambda_mu = g_lambda_sigma = g_lambda_skew = 0.0
    g_sigma_slope = g_sigma_intercept = g_sigma_sigma = 0.0
    lp += -(lambda_mu - L) ** 2 / 8.0;  g_lambda_mu += -(lambda_mu - L) / 4.0
    lp += -lambda_sigma ** 2 / 8.0;     g_lambda_sigma += -lambda_sigma / 4.0
    lp += -lambda_skew ** 2 / 2.0;      g_lambda_skew += -lambda_skew
    lp += -sigma_intercept ** 2 / 8.0;  g_sigma_intercept += -sigma_intercept / 4.0
    lp += -sigma_slope ** 2 / 8.0;      g_sigma_slope += -sigma_slope / 4.0
    lp += -sigma_sigma ** 2 / 8.0;      g_sigma_sigma += -sigma_sigma / 4.0
    if not propto:
        lp += 5 * (-HALF_LOG_2PI - math.log(2.0)) + (-HALF_LOG_2PI)

    # intercept ~ skew_normal(lambda_mu + lambda_mu_mu, lambda_sigma, lambda_skew)  (:219, L twice)
    xi = lambda_mu + L
    om = lambda_sigma
    a = lambda_skew
    z = (intercept - xi) / om
    t = -a * z / math.sqrt(2.0)
    log_erfc = np.log(sp.erfcx(t)) - t * t          # log erfc(t), stable for large positive t
    neg = t < 0
    log_erfc[neg] = np.log(sp.erfc(t[neg]))
    lp += -G * math.log(om) + math.fsum(-0.5 * z * z + log_erfc)
    if not propto:
        lp += -G * HALF_LOG_2PI
    # r = a*sqrt(2/pi)*exp(-a^2 z^2/2)/erfc(-a z/sqrt2) = a*sqrt(2/pi)/erfcx(t)  (exp(-t^2)/erfc(t) = 1/erfcx(t))
    r = a * SQRT_2_OVER_PI / sp.erfcx(t)
    dz = -z + r                                     # d/dz of the per-gene term
    g_intercept += dz / om
    g_lambda_mu += -dz.sum() / om
    g_lambda_sigma += (-G + (-dz * z).sum()) / om
    g_lambda_skew += (SQRT_2_OVER_PI / sp.erfcx(t) * z).sum()

    if C >= 2:                                      # :220
        a1 = p["alpha_sub_1"]
        lp += -np.abs(a1).sum()
        g[Lo.o_alpha1:Lo.o_alpha1 + K] += -np.sign(a1)
        if not propto:
            lp += -K * math.log(2.0)
    if C >= 3:                                      # :221
        a2 = p["alpha_2"]
        lp += -(a2 * a2).sum() / 12.5
        g[Lo.o_alpha2:Lo.o_alpha2 + (C - 2) * K] += (-a2 / 6.25).T.reshape(-1)
        if not propto:
            lp += -(C - 2) * K * (HALF_LOG_2PI + math.log(2.5))

    # sigma_raw ~ normal(sigma_slope * intercept + sigma_intercept, sigma_sigma)  (:223)
    m = sigma_slope * intercept + sigma_intercept
    e = (sigma_raw - m) / sigma_sigma
    lp += -G * math.log(sigma_sigma) - 0.5 * math.fsum(e * e)
    if not propto:
        lp += -G * HALF_LOG_2PI
    g_sigma_raw += -e / sigma_sigma
    g_m = e / sigma_sigma
    g_intercept += sigma_slope * g_m
    g_sigma_slope += (g_m * intercept).sum()
    g_sigma_intercept += g_m.sum()
    g_sigma_sigma += -G / sigma_sigma + (e * e).sum() / sigma_sigma

    # ---- constraints / Jacobians (:183-197) ---------------------------------------------
    g[0] = g_lambda_mu
    g[1] = g_lambda_sigma * lambda_sigma
    g[2] = g_lambda_skew
    g[Lo.o_intercept:Lo.o_intercept + G] = g_intercept
    g[Lo.o_sigma_raw:Lo.o_sigma_raw + G] = g_sigma_raw
    g[Lo.o_tail] = g_sigma_slope * sigma_slope      # d sigma_slope/du = -exp(u) = sigma_slope
    g[Lo.o_tail + 1] = g_sigma_intercept
    g[Lo.o_tail + 2] = g_sigma_sigma * sigma_sigma
    if jacobian:
        lp += p["u_ls"] + p["u_ss"] + p["u_sg"]
        g[1] += 1.0
        g[Lo.o_tail] += 1.0
        g[Lo.o_tail + 2] += 1.0
    return float(lp), g


def constrain(d: ModelData, theta: np.ndarray) -> dict:
    """write_array for the quantities the R side consumes (alpha_sub_1, sigma_raw, lambda_log_param)."""
    p = unpack(np.asarray(theta, dtype=np.float64), d.G, d.K, d.C)
    alpha = alpha_matrix(p, d.G, d.K, d.C)
    return dict(alpha_sub_1=p["alpha_sub_1"].copy(), sigma_raw=p["sigma_raw"].copy(),
                lambda_log_param=(d.X @ alpha))     # [S, G] as in :205
