"""40-digit mpmath evaluation of the ppcseq Stan model -- the TRUTH the fp64 oracles and the CUDA
path are compared against.  ORACLE ONLY (see oracle/__init__.py; parity unpinned: the reference has
no golden vectors for log_prob/grad and cannot be executed in this image).

Semantics follow /root/reference/inst/stan/negBinomial_MPI.stan line by line:
  :183-197 constraints, :203-205 transformed parameters, :210-223 priors, :97-115 likelihood with
  the exclusion subtraction, :226 target += sum(map_rect(...)).
`log_prob` is a direct transcription (no algebra); `log_prob_grad` adds hand-derived partials, and
`numeric_grad` differentiates `log_prob` numerically at high precision so the partials can be
verified independently (tests/test_oracle.py).
"""
from __future__ import annotations

import mpmath as mp
import numpy as np

from .model_np import Layout, ModelData

mp.mp.dps = 40


def _unpack(theta, G, K, C):
    Lo = Layout(G, K, C)
    R = max(0, C - 2)
    th = [mp.mpf(float(t)) if not isinstance(t, mp.mpf) else t for t in theta]
    assert len(th) == Lo.D
    alpha_2 = [[th[Lo.o_alpha2 + k * R + r] for k in range(K)] for r in range(R)]   # [r][k]
    return Lo, dict(u_lm=th[0], u_ls=th[1], lambda_skew=th[2],
                    intercept=th[Lo.o_intercept:Lo.o_intercept + G],
                    alpha_sub_1=th[Lo.o_alpha1:Lo.o_alpha1 + K], alpha_2=alpha_2,
                    sigma_raw=th[Lo.o_sigma_raw:Lo.o_sigma_raw + G],
                    u_ss=th[Lo.o_tail], sigma_intercept=th[Lo.o_tail + 1], u_sg=th[Lo.o_tail + 2])


def _nb2_log_lpmf(n, eta, phi):
    """neg_binomial_2_log_lpmf(n | eta, phi) with all constants (it is called as a function, :98)."""
    mu = mp.e ** eta
    return (mp.loggamma(n + phi) - mp.loggamma(n + 1) - mp.loggamma(phi)
            + n * eta + phi * mp.log(phi) - (n + phi) * mp.log(mu + phi))


def log_prob(d: ModelData, theta, propto=True, jacobian=True, genes=None):
    """Direct transcription.  `genes` restricts the gene-level terms (for fast numeric partials)."""
    G, S, C, K = d.G, d.S, d.C, d.K
    Lo, p = _unpack(theta, G, K, C)
    L = mp.mpf(d.lambda_mu_mu)
    lambda_mu = p["u_lm"] + L
    lambda_sigma = mp.e ** p["u_ls"]
    lambda_skew = p["lambda_skew"]
    sigma_slope = -(mp.e ** p["u_ss"])
    sigma_intercept = p["sigma_intercept"]
    sigma_sigma = mp.e ** p["u_sg"]
    half_log_2pi = mp.log(2 * mp.pi) / 2
    X = [[mp.mpf(float(v)) for v in row] for row in d.X]
    ex = [mp.mpf(float(v)) for v in d.exposure]

    lp = mp.mpf(0)
    if genes is None:
        lp += -(lambda_mu - L) ** 2 / 8 - lambda_sigma ** 2 / 8 - lambda_skew ** 2 / 2
        lp += -sigma_intercept ** 2 / 8 - sigma_slope ** 2 / 8 - sigma_sigma ** 2 / 8
        if not propto:
            lp += 5 * (-half_log_2pi - mp.log(2)) - half_log_2pi
            lp += -G * half_log_2pi            # skew normal
            lp += -G * half_log_2pi            # sigma_raw normal
            if C >= 2:
                lp += -K * mp.log(2)
            if C >= 3:
                lp += -(C - 2) * K * (half_log_2pi + mp.log(mp.mpf("2.5")))
        if jacobian:
            lp += p["u_ls"] + p["u_ss"] + p["u_sg"]
    for g in (range(G) if genes is None else genes):
        ic = p["intercept"][g]
        sr = p["sigma_raw"][g]
        xi = lambda_mu + L                     # :219 -- lambda_mu_mu enters twice, as written
        z = (ic - xi) / lambda_sigma
        lp += -mp.log(lambda_sigma) - z * z / 2 + mp.log(mp.erfc(-lambda_skew * z / mp.sqrt(2)))
        m = sigma_slope * ic + sigma_intercept
        lp += -mp.log(sigma_sigma) - ((sr - m) / sigma_sigma) ** 2 / 2
        if g < K and C >= 2:
            lp += -abs(p["alpha_sub_1"][g])
        if g < K and C >= 3:
            for r in range(C - 2):
                lp += -p["alpha_2"][r][g] ** 2 / mp.mpf("12.5")
        phi = mp.e ** (-sr)
        for s in range(S):
            if d.exclude is not None and d.exclude[g, s]:
                continue                      # lpmf(all) - lpmf(excluded)  (:97-115)
            eta = ex[s] + X[s][0] * ic
            if g < K and C >= 2:
                eta += X[s][1] * p["alpha_sub_1"][g]
                for r in range(C - 2):
                    eta += X[s][2 + r] * p["alpha_2"][r][g]
            lp += _nb2_log_lpmf(int(d.counts[g, s]), eta, phi)
    return lp


def log_prob_grad(d: ModelData, theta, propto=True, jacobian=True):
    """lp (mpf) and gradient (list of mpf) with hand-derived partials, all in 40-digit arithmetic."""
    G, S, C, K = d.G, d.S, d.C, d.K
    Lo, p = _unpack(theta, G, K, C)
    R = max(0, C - 2)
    L = mp.mpf(d.lambda_mu_mu)
    lambda_mu = p["u_lm"] + L
    lambda_sigma = mp.e ** p["u_ls"]
    a = p["lambda_skew"]
    sigma_slope = -(mp.e ** p["u_ss"])
    sigma_intercept = p["sigma_intercept"]
    sigma_sigma = mp.e ** p["u_sg"]
    X = [[mp.mpf(float(v)) for v in row] for row in d.X]
    ex = [mp.mpf(float(v)) for v in d.exposure]
    lp = log_prob(d, theta, propto, jacobian)
    grad = [mp.mpf(0)] * Lo.D
    g_lm = -(lambda_mu - L) / 4
    g_ls = -lambda_sigma / 4
    g_sk = -a
    g_si = -sigma_intercept / 4
    g_ssl = -sigma_slope / 4
    g_ssg = -sigma_sigma / 4
    xi = lambda_mu + L
    sq2 = mp.sqrt(2)
    for g in range(G):
        ic = p["intercept"][g]
        sr = p["sigma_raw"][g]
        phi = mp.e ** (-sr)
        z = (ic - xi) / lambda_sigma
        t = -a * z / sq2
        ratio = mp.sqrt(2 / mp.pi) * mp.e ** (-t * t) / mp.erfc(t)      # sqrt(2/pi) exp(-a^2 z^2/2)/erfc(.)
        dz = -z + a * ratio
        g_ic = dz / lambda_sigma
        g_lm += -dz / lambda_sigma
        g_ls += (-1 - dz * z) / lambda_sigma
        g_sk += ratio * z
        m = sigma_slope * ic + sigma_intercept
        e = (sr - m) / sigma_sigma
        g_sr = -e / sigma_sigma
        g_m = e / sigma_sigma
        g_ic += sigma_slope * g_m
        g_ssl += g_m * ic
        g_si += g_m
        g_ssg += -1 / sigma_sigma + e * e / sigma_sigma
        d_alpha = [mp.mpf(0)] * C
        d_phi = mp.mpf(0)
        for s in range(S):
            if d.exclude is not None and d.exclude[g, s]:
                continue
            n = int(d.counts[g, s])
            eta = ex[s] + X[s][0] * ic
            if g < K and C >= 2:
                eta += X[s][1] * p["alpha_sub_1"][g]
                for r in range(R):
                    eta += X[s][2 + r] * p["alpha_2"][r][g]
            mu = mp.e ** eta
            de = n - (n + phi) * mu / (mu + phi)
            d_phi += ((mu - n) / (mu + phi) + mp.log(phi) - mp.log(mu + phi)
                      - mp.digamma(phi) + mp.digamma(n + phi))
            for c in range(C):
                d_alpha[c] += X[s][c] * de
        grad[Lo.o_intercept + g] = g_ic + d_alpha[0]
        grad[Lo.o_sigma_raw + g] = g_sr - phi * d_phi
        if g < K:
            if C >= 2:
                a1 = p["alpha_sub_1"][g]
                grad[Lo.o_alpha1 + g] = d_alpha[1] - mp.sign(a1)
            for r in range(R):
                grad[Lo.o_alpha2 + g * R + r] = d_alpha[2 + r] - p["alpha_2"][r][g] / mp.mpf("6.25")
    jac = 1 if jacobian else 0
    grad[0] = g_lm
    grad[1] = g_ls * lambda_sigma + jac
    grad[2] = g_sk
    grad[Lo.o_tail] = g_ssl * sigma_slope + jac
    grad[Lo.o_tail + 1] = g_si
    grad[Lo.o_tail + 2] = g_ssg * sigma_sigma + jac
    return lp, grad


def numeric_grad(d: ModelData, theta, idx, propto=True, jacobian=True, h=None):
    """Central difference of `log_prob` in the components `idx`, at 60 digits (independent check)."""
    old = mp.mp.dps
    mp.mp.dps = 60
    try:
        h = mp.mpf(10) ** (-20) if h is None else h
        Lo = Layout(d.G, d.K, d.C)
        th = [mp.mpf(float(t)) for t in theta]
        out = []
        for i in idx:
            genes = None
            if Lo.o_intercept <= i < Lo.o_intercept + d.G:
                genes = [i - Lo.o_intercept]
            elif Lo.o_alpha1 <= i < Lo.o_alpha1 + d.K:
                genes = [i - Lo.o_alpha1]
            elif Lo.o_alpha2 <= i < Lo.o_sigma_raw:
                genes = [(i - Lo.o_alpha2) // max(1, d.C - 2)]
            elif Lo.o_sigma_raw <= i < Lo.o_tail:
                genes = [i - Lo.o_sigma_raw]
            tp = list(th); tp[i] = th[i] + h
            tm = list(th); tm[i] = th[i] - h
            out.append((log_prob(d, tp, propto, jacobian, genes) - log_prob(d, tm, propto, jacobian, genes)) / (2 * h))
        return out
    finally:
        mp.mp.dps = old


def to_float(lp, grad):
    return float(lp), np.array([float(g) for g in grad])
