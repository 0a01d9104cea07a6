"""A LITERAL, un-optimised transcription of the reference's Stan program, gradients by reverse-mode autodiff.
ORACLE-OF-THE-ORACLE: TEST INFRASTRUCTURE ONLY (imported by tests/ alone).

Purpose: pin oracle/model_mp.py, oracle/model_np.py and oracle/ppcseq_oracle.c with something they did not write.
Those three share one hand-derived algebra (dense gene-major arrays, exclusion as a 0/1 weight, closed-form
gradients).  This file shares none of it:

  * data go through the reference's own map_rect PACKING -- `format_for_MPI` (R/utilities.R:125-174: genes dealt
    cyclically to shards, rows ordered by G inside a shard), `counts_package` (R/utilities.R:1455-1466:
    [M, N, S, G_per_shard, symbol_end(M+1), sample_idx(N), counts(N), n_exclude, to_exclude...]) and
    `get_outlier_data_to_exlude` (R/utilities.R:321-359: shard-local 1-based row numbers);
  * `lp_reduce` (inst/stan/negBinomial_MPI.stan:58-120) unpacks that integer package with the Stan program's own
    index arithmetic, and the excluded points are SUBTRACTED (:105-115), not masked;
  * `get_reference_parameters_MPI` (:32-56) builds the per-shard parameter vectors (column-major lambda block,
    sigma block, zero buffer, exposure) exactly as written;
  * constrained parameters, `~` statements and the Jacobian terms follow Stan's documented transforms
    (lower = 0: exp, upper = 0: -exp, offset: add) literally; densities are written in their textbook form
    (lgamma / log-sum-exp / log_ndtr -- log erfc(-a z / sqrt 2) = log 2 + log Phi(a z)) and, in `log_prob_scipy`, taken from scipy.stats
    (nbinom, skewnorm, norm, laplace) with no formula of this repo at all;
  * the gradient is torch.autograd on that graph -- the mechanism Stan itself uses -- not a derived formula.

What this does NOT pin (stated in oracle/__init__.py): the bit-level behaviour of Stan Math's
neg_binomial_2_log_lpmf for phi > 1e5 in StanHeaders <= 2.21, edgeR's TMM, R's quantile().
fp64 lgamma differences lose digits at large counts, so comparisons against this file use
moderate counts and a 1e-9 / 1e-7 (lp / gradient) tolerance; the 1e-10 claims rest on mpmath.
"""
from __future__ import annotations

import math

import numpy as np


# ---------------------------------------------------------------------------------------------------------
# R side: format_for_MPI + counts_package + to_exclude_MPI  (1-based everywhere, as in R)
# ---------------------------------------------------------------------------------------------------------
def pack_for_map_rect(counts, exclude, shards):
    """counts [G, S] (gene g = G index g+1, sample s = S index s+1; every gene's rows in S order, as the
    reference assumes), exclude bool [G, S] or None.  Returns a dict with the Stan data objects the
    likelihood reads: M, N, n_shards, G_per_shard, G_ind [n_shards][M], counts_package [n_shards][CP]."""
    G, S = counts.shape
    # idx_MPI = head(rep(1:shards, ceiling(G / shards)), G): gene with G index g goes to shard ((g-1) %% shards) + 1
    idx_mpi = [(g % shards) + 1 for g in range(G)]
    shard_ids = sorted(set(idx_mpi))
    n_shards = min(shards, len(shard_ids))
    genes_of = {s: [g + 1 for g in range(G) if idx_mpi[g] == s] for s in shard_ids}       # arranged by G
    G_per_shard = [len(genes_of[s]) for s in shard_ids]
    M = max(G_per_shard)
    N = max(len(genes_of[s]) * S for s in shard_ids)
    rows_excl = []
    for s in shard_ids:
        ex = []
        row = 0
        for g1 in genes_of[s]:
            for s1 in range(1, S + 1):
                row += 1                                              # read_count_MPI_row
                if exclude is not None and exclude[g1 - 1, s1 - 1]:
                    ex.append(row)
        rows_excl.append(ex)
    max_ex = max(1, max(len(e) for e in rows_excl))                   # a dummy row when a shard has none
    package, G_ind = [], []
    for k, s in enumerate(shard_ids):
        genes = genes_of[s]
        symbol_end = [0]
        for _ in genes:
            symbol_end.append(symbol_end[-1] + S)
        symbol_end += [0] * (M + 1 - len(symbol_end))                 # replace(is.na(.), 0)
        sample_idx, cnt = [], []
        for g1 in genes:
            for s1 in range(1, S + 1):
                sample_idx.append(s1)
                cnt.append(int(counts[g1 - 1, s1 - 1]))
        sample_idx += [0] * (N - len(sample_idx))
        cnt += [0] * (N - len(cnt))
        ex = rows_excl[k]
        if exclude is not None and exclude.any():
            tail = [len(ex)] + ex + [0] * (max_ex - len(ex))
        else:
            tail = [0]                                                # matrix(rep(0, shards))
        package.append([M, N, S, len(genes)] + symbol_end + sample_idx + cnt + tail)
        G_ind.append(genes + [0] * (M - len(genes)))
    return dict(M=M, N=N, S=S, G=G, n_shards=n_shards, G_per_shard=G_per_shard, G_ind=G_ind, counts_package=package)


# ---------------------------------------------------------------------------------------------------------
# Stan side
# ---------------------------------------------------------------------------------------------------------
def _neg_binomial_2_log_lpmf(torch, n, eta, phi):
    """sum of the NB2-log pmf, textbook form with log-sum-exp:
       lchoose(n + phi - 1, n) + n (eta - lse) + phi (log phi - lse),  lse = log(exp(eta) + phi)."""
    lse = torch.logaddexp(eta, torch.log(phi))
    lchoose = torch.lgamma(n + phi) - torch.lgamma(n + 1.0) - torch.lgamma(phi)
    return torch.sum(lchoose + n * (eta - lse) + phi * (torch.log(phi) - lse))


def _lp_reduce(torch, local_parameters, int_data):
    """negBinomial_MPI.stan:58-120, 1-based slices turned into Python slices one by one."""
    M, N, S, Gps = int_data[0], int_data[1], int_data[2], int_data[3]
    symbol_end = int_data[4:4 + 1 + M]                                  # int_data[(4+1):(4+1+M)]
    sample_idx = int_data[4 + 1 + M:4 + 1 + M + N]                      # int_data[(4+1+M+1):(4+1+M+1+N-1)]
    counts = int_data[4 + 1 + M + N:4 + 1 + M + N + N]                  # int_data[(4+1+M+1+N-1+1):(4+1+M+1+N-1+N)]
    size_exclude = int_data[4 + 1 + M + N + N]                          # int_data[(4+1+M+1+N-1+N+1)]
    to_exclude = int_data[4 + 1 + M + N + N + 1:4 + 1 + M + N + N + 1 + size_exclude]
    lambda_MPI = local_parameters[0:Gps * S]
    sigma_MPI = local_parameters[Gps * S:Gps * S + Gps]
    exposure_rate = local_parameters[M * S + M:]
    n_rows = symbol_end[Gps]                                            # symbol_end[G_per_shard+1]
    sigma_c = torch.cat([sigma_MPI[g].repeat(symbol_end[g + 1] - symbol_end[g]) for g in range(Gps)])
    idx = torch.tensor([i - 1 for i in sample_idx[:n_rows]], dtype=torch.long)
    cnt = torch.tensor(counts[:n_rows], dtype=torch.float64)
    eta = exposure_rate[idx] + lambda_MPI
    lp = _neg_binomial_2_log_lpmf(torch, cnt, eta, sigma_c)
    if size_exclude > 0:
        ex = torch.tensor([i - 1 for i in to_exclude], dtype=torch.long)
        lp = lp - _neg_binomial_2_log_lpmf(torch, cnt[ex], eta[ex], sigma_c[ex])
    return lp


def _get_reference_parameters_MPI(torch, pk, lambda_log, sigma, exposure_rate):
    """:32-56.  lambda_log is [S, G]; to_vector of a column subset is column-major = gene by gene."""
    M, S = pk["M"], pk["S"]
    out = []
    for i in range(pk["n_shards"]):
        gi = [g - 1 for g in pk["G_ind"][i][:pk["G_per_shard"][i]]]
        size_buffer = (M * S + M) - (len(gi) * S + len(gi))
        out.append(torch.cat([lambda_log[:, gi].T.reshape(-1), sigma[gi],
                              torch.zeros(size_buffer, dtype=torch.float64), exposure_rate]))
    return out


def _normal_lpdf(torch, y, mu, sigma):
    return torch.sum(-0.5 * math.log(2.0 * math.pi) - torch.log(sigma) - 0.5 * ((y - mu) / sigma) ** 2)


def _skew_normal_lpdf(torch, y, xi, omega, alpha):
    z = (y - xi) / omega
    # log erfc(-alpha z / sqrt 2) = log(2 Phi(alpha z)); the library's log_ndtr keeps it finite where erfc underflows
    return torch.sum(-0.5 * math.log(2.0 * math.pi) - torch.log(omega) - 0.5 * z * z
                     + math.log(2.0) + torch.special.log_ndtr(alpha * z))


def _double_exponential_lpdf(torch, y, mu, sigma):
    return torch.sum(-math.log(2.0) - torch.log(sigma) - torch.abs(y - mu) / sigma)


def log_prob_grad(counts, X, exposure, K, theta, lambda_mu_mu=5.612671, exclude=None, jacobian=True, shards=3):
    """log_prob<propto = false, jacobian>(theta) with every constant, and its gradient by autograd.
    theta in Stan's declaration order (:180-199)."""
    import torch
    counts = np.asarray(counts)
    G, S = counts.shape
    C = X.shape[1]
    R = max(0, C - 2)
    pk = pack_for_map_rect(counts, exclude, shards)
    t64 = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=torch.float64)
    th = t64(theta).clone().requires_grad_(True)
    one = t64(1.0)
    # ---- parameters block: unconstrained -> constrained, with the log-Jacobians Stan adds --------------------
    o = 0
    lambda_mu = th[o] + lambda_mu_mu; o += 1                            # real<offset = lambda_mu_mu>
    lambda_sigma = torch.exp(th[o]); jac = th[o]; o += 1                # real<lower = 0>
    lambda_skew = th[o]; o += 1
    intercept = th[o:o + G]; o += G
    alpha_sub_1 = th[o:o + K]; o += K
    alpha_2 = th[o:o + R * K].reshape(K, R).T if R else th[o:o].reshape(0, K); o += R * K   # matrix[R, K], column-major
    sigma_raw = th[o:o + G]; o += G
    sigma_slope = -torch.exp(th[o]); jac = jac + th[o]; o += 1          # real<upper = 0>
    sigma_intercept = th[o]; o += 1
    sigma_sigma = torch.exp(th[o]); jac = jac + th[o]; o += 1           # real<lower = 0>
    assert o == th.shape[0]
    # ---- transformed parameters (:200-206) ------------------------------------------------------------------------
    sigma = 1.0 / torch.exp(sigma_raw)
    if C == 1:                                                          # merge_coefficients (:122-139)
        alpha = intercept.reshape(1, G)
    else:
        row2 = torch.cat([alpha_sub_1, torch.zeros(G - K, dtype=torch.float64)]).reshape(1, G)
        rest = torch.cat([alpha_2, torch.zeros((R, G - K), dtype=torch.float64)], dim=1)
        alpha = torch.cat([intercept.reshape(1, G), row2, rest], dim=0)
    lambda_log_param = t64(X) @ alpha                                   # [S, G]
    # ---- model block (:208-241) -----------------------------------------------------------------------------------
    target = _normal_lpdf(torch, lambda_mu, t64(lambda_mu_mu), 2.0 * one)
    target = target + _normal_lpdf(torch, lambda_sigma, 0.0 * one, 2.0 * one)
    target = target + _normal_lpdf(torch, lambda_skew, 0.0 * one, one)
    target = target + _normal_lpdf(torch, sigma_intercept, 0.0 * one, 2.0 * one)
    target = target + _normal_lpdf(torch, sigma_slope, 0.0 * one, 2.0 * one)
    target = target + _normal_lpdf(torch, sigma_sigma, 0.0 * one, 2.0 * one)
    target = target + _skew_normal_lpdf(torch, intercept, lambda_mu + lambda_mu_mu, lambda_sigma, lambda_skew)
    if C >= 2:
        target = target + _double_exponential_lpdf(torch, alpha_sub_1, 0.0 * one, one)
    if C >= 3:
        target = target + _normal_lpdf(torch, alpha_2.T.reshape(-1), 0.0 * one, 2.5 * one)
    target = target + _normal_lpdf(torch, sigma_raw, sigma_slope * intercept + sigma_intercept, sigma_sigma)
    shards_par = _get_reference_parameters_MPI(torch, pk, lambda_log_param, sigma, t64(exposure))
    for i in range(pk["n_shards"]):                                     # sum(map_rect(lp_reduce, ...))
        target = target + _lp_reduce(torch, shards_par[i], pk["counts_package"][i])
    if jacobian:
        target = target + jac
    (g,) = torch.autograd.grad(target, th)
    return float(target.detach()), g.detach().numpy()


def log_prob_scipy(counts, X, exposure, K, theta, lambda_mu_mu=5.612671, exclude=None, jacobian=True):
    """The same log-density with EVERY density taken from scipy.stats (no lgamma / erfc formula of this repo)."""
    from scipy import stats
    counts = np.asarray(counts)
    G, S = counts.shape
    C = X.shape[1]
    R = max(0, C - 2)
    th = np.asarray(theta, dtype=np.float64)
    o = 0
    lambda_mu = th[o] + lambda_mu_mu; o += 1
    lambda_sigma = math.exp(th[o]); jac = th[o]; o += 1
    lambda_skew = th[o]; o += 1
    intercept = th[o:o + G]; o += G
    alpha_sub_1 = th[o:o + K]; o += K
    alpha_2 = th[o:o + R * K].reshape(K, R).T; o += R * K
    sigma_raw = th[o:o + G]; o += G
    sigma_slope = -math.exp(th[o]); jac += th[o]; o += 1
    sigma_intercept = th[o]; o += 1
    sigma_sigma = math.exp(th[o]); jac += th[o]; o += 1
    sigma = 1.0 / np.exp(sigma_raw)
    alpha = np.zeros((C, G))
    alpha[0] = intercept
    if C >= 2:
        alpha[1, :K] = alpha_sub_1
    if C >= 3:
        alpha[2:, :K] = alpha_2
    mu = np.exp(X @ alpha + np.asarray(exposure)[:, None]).T            # [G, S]
    lp = stats.norm.logpdf(lambda_mu, lambda_mu_mu, 2) + stats.norm.logpdf(lambda_sigma, 0, 2)
    lp += stats.norm.logpdf(lambda_skew, 0, 1) + stats.norm.logpdf(sigma_intercept, 0, 2)
    lp += stats.norm.logpdf(sigma_slope, 0, 2) + stats.norm.logpdf(sigma_sigma, 0, 2)
    lp += stats.skewnorm.logpdf(intercept, lambda_skew, loc=lambda_mu + lambda_mu_mu, scale=lambda_sigma).sum()
    if C >= 2:
        lp += stats.laplace.logpdf(alpha_sub_1, 0, 1).sum()
    if C >= 3:
        lp += stats.norm.logpdf(alpha_2, 0, 2.5).sum()
    lp += stats.norm.logpdf(sigma_raw, sigma_slope * intercept + sigma_intercept, sigma_sigma).sum()
    ll = stats.nbinom.logpmf(counts, sigma[:, None], sigma[:, None] / (sigma[:, None] + mu))
    lp += ll.sum()
    if exclude is not None:
        lp -= ll[exclude].sum()                                         # the subtraction of :105-115
    return float(lp + (jac if jacobian else 0.0))
