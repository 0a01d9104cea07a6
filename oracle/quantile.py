"""R `quantile` type 7, mean/sd and the outlier flags.  ORACLE ONLY (parity unpinned: R is absent,
so bit-exactness is defined w.r.t. exactly the expression and operation order written here; the definition is
cross-checked against NumPy / SciPy / pandas type-7 quantiles in tests/test_oracle_pin.py).

Follows /root/reference/R/utilities.R:
  :689-691, :770-776  quantile(c(p, 1-p)), mean, sd over the draws of one (sample, gene) pair
  :659-661            ppc = between(count, .lower, .upper) (inclusive); is higher than mean
  :502-510            is_group_right, is group high, deleterious_outliers
  :597, :604          ppc_samples_failed, tot_deleterious_outliers
R's quantile.default(type=7) *(external)*: index = 1 + (n-1)*p; lo = floor(index); hi = ceiling(index);
x = sort(x, partial = unique(c(lo, hi))); qs = x[lo]; h = index - lo;
qs = (1 - h)*qs + h*x[hi] wherever index > lo and x[hi] != qs.
"""
from __future__ import annotations

import math

import numpy as np


def quantile_type7_sorted(x_sorted: np.ndarray, p: float) -> float:
    n = len(x_sorted)
    index = 1.0 + (n - 1) * p                 # fp64, this order
    lo = math.floor(index)
    hi = math.ceil(index)
    q = float(x_sorted[lo - 1])
    if index > lo and float(x_sorted[hi - 1]) != q:
        h = index - lo
        q = (1.0 - h) * q + h * float(x_sorted[hi - 1])     # two products, one add, no FMA
    return q


def summarise_draws(draws: np.ndarray, p: float):
    """draws [n_draws, n_pairs] (integer-valued) -> lower, upper, mean, sd per pair.

    `.lower` is the type-7 quantile at p, `.upper` at (1 - p) with the subtraction done in fp64
    (R/utilities.R:691).  mean is the exact integer sum divided by n; sd uses n-1.
    """
    n, m = draws.shape
    xs = np.sort(np.asarray(draws, dtype=np.float64), axis=0)
    p_hi = 1.0 - p
    lower = np.array([quantile_type7_sorted(xs[:, j], p) for j in range(m)])
    upper = np.array([quantile_type7_sorted(xs[:, j], p_hi) for j in range(m)])
    mean = np.empty(m)
    sd = np.empty(m)
    for j in range(m):
        col = [int(v) for v in draws[:, j]]
        s1 = sum(col)
        s2 = sum(v * v for v in col)
        mean[j] = s1 / n                                   # correctly rounded exact quotient
        num = n * s2 - s1 * s1                             # exact integer, >= 0
        sd[j] = math.sqrt(float(num) / float(n * (n - 1))) if n > 1 else float("nan")
    return lower, upper, mean, sd


def flags(counts: np.ndarray, lower, upper, mean, slope, X: np.ndarray):
    """counts [K,S] int, lower/upper/mean [K,S], slope [K] (posterior mean of alpha_sub_1), X [S,C].

    Returns dict(ppc [K,S] bool, deleterious [K,S] bool or None, ppc_samples_failed [K],
    tot_deleterious_outliers [K] or None).
    """
    c = counts.astype(np.float64)
    ppc = (c >= lower) & (c <= upper)                      # dplyr::between is inclusive
    higher = (~ppc) & (c > mean)
    out = dict(ppc=ppc, is_higher_than_mean=higher, ppc_samples_failed=(~ppc).sum(axis=1).astype(np.int32),
               deleterious=None, tot_deleterious_outliers=None)
    if X.shape[1] > 1:
        col = X[:, 1]
        right = col > col.mean()                           # R/utilities.R:502
        sl = np.asarray(slope, dtype=np.float64)[:, None]
        group_high = ((sl > 0) & right[None, :]) | ((sl < 0) & ~right[None, :])
        dele = (~ppc) & (higher == group_high)
        out["deleterious"] = dele
        out["tot_deleterious_outliers"] = dele.sum(axis=1).astype(np.int32)
    return out
