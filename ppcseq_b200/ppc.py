"""Posterior-predictive summaries and outlier flags (host-side mirror of fit_to_counts_rng /
check_if_within_posterior / add_deleterious_if_covariate_exists, reference R/utilities.R:685-703,
:651-663, :493-513)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int32_p, c_uint8_p, check


def _dp(a): return a.ctypes.data_as(c_double_p)


def summarise_draws(draws, p: float, device: int = 0):
    """draws [n_draws, n_pairs] (integer-valued) -> (lower, upper, mean, sd), R quantile type 7 at p, 1-p."""
    d = np.ascontiguousarray(draws, dtype=np.float64)
    if d.ndim != 2:
        raise ValueError("draws must be [n_draws, n_pairs]")
    n, m = d.shape
    out = [np.empty(m) for _ in range(4)]
    check(_lib.lib().ppcseq_summarise_draws(int(device), _dp(d), n, m, float(p), *[_dp(o) for o in out]))
    return tuple(out)


def flags(model, lower, upper, mean, slope=None):
    """[K,S] credible bounds + means (+ slope [K]) -> dict(ppc, deleterious, ppc_samples_failed, tot_deleterious_outliers)."""
    K, S = model.K, model.S
    lo, up, mu = (np.ascontiguousarray(a, dtype=np.float64).reshape(K, S) for a in (lower, upper, mean))
    ppc = np.empty((K, S), np.uint8)
    failed = np.empty(K, np.int32)
    has_cov = model.C > 1
    dele = np.empty((K, S), np.uint8) if has_cov else None
    tot = np.empty(K, np.int32) if has_cov else None
    sl = np.ascontiguousarray(slope, dtype=np.float64) if has_cov else None
    check(_lib.lib().ppcseq_flags(
        model.handle, _dp(lo), _dp(up), _dp(mu), _dp(sl) if has_cov else None,
        ppc.ctypes.data_as(c_uint8_p), dele.ctypes.data_as(c_uint8_p) if has_cov else None,
        failed.ctypes.data_as(c_int32_p), tot.ctypes.data_as(c_int32_p) if has_cov else None))
    return dict(ppc=ppc.astype(bool), deleterious=None if dele is None else dele.astype(bool),
                ppc_samples_failed=failed, tot_deleterious_outliers=tot)
