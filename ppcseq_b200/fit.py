"""Fit handle: posterior draws resident in HBM (the stand-in for the stanfit object that
rstan::sampling / rstan::vb return, reference R/utilities.R:1482-1513) and the four queries the
reference makes on it (R/utilities.R:689-691, :738-743, :1252-1255, :796)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import c_double_p, check


def _dp(a): return a.ctypes.data_as(c_double_p)


class Fit:
    def __init__(self, model, handle):
        self.model = model
        self._h = handle
        n = ctypes.c_int32()
        check(_lib.lib().ppcseq_fit_num_draws(self._h, n))
        self.n_draws = n.value

    @classmethod
    def from_draws(cls, model, theta_draws):
        """theta_draws [n, D]: unconstrained draws the caller already holds."""
        th = np.ascontiguousarray(theta_draws, dtype=np.float64)
        if th.ndim != 2 or th.shape[1] != model.D:
            raise ValueError(f"theta_draws must be [n, {model.D}]")
        h = ctypes.c_void_p()
        check(_lib.lib().ppcseq_fit_from_draws(model.handle, _dp(th), th.shape[0], ctypes.byref(h)))
        return cls(model, h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().ppcseq_fit_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def draws(self, begin: int, count: int) -> np.ndarray:
        """[n_draws, count] draws of `count` consecutive unconstrained parameters."""
        out = np.empty((count, self.n_draws))
        check(_lib.lib().ppcseq_fit_get_draws(self._h, begin, count, _dp(out)))
        return out.T.copy()

    def param_mean(self, begin: int, count: int) -> np.ndarray:
        out = np.empty(count)
        check(_lib.lib().ppcseq_fit_param_mean(self._h, begin, count, _dp(out)))
        return out

    def info(self, n: int = 16) -> np.ndarray:
        out = np.zeros(n)
        check(_lib.lib().ppcseq_fit_info(self._h, _dp(out), n))
        return out

    def slope(self) -> np.ndarray:
        """posterior mean of alpha_sub_1[g], g < K (summary_to_tibble, R/utilities.R:1250-1263, :1531)."""
        m = self.model
        return self.param_mean(m.layout.o_alpha1, m.K)

    def ppc_summary(self, p: float, exact: bool = True, n_draws: int = 0, truncation_compensation: float = 1.0,
                    seed: int = 1):
        """(.lower, .upper, mean, sd), each [K, S]."""
        m = self.model
        out = [np.empty((m.K, m.S)) for _ in range(4)]
        check(_lib.lib().ppcseq_ppc_summary(self._h, int(exact), int(n_draws), float(p), float(truncation_compensation),
                                            int(seed), *[_dp(o) for o in out]))
        return tuple(out)

    def ppc_draws(self, truncation_compensation: float = 1.0, seed: int = 1) -> np.ndarray:
        """counts_rng [n_draws, K, S] (small problems only)."""
        m = self.model
        out = np.empty((self.n_draws, m.K * m.S))
        check(_lib.lib().ppcseq_ppc_draws(self._h, float(truncation_compensation), int(seed), _dp(out)))
        return out.reshape(self.n_draws, m.K, m.S)
