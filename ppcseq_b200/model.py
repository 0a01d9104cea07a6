"""Host-side handle of the GPU-resident model: the stand-in for `stanmodels$negBinomial_MPI`
(reference R/stanmodels.R:10-25) plus rstan's `log_prob` / `grad_log_prob` on it."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import c_double_p, c_int32_p, check


@dataclass(frozen=True)
class Layout:
    """Offsets into Stan's unconstrained vector (inst/stan/negBinomial_MPI.stan:180-199)."""
    G: int
    K: int
    C: int

    @property
    def R(self): return max(0, self.C - 2)
    @property
    def o_intercept(self): return 3
    @property
    def o_alpha1(self): return 3 + self.G
    @property
    def o_alpha2(self): return 3 + self.G + self.K
    @property
    def o_sigma_raw(self): return self.o_alpha2 + self.R * self.K
    @property
    def o_tail(self): return self.o_sigma_raw + self.G
    @property
    def D(self): return self.o_tail + 3


def layout(G: int, K: int, C: int) -> Layout:
    return Layout(G, K, C)


def _dp(a): return a.ctypes.data_as(c_double_p)
def _ip(a): return a.ctypes.data_as(c_int32_p)


class NBModel:
    """counts [G,S] int32 gene-major, X [S,C] model.matrix, exposure_rate [S], K = how_many_to_check."""

    def __init__(self, counts, X, exposure_rate, K, lambda_mu_mu=5.612671, device=0, shard=None, devices=None):
        """device: CUDA ordinal of a single-GPU model.  devices: list of ordinals -> ONE handle whose genes are split
        over those GPUs inside this process (ppcseq_model_create_multi; every method below then works on the global
        problem).  shard=(G_total, g_begin): this process's block of a model sharded over several processes."""
        L = _lib.lib()
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.float64)
        ex = np.ascontiguousarray(exposure_rate, dtype=np.float64)
        if counts.ndim != 2 or X.ndim != 2 or X.shape[0] != counts.shape[1] or ex.shape != (counts.shape[1],):
            raise ValueError("shape mismatch: counts [G,S], X [S,C], exposure_rate [S]")
        G, S = counts.shape
        C = X.shape[1]
        self._h = ctypes.c_void_p()
        self.device = device
        self.devices = None if devices is None else [int(d) for d in devices]
        if devices is not None:
            if shard is not None:
                raise ValueError("devices= (one process, several GPUs) and shard= (one process per GPU) are exclusive")
            dv = np.ascontiguousarray(self.devices, dtype=np.int32)
            check(L.ppcseq_model_create_multi(G, S, C, int(K), _ip(counts), _dp(X), _dp(ex), float(lambda_mu_mu),
                                              len(dv), _ip(dv), ctypes.byref(self._h)))
            self.device = self.devices[0]
        elif shard is None:
            check(L.ppcseq_model_create(G, S, C, int(K), _ip(counts), _dp(X), _dp(ex), float(lambda_mu_mu),
                                        int(device), ctypes.byref(self._h)))
        else:
            G_total, g_begin = shard
            check(L.ppcseq_model_create_shard(int(G_total), int(K), int(g_begin), int(g_begin) + G, S, C,
                                              _ip(counts), _dp(X), _dp(ex), float(lambda_mu_mu), int(device),
                                              ctypes.byref(self._h)))
        g, s, c, k, d = (ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64())
        check(L.ppcseq_model_dims(self._h, g, s, c, k, d))
        self.G, self.S, self.C, self.K, self.D = g.value, s.value, c.value, k.value, d.value
        self.layout = Layout(self.G, self.K, self.C)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().ppcseq_model_free(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_exclusion(self, pairs):
        """pairs: int array [n,2] of 0-based (g, s) dropped from the likelihood (pass 2)."""
        p = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        check(_lib.lib().ppcseq_model_set_exclusion(self._h, _ip(p), p.shape[0]))

    def set_design_path(self, mode: int):
        check(_lib.lib().ppcseq_model_set_design_path(self._h, int(mode)))

    def log_prob_grad(self, theta, propto=True, jacobian=True, out=None):
        """theta [D] or [B,D] (host) -> (lp, grad) with matching leading shape.

        B > 1 runs as a three-stage pipeline inside the library (theta b+1 host->device, evaluation b, gradient b-1
        device->host); pass page-locked arrays for theta and `out=(lp[B], grad[B,D])` to let the copies overlap."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        single = th.ndim == 1
        th2 = th.reshape(1, -1) if single else th
        if th2.shape[1] != self.D:
            raise ValueError(f"theta has {th2.shape[1]} columns, model dimension is {self.D}")
        B = th2.shape[0]
        if out is not None:
            lp, grad = out
            if (lp.dtype != np.float64 or grad.dtype != np.float64 or not lp.flags.c_contiguous or not grad.flags.c_contiguous
                    or lp.size != B or grad.size != th2.size):
                raise ValueError("out must be C-contiguous float64 arrays of B and B x D elements")
        else:
            lp = np.empty(B)
            grad = np.empty_like(th2)
        check(_lib.lib().ppcseq_log_prob_grad(self._h, B, _dp(th2), int(propto), int(jacobian), _dp(lp), _dp(grad)))
        return (float(lp.reshape(-1)[0]), grad.reshape(B, -1)[0]) if single else (lp, grad)

    def exposure_grad(self, theta):
        """OPTIONAL, outside every parity claim: d log_prob / d exposure_rate[s], an S-vector (exposure is data in the
        reference; BASELINE config 5's "exposure-gradient")."""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.shape != (self.D,):
            raise ValueError(f"theta must have {self.D} entries")
        out = np.empty(self.S)
        check(_lib.lib().ppcseq_exposure_grad(self._h, _dp(th), _dp(out)))
        return out
