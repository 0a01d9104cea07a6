"""The two sampler calls of the reference's do_inference() (R/utilities.R:1482-1513), GPU-resident.

  sample_nuts    <- rstan::sampling(stanmodels$negBinomial_MPI, chains, iter, warmup = 150, seed, init = "random", ...)
  advi           <- rstan::vb(model, output_samples, iter = 50000, tol_rel_obj = 0.005, ...)
  vb_iterative   <- the retry wrapper of R/utilities.R:246-278
  find_optimal_number_of_chains <- R/utilities.R:291-303
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _lib
from ._lib import EDIVERGED, AdviOpts, NutsOpts, PpcseqError, c_double_p, check
from .fit import Fit

WARMUP = 150          # R/utilities.R:1503


def find_optimal_number_of_chains(how_many_posterior_draws: int, max_number_to_check: int = 100, warmup: int = WARMUP) -> int:
    """argmin over c of draws/c + warmup*c, ties to the smaller c; the reference's tibble starts at c = 2
    (R/utilities.R:291-303)."""
    best, best_tot = None, None
    for c in range(2, max_number_to_check + 1):
        tot = how_many_posterior_draws / c + warmup * c
        if best is None or tot < best_tot:
            best, best_tot = c, tot
    return best


def sample_nuts(model, *, chains: int, iter: int, warmup: int = WARMUP, seed: int = 1, max_treedepth: int = 10,
                adapt_delta: float = 0.8, init=None, threads: int = 0) -> Fit:
    L = _lib.lib()
    o = NutsOpts()
    check(L.ppcseq_nuts_default_opts(ctypes.byref(o)))
    o.chains, o.iter, o.warmup, o.seed = int(chains), int(iter), int(warmup), int(seed) & (2**64 - 1)
    o.max_treedepth, o.adapt_delta, o.threads = int(max_treedepth), float(adapt_delta), int(threads)
    keep = None
    if init is not None:
        keep = np.ascontiguousarray(init, dtype=np.float64).reshape(chains, model.D)
        o.init = keep.ctypes.data_as(c_double_p)
    h = ctypes.c_void_p()
    check(L.ppcseq_sample_nuts(model.handle, ctypes.byref(o), ctypes.byref(h)))
    del keep
    return Fit(model, h)


def advi(model, *, output_samples: int, iter: int = 50000, tol_rel_obj: float = 0.005, seed: int = 1,
         grad_samples: int = 1, elbo_samples: int = 100, eval_elbo: int = 100, adapt_iter: int = 50, init=None) -> Fit:
    L = _lib.lib()
    o = AdviOpts()
    check(L.ppcseq_advi_default_opts(ctypes.byref(o)))
    o.iter, o.output_samples, o.tol_rel_obj, o.seed = int(iter), int(output_samples), float(tol_rel_obj), int(seed) & (2**64 - 1)
    o.grad_samples, o.elbo_samples, o.eval_elbo, o.adapt_iter = int(grad_samples), int(elbo_samples), int(eval_elbo), int(adapt_iter)
    keep = None
    if init is not None:
        keep = np.ascontiguousarray(init, dtype=np.float64).reshape(model.D)
        o.init = keep.ctypes.data_as(c_double_p)
    h = ctypes.c_void_p()
    check(L.ppcseq_advi_meanfield(model.handle, ctypes.byref(o), ctypes.byref(h)))
    del keep
    return Fit(model, h)


def vb_iterative(model, *, output_samples: int, iter: int, tol_rel_obj: float, seed: int = 1, max_attempts: int = 5) -> Fit:
    """Retry ADVI on failure, as vb_iterative does (R/utilities.R:246-278).  The reference's loop never
    terminates on a persistent failure (its counter is not incremented in scope); here it is bounded."""
    last = None
    for attempt in range(max_attempts):
        try:
            return advi(model, output_samples=output_samples, iter=iter, tol_rel_obj=tol_rel_obj, seed=seed + 7919 * attempt)
        except PpcseqError as e:
            last = e
            if e.rc != EDIVERGED:          # only "the algorithm failed" is retried, as rstan::vb errors are
                raise
    raise last
