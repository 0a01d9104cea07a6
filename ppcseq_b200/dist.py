"""Gene sharding across GPUs (one process per GPU).

The reference shards genes over map_rect workers (inst/stan/negBinomial_MPI.stan:226-240, packing in
R/utilities.R:125-174).  Here genes are block-partitioned over ranks; every gene-level parameter,
its counts row and its gradient live only on the owning rank, the 6 scalar hyper-parameters are
replicated, and one all-reduce(SUM) of 8 doubles per evaluation couples the ranks.
"""
from __future__ import annotations

import numpy as np

from .model import Layout


def shard_range(G: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced gene block of `rank` (first G % world ranks get one extra gene)."""
    base, extra = divmod(G, world)
    g0 = rank * base + min(rank, extra)
    return g0, g0 + base + (1 if rank < extra else 0)


def local_K(K: int, g0: int, g1: int) -> int:
    return max(0, min(g1 - g0, K - g0))


def local_theta(theta: np.ndarray, G: int, K: int, C: int, g0: int, g1: int) -> np.ndarray:
    """Slice a global unconstrained vector into the local vector of genes [g0, g1)."""
    lay = Layout(G, K, C)
    Gl, Kl = g1 - g0, local_K(K, g0, g1)
    ll = Layout(Gl, Kl, C)
    R = lay.R
    out = np.empty(ll.D)
    out[:3] = theta[:3]
    out[ll.o_intercept:ll.o_intercept + Gl] = theta[lay.o_intercept + g0:lay.o_intercept + g1]
    out[ll.o_alpha1:ll.o_alpha1 + Kl] = theta[lay.o_alpha1 + g0:lay.o_alpha1 + g0 + Kl]
    out[ll.o_alpha2:ll.o_alpha2 + R * Kl] = theta[lay.o_alpha2 + R * g0:lay.o_alpha2 + R * (g0 + Kl)]
    out[ll.o_sigma_raw:ll.o_sigma_raw + Gl] = theta[lay.o_sigma_raw + g0:lay.o_sigma_raw + g1]
    out[ll.o_tail:] = theta[lay.o_tail:]
    return out


def scatter_local_grad(grad_global: np.ndarray, grad_local: np.ndarray, G: int, K: int, C: int, g0: int, g1: int,
                       write_hyper: bool) -> None:
    """Inverse of `local_theta` for gradients: writes the gene block (and optionally the 6 hyper slots)."""
    lay = Layout(G, K, C)
    Gl, Kl = g1 - g0, local_K(K, g0, g1)
    ll = Layout(Gl, Kl, C)
    R = lay.R
    grad_global[lay.o_intercept + g0:lay.o_intercept + g1] = grad_local[ll.o_intercept:ll.o_intercept + Gl]
    grad_global[lay.o_alpha1 + g0:lay.o_alpha1 + g0 + Kl] = grad_local[ll.o_alpha1:ll.o_alpha1 + Kl]
    grad_global[lay.o_alpha2 + R * g0:lay.o_alpha2 + R * (g0 + Kl)] = grad_local[ll.o_alpha2:ll.o_alpha2 + R * Kl]
    grad_global[lay.o_sigma_raw + g0:lay.o_sigma_raw + g1] = grad_local[ll.o_sigma_raw:ll.o_sigma_raw + Gl]
    if write_hyper:
        grad_global[:3] = grad_local[:3]
        grad_global[lay.o_tail:] = grad_local[ll.o_tail:]
