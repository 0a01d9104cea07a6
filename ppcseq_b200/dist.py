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


def connect(model, rank: int, world: int, channels: int = 1, cap: int = 8) -> None:
    """Fuse the cross-shard all-reduce into the log_prob kernel (ppcseq_comm_create / _connect): exchanges the
    64-byte IPC handles of the per-rank mailboxes through torch.distributed, then maps every peer's mailbox.
    After this, `model.log_prob_grad*` on the shard returns the global lp and hyper-gradients."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import _lib
    from ._lib import c_uint8_p, check

    L = _lib.lib()
    nb = 64
    mine = np.zeros(nb, np.uint8)
    check(L.ppcseq_comm_create(model.handle, rank, world, channels, cap, mine.ctypes.data_as(c_uint8_p)))
    if world == 1:
        allh = mine.copy()
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(mine).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        allh = np.concatenate([o.cpu().numpy() for o in out])
    check(L.ppcseq_comm_connect(model.handle, np.ascontiguousarray(allh).ctypes.data_as(c_uint8_p)))
    if world > 1:
        dist.barrier()


def comm_timed_out(model) -> bool:
    import ctypes

    from . import _lib
    from ._lib import check
    flag = ctypes.c_int32()
    check(_lib.lib().ppcseq_comm_status(model.handle, ctypes.byref(flag)))
    return bool(flag.value)
