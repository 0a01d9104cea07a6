"""CPU, world_size 2 over gloo: the gene-sharding host logic (ppcseq_b200/dist.py) and the additive structure the
multi-GPU path relies on -- the sum over shards of the shard-local results (gene blocks final, hyper parts summed)
reproduces the unsharded oracle.  No GPU, no product kernels: each rank evaluates its shard with the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model_np
from ppcseq_b200 import dist as pdist
from tests.helpers import small_problem

G, S, C, K = 23, 9, 3, 10


def _hyper_only(d, theta):
    """log_prob and gradient of a zero-gene model = the 6 hyper-priors + Jacobians (what every shard repeats)."""
    d0 = model_np.ModelData(d.counts[:0], d.X, d.exposure, 0)
    th0 = np.concatenate([theta[:3], theta[-3:]])
    return model_np.log_prob_grad(d0, th0)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = small_problem(G, S, C, K, seed=17, exclude_frac=0.05, big=True)
    theta = np.random.default_rng(3).uniform(-2, 2, model_np.dim(G, K, C))
    g0, g1 = pdist.shard_range(G, rank, world)
    Kl = pdist.local_K(K, g0, g1)
    dl = model_np.ModelData(d.counts[g0:g1], d.X, d.exposure, Kl, exclude=d.exclude[g0:g1])
    thl = pdist.local_theta(theta, G, K, C, g0, g1)
    lp_l, g_l = model_np.log_prob_grad(dl, thl)
    lp_h, g_h = _hyper_only(d, theta)
    # all-reduce(SUM) of [lp, 6 hyper-gradients] with the repeated hyper-prior part removed on ranks > 0
    part = np.concatenate([[lp_l], g_l[:3], g_l[-3:]])
    if rank > 0:
        part -= np.concatenate([[lp_h], g_h])
    t = torch.from_numpy(part)
    dist.all_reduce(t)
    # gene blocks: gather the local gradients into the global layout
    gg = np.zeros_like(theta)
    pdist.scatter_local_grad(gg, g_l, G, K, C, g0, g1, write_hyper=False)
    tg = torch.from_numpy(gg)
    dist.all_reduce(tg)
    gg = tg.numpy()
    gg[:3] = t.numpy()[1:4]
    gg[-3:] = t.numpy()[4:7]
    if rank == 0:
        np.savez(out, lp=t.numpy()[0], grad=gg)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sum_equals_unsharded_oracle(world, tmp_path):
    out = str(tmp_path / "res.npz")
    port = 29500 + os.getpid() % 500 + world
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r = np.load(out)
    d = small_problem(G, S, C, K, seed=17, exclude_frac=0.05, big=True)
    theta = np.random.default_rng(3).uniform(-2, 2, model_np.dim(G, K, C))
    lp, g = model_np.log_prob_grad(d, theta)
    assert abs(r["lp"] - lp) <= 1e-12 * abs(lp)
    assert np.allclose(r["grad"], g, rtol=1e-11, atol=1e-11)


def test_shard_ranges_partition_the_genes():
    for Gt, W in [(10, 1), (10, 3), (7, 8), (60000, 8)]:
        r = [pdist.shard_range(Gt, k, W) for k in range(W)]
        assert r[0][0] == 0 and r[-1][1] == Gt and all(r[i][1] == r[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
    assert pdist.local_K(5, 0, 4) == 4 and pdist.local_K(5, 4, 8) == 1 and pdist.local_K(5, 8, 12) == 0


def test_local_theta_roundtrip():
    rng = np.random.default_rng(0)
    theta = rng.normal(size=model_np.dim(G, K, C))
    rebuilt = np.zeros_like(theta)
    for rank in range(3):
        g0, g1 = pdist.shard_range(G, rank, 3)
        thl = pdist.local_theta(theta, G, K, C, g0, g1)
        assert thl.shape == (model_np.dim(g1 - g0, pdist.local_K(K, g0, g1), C),)
        pdist.scatter_local_grad(rebuilt, thl, G, K, C, g0, g1, write_hyper=(rank == 0))
    assert np.array_equal(rebuilt, theta)
