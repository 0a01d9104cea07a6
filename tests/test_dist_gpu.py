"""GPU, 2 ranks on 2 GPUs (skipped on a single-GPU box): gene shards with the all-reduce fused into the log_prob
kernel (peer mailboxes over NVLink) against the unsharded C oracle; every rank must hold bitwise the same lp and
hyper-gradients."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G, S, C, K = 600, 64, 3, 300


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import model_np
    from ppcseq_b200 import NBModel
    from ppcseq_b200 import dist as pdist
    from tests.helpers import small_problem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = small_problem(G, S, C, K, seed=23, exclude_frac=0.02, big=True)
    thetas = np.random.default_rng(5).uniform(-2, 2, (3, model_np.dim(G, K, C)))
    g0, g1 = pdist.shard_range(G, rank, world)
    m = NBModel(d.counts[g0:g1], d.X, d.exposure, K, device=rank, shard=(G, g0))
    m.set_exclusion(np.argwhere(d.exclude[g0:g1]))
    pdist.connect(m, rank, world, channels=1, cap=4)
    res = []
    for rep in range(2):                                 # two rounds: both mailbox parities
        for th in thetas:
            lp, gl = m.log_prob_grad(pdist.local_theta(th, G, K, C, g0, g1))
            res.append((lp, gl))
    lps, gls = m.log_prob_grad(np.stack([pdist.local_theta(th, G, K, C, g0, g1) for th in thetas]))   # batched
    assert not pdist.comm_timed_out(m)
    np.savez(out % rank, lp=np.array([r[0] for r in res]), grads=np.stack([r[1] for r in res]), lpb=lps, gb=gls)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_allreduce_two_ranks(tmp_path, built_lib):
    import torch.multiprocessing as mp
    from oracle import c_oracle, model_np
    from ppcseq_b200 import dist as pdist
    from tests.helpers import grad_err, rel, small_problem
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_worker, args=(2, 29650 + os.getpid() % 300, out), nprocs=2, join=True)
    r = [np.load(out % k) for k in range(2)]
    d = small_problem(G, S, C, K, seed=23, exclude_frac=0.02, big=True)
    thetas = np.random.default_rng(5).uniform(-2, 2, (3, model_np.dim(G, K, C)))
    assert np.array_equal(r[0]["lp"], r[1]["lp"])                        # bitwise identical on both ranks
    assert np.array_equal(r[0]["grads"][:, :3], r[1]["grads"][:, :3]) and np.array_equal(r[0]["grads"][:, -3:], r[1]["grads"][:, -3:])
    assert np.array_equal(r[0]["lp"][:3], r[0]["lp"][3:]) and np.array_equal(r[0]["lpb"], r[0]["lp"][:3])
    for i, th in enumerate(thetas):
        lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=2)
        g = np.zeros_like(th)
        for k in range(2):
            g0, g1 = pdist.shard_range(G, k, 2)
            pdist.scatter_local_grad(g, r[k]["grads"][i], G, K, C, g0, g1, write_hyper=(k == 0))
        assert rel(r[0]["lp"][i], lp_ref) < 1e-10 and grad_err(g, g_ref) < 1e-10
