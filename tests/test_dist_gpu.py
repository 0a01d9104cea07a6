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


def _sampler_worker(rank, world, port, out):
    import torch.distributed as dist
    from ppcseq_b200 import NBModel, inference
    from ppcseq_b200 import dist as pdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nuts_golden.npz"))
    Gt, Kt = g["counts"].shape[0], int(g["K"])
    g0, g1 = pdist.shard_range(Gt, rank, world)
    m = NBModel(g["counts"][g0:g1], g["X"], g["exposure"], Kt, device=rank, shard=(Gt, g0))
    pdist.connect(m, rank, world, channels=1 + 4, cap=100)
    fit = inference.sample_nuts(m, chains=4, iter=150 + 600, warmup=150, seed=21)
    dr = fit.draws(0, m.D)
    vb = inference.advi(m, output_samples=1000, iter=20000, tol_rel_obj=0.005, seed=4)
    dv = vb.draws(0, m.D)
    assert not pdist.comm_timed_out(m)
    np.savez(out % rank, nuts=dr, advi=dv, info=fit.info(8), vinfo=vb.info(8))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(240)
def test_sharded_samplers_two_ranks(tmp_path, built_lib):
    """NUTS and ADVI with the genes split over 2 GPUs: both ranks must take bitwise the same decisions (identical
    hyper-parameter draws), and the assembled posterior must be concordant with the CPU oracle sampler."""
    import torch.multiprocessing as mp
    from oracle import model_np
    from ppcseq_b200 import dist as pdist
    out = str(tmp_path / "s%d.npz")
    mp.spawn(_sampler_worker, args=(2, 29350 + os.getpid() % 300, out), nprocs=2, join=True)
    r = [np.load(out % k) for k in range(2)]
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nuts_golden.npz"))
    Gt, Kt, C = g["counts"].shape[0], int(g["K"]), g["X"].shape[1]
    for key in ("nuts", "advi"):
        a, b = r[0][key], r[1][key]
        assert np.array_equal(a[:, :3], b[:, :3]) and np.array_equal(a[:, -3:], b[:, -3:]), key   # replicated hyper draws
    assert np.array_equal(r[0]["info"][[1, 3, 4, 7]], r[1]["info"][[1, 3, 4, 7]])           # same evals / divergences / tree sizes
    # assemble the global posterior mean from the shards and compare with the oracle sampler
    mean = np.zeros(model_np.dim(Gt, Kt, C))
    for k in range(2):
        g0, g1 = pdist.shard_range(Gt, k, 2)
        pdist.scatter_local_grad(mean, r[k]["nuts"].mean(axis=0), Gt, Kt, C, g0, g1, write_hyper=(k == 0))
    z = np.abs(mean - g["mean"]) / g["sd"]
    assert np.delete(z, len(z) - 1).max() < 0.35 and np.percentile(z, 90) < 0.2, float(z.max())
