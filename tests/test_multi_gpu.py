"""GPU: ONE process, SEVERAL GPUs -- the handle of ppcseq_model_create_multi (what a single R process calling
identify_outliers() would hold; reference: one DLL / one process, R/stanmodels.R:10-25, src/RcppExports.cpp:15-25,
likelihood sharded inside it by map_rect, inst/stan/negBinomial_MPI.stan:226-240).

  * devices = [0] (runs on any box): the parent machinery (global <-> local index algebra, shard thread, fit
    queries) must reproduce the plain single-device handle bit for bit;
  * devices = [0, 1] (skipped on a single-GPU box): lp / hyper-gradients against the oracle and the unsharded GPU
    result, gene-block gradients bitwise, PPC bitwise (Philox streams keyed by the global pair), samplers and
    identify_outliers() end to end.
"""
import os

import numpy as np
import pytest
import torch

from tests.helpers import grad_err, rel, small_problem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
two_gpus = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
G, S, C, K = 601, 64, 3, 301          # odd sizes: unequal shards, K ends inside the first shard


def _problem():
    from oracle import model_np
    d = small_problem(G, S, C, K, seed=23, exclude_frac=0.02, big=True)
    thetas = np.random.default_rng(5).uniform(-2, 2, (3, model_np.dim(G, K, C)))
    return d, thetas


def _check_against_single(devices):
    from oracle import c_oracle
    from ppcseq_b200 import Fit, NBModel
    from ppcseq_b200 import ppc as P
    d, thetas = _problem()
    one = NBModel(d.counts, d.X, d.exposure, K, device=0)
    mul = NBModel(d.counts, d.X, d.exposure, K, devices=devices)
    assert (mul.G, mul.S, mul.C, mul.K, mul.D) == (one.G, one.S, one.C, one.K, one.D)
    pairs = np.argwhere(d.exclude)
    one.set_exclusion(pairs)
    mul.set_exclusion(pairs[::-1])                       # order of the list must not matter
    lay = one.layout
    for rep in range(2):                                 # both mailbox parities
        for th in thetas:
            lp1, g1 = one.log_prob_grad(th)
            lpm, gm = mul.log_prob_grad(th)
            lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=2)
            assert rel(lpm, lp_ref) < 1e-10 and grad_err(gm, g_ref) < 1e-10
            assert np.array_equal(gm[3:lay.o_tail], g1[3:lay.o_tail])            # gene blocks: bitwise the unsharded ones
            if len(devices) == 1:
                assert lpm == lp1 and np.array_equal(gm, g1)
            else:
                assert rel(lpm, lp1) < 1e-13
    lpb, gb = mul.log_prob_grad(thetas)                  # batched (3-stage pipeline inside every shard)
    for i, th in enumerate(thetas):
        lpm, gm = mul.log_prob_grad(th)
        assert lpb[i] == lpm and np.array_equal(gb[i], gm)
    for mode in (2, 1, 0):                               # the other likelihood paths through the parent
        mul.set_design_path(mode); one.set_design_path(mode)
        lpm, gm = mul.log_prob_grad(thetas[0]); lp1, g1 = one.log_prob_grad(thetas[0])
        assert rel(lpm, lp1) < 1e-13 and np.array_equal(gm[3:lay.o_tail], g1[3:lay.o_tail])
    # fit queries + PPC + flags on imported draws: bitwise the single-device results
    draws = thetas[0][None, :] * 0.2 + 0.05 * np.random.default_rng(1).standard_normal((96, one.D))
    draws[:, lay.o_intercept:lay.o_intercept + G] += 4.0
    f1, fm = Fit.from_draws(one, draws), Fit.from_draws(mul, draws)
    assert fm.n_draws == 96
    assert np.array_equal(fm.draws(0, one.D), f1.draws(0, one.D)) and np.array_equal(fm.draws(0, one.D), draws)
    assert np.array_equal(fm.draws(lay.o_alpha1 + 7, 400), f1.draws(lay.o_alpha1 + 7, 400))   # a range across blocks and shards
    assert np.array_equal(fm.param_mean(0, one.D), f1.param_mean(0, one.D))
    assert np.array_equal(fm.slope(), f1.slope())
    for exact, nd in ((True, 0), (False, 500)):
        s1 = f1.ppc_summary(0.05, exact=exact, n_draws=nd, truncation_compensation=0.7352941, seed=9)
        sm = fm.ppc_summary(0.05, exact=exact, n_draws=nd, truncation_compensation=0.7352941, seed=9)
        for a, b in zip(s1, sm):
            assert np.array_equal(a, b)
    assert np.array_equal(fm.ppc_draws(seed=4), f1.ppc_draws(seed=4))
    fl1 = P.flags(one, s1[0], s1[1], s1[2], f1.slope())
    flm = P.flags(mul, sm[0], sm[1], sm[2], fm.slope())
    for k in fl1:
        assert np.array_equal(fl1[k], flm[k]), k
    assert fl1["ppc_samples_failed"].sum() > 0
    mul.set_exclusion(np.empty((0, 2), np.int32)); one.set_exclusion(np.empty((0, 2), np.int32))
    lpm, gm = mul.log_prob_grad(thetas[1]); lp1, g1 = one.log_prob_grad(thetas[1])
    assert rel(lpm, lp1) < 1e-13 and np.array_equal(gm[3:lay.o_tail], g1[3:lay.o_tail])
    for h in (f1, fm):
        h.close()
    one.close(); mul.close()


def test_parent_with_one_device_is_the_plain_handle(built_lib):
    _check_against_single([0])


@two_gpus
def test_two_devices_one_process(built_lib):
    _check_against_single([0, 1])
    _check_against_single([1, 0])                        # device order is the caller's choice


def test_parent_rejects_device_pointer_entry_points(built_lib):
    import ctypes

    from ppcseq_b200 import NBModel, PpcseqError, _lib
    d, _ = _problem()
    mul = NBModel(d.counts, d.X, d.exposure, K, devices=[0])
    L = _lib.lib()
    with pytest.raises(PpcseqError) as e:
        _lib.check(L.ppcseq_log_prob_grad_device(mul.handle, 1, None, 1, 1, None, None, None))
    assert e.value.rc == 4                                # PPCSEQ_ESTATE
    buf = (ctypes.c_uint8 * 64)()
    with pytest.raises(PpcseqError):
        _lib.check(L.ppcseq_comm_create(mul.handle, 0, 1, 1, 1, buf))
    with pytest.raises(PpcseqError):
        NBModel(d.counts, d.X, d.exposure, K, devices=[0, 0])
    with pytest.raises(PpcseqError):
        NBModel(d.counts, d.X, d.exposure, K, devices=[99])


@two_gpus
@pytest.mark.timeout(300)
def test_samplers_two_devices_one_process(built_lib):
    """NUTS and ADVI on a 2-GPU handle: the assembled posterior is concordant with the CPU oracle sampler (as the
    single-GPU and the 2-process tests require), reproducible for a fixed seed, and the hyper-parameter draws are
    those of every shard."""
    from ppcseq_b200 import NBModel, inference
    g = np.load(os.path.join(GOLD, "nuts_golden.npz"))
    m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]), devices=[0, 1])
    fit = inference.sample_nuts(m, chains=4, iter=150 + 600, warmup=150, seed=21)
    assert fit.n_draws == 2400 and fit.info(8)[0] == 1
    dr = fit.draws(0, m.D)
    z = np.abs(dr.mean(axis=0) - g["mean"]) / g["sd"]
    assert np.delete(z, len(z) - 1).max() < 0.35 and np.percentile(z, 90) < 0.2, float(z.max())
    again = inference.sample_nuts(m, chains=4, iter=150 + 600, warmup=150, seed=21).draws(0, m.D)
    assert np.array_equal(dr, again)
    vb = inference.advi(m, output_samples=1000, iter=20000, tol_rel_obj=0.005, seed=4)
    lay = m.layout
    dv = vb.draws(lay.o_intercept, m.G)
    ref_m, ref_s = g["mean"][lay.o_intercept:lay.o_intercept + m.G], g["sd"][lay.o_intercept:lay.o_intercept + m.G]
    assert (np.abs(dv.mean(axis=0) - ref_m) / ref_s).max() < 1.0
    # more chains than the default mailbox geometry provides: the parent re-wires its mailboxes
    f12 = inference.sample_nuts(m, chains=12, iter=40, warmup=20, seed=2)
    assert f12.n_draws == 240 and np.isfinite(f12.draws(0, m.D)).all()
    m.close()


@two_gpus
@pytest.mark.parametrize("vb", [True, False])
def test_identify_outliers_two_devices(vb, built_lib):
    """tests/testthat/test-ppcSeq.R:26-30 through identify_outliers(devices = [0, 1]): c(0, 1, 0)."""
    from ppcseq_b200.api import identify_outliers
    from tests.test_inference_gpu import _tidy
    z, df = _tidy("bundled_test53.npz")
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value", significance="PValue",
                            do_check="is_significant", percent_false_positive_genes=1, approximate_posterior_inference=vb,
                            how_many_negative_controls=50, cores=4, seed=7, devices=[0, 1])
    assert list(res["symbol"]) == ["SLC16A12", "CYP1A1", "ART3"]
    assert list(res["tot_deleterious_outliers"].astype(int)) == [0, 1, 0]
