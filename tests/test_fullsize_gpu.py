"""GPU parity at BASELINE.json's full sizes.

  (0) THE WHOLE PROBLEM AGAINST THE ORACLE, configs 2, 3, 4 and 5 (20k x 21, 60k x 500 with the pass-2 mask,
      20k x 2,000, 60k x 5,000): lp and the complete gradient of the CUDA path against the C oracle
      (oracle/ppcseq_oracle.c on all host cores -- ~0.1 s per evaluation at config 3, ~1 s at config 5) at the
      generating truth, at two points ~ U(-2,2)^D and at a point that forces the streaming fallback (phase B2 of
      the moment kernel) on ~10 % of the genes; 1e-10 with the scaling rule of SURVEY.md 7.2;
and, on config 3, size-independent properties:
  (1) two independent formulations agreeing: the per-element kernel (every count visited) against the
      Chebyshev-moment / Taylor-series kernel (data-only sufficient statistics), 1e-10;
  (2) additivity over gene shards: three shard models (partial sums, summed on the host in rank order, finalised)
      reproduce the unsharded lp and hyper-gradients, and their gene blocks are bitwise the unsharded ones;
  (3) the gradient being the derivative of lp: central differences along random directions;
  (4) a random 1,500-gene slice of the full problem against the C oracle (the oracle finishes that in seconds);
  (5) bitwise reproducibility.
"""
import ctypes

import numpy as np
import pytest

from tests.helpers import grad_err, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    from ppcseq_b200 import NBModel, synthetic
    w = synthetic.make("cfg3_60kx500")
    m = NBModel(w.counts, w.X, w.exposure, w.K)
    m.set_exclusion(w.exclude_pairs)
    return w, m


def _streaming_theta(w, seed=17):
    """theta_true with sigma_raw pushed to U(-7, -4) (phi = 55 .. 1100) on ~10 % of the genes: phi > 0.2 min{n >= 64}
    there, so those genes leave the data-only Taylor series and stream their count rows (phase B2)."""
    from ppcseq_b200 import layout
    rng = np.random.default_rng(seed)
    lay = layout(w.G, w.K, w.C)
    th = w.theta_true.copy()
    pick = rng.random(w.G) < 0.10
    th[lay.o_sigma_raw:lay.o_sigma_raw + w.G][pick] = rng.uniform(-7.0, -4.0, int(pick.sum()))
    return th


@pytest.mark.parametrize("name", ["cfg2_20kx21", "cfg3_60kx500", "cfg4_20kx2000", "cfg5_60kx5000"])
def test_whole_problem_against_the_oracle(name, built_lib):
    import os

    from oracle import c_oracle, model_np
    from ppcseq_b200 import NBModel, synthetic
    w = synthetic.make(name)
    excl = None
    if len(w.exclude_pairs):
        excl = np.zeros((w.G, w.S), bool)
        excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
    d = model_np.ModelData(w.counts, w.X, w.exposure, w.K, exclude=excl)
    m = NBModel(w.counts, w.X, w.exposure, w.K)
    if excl is not None:
        m.set_exclusion(w.exclude_pairs)
    ths = np.vstack([w.theta_true, synthetic.random_thetas(w, 2, seed=9), _streaming_theta(w)])
    lps, gs = m.log_prob_grad(ths)
    cores = os.cpu_count() or 1
    for i, th in enumerate(ths):
        lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=cores)
        assert np.isfinite(lp_ref)
        assert rel(lps[i], lp_ref) < 1e-10 and grad_err(gs[i], g_ref) < 1e-10, (name, i, rel(lps[i], lp_ref), grad_err(gs[i], g_ref))
    m.close()


def test_heavy_exclusion_at_scale_against_the_oracle(built_lib):
    """Config 4 (20,000 x 2,000) with 3 % of the points excluded (60 per gene -- what pass 2 looks like at thousands of
    samples): the excluded-point-moment mode of the kernel against the C oracle on the whole problem."""
    import os

    from oracle import c_oracle, model_np
    from ppcseq_b200 import NBModel, synthetic
    w = synthetic.make("cfg4_20kx2000")
    rng = np.random.default_rng(31)
    excl = rng.random((w.G, w.S)) < 0.03
    d = model_np.ModelData(w.counts, w.X, w.exposure, w.K, exclude=excl)
    m = NBModel(w.counts, w.X, w.exposure, w.K)
    m.set_exclusion(np.argwhere(excl))
    ths = np.vstack([w.theta_true, synthetic.random_thetas(w, 1, seed=3)])
    lps, gs = m.log_prob_grad(ths)
    for i, th in enumerate(ths):
        lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=os.cpu_count() or 1)
        assert rel(lps[i], lp_ref) < 1e-10 and grad_err(gs[i], g_ref) < 1e-10, (i, rel(lps[i], lp_ref), grad_err(gs[i], g_ref))
    m.close()


def test_two_formulations_agree(full, built_lib):
    from ppcseq_b200 import synthetic
    w, m = full
    ths = np.vstack([w.theta_true, synthetic.random_thetas(w, 2, seed=9)])
    m.set_design_path(2)
    lp2, g2 = m.log_prob_grad(ths)
    m.set_design_path(3)
    lp3, g3 = m.log_prob_grad(ths)
    m.set_design_path(0)
    for i in range(len(ths)):
        assert rel(lp3[i], lp2[i]) < 1e-10 and grad_err(g3[i], g2[i]) < 1e-10, i


def test_batched_launch_is_bitwise_the_single_launch(full, built_lib):
    """B thetas in one launch (grid.y = B, several waves of CTAs) against one launch per theta, repeated: guards the
    record ring (a TMA-fed variant of it once read the next batch in one warp out of ~10^5 for want of a proxy fence
    before the refill -- one wrong 32-gene tile in ~4 % of batched launches -- which only this test exposed) and the
    grid reduction's arrival protocol across waves."""
    from ppcseq_b200 import synthetic
    w, m = full
    B = 6
    ths = synthetic.random_thetas(w, B, seed=21)
    ths[0] = w.theta_true
    one = [m.log_prob_grad(ths[i]) for i in range(B)]
    for _ in range(10):
        lp, g = m.log_prob_grad(ths)
        for b in range(B):
            assert lp[b] == one[b][0] and np.array_equal(g[b], one[b][1]), b


def test_reproducible_bitwise(full, built_lib):
    from ppcseq_b200 import synthetic
    w, m = full
    th = synthetic.random_thetas(w, 1, seed=4)[0]
    a = m.log_prob_grad(th)
    b = m.log_prob_grad(th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_gradient_is_the_derivative_of_lp(full, built_lib):
    w, m = full
    rng = np.random.default_rng(2)
    th = w.theta_true + 0.01 * rng.standard_normal(w.D)
    lp0, g = m.log_prob_grad(th)
    for _ in range(3):
        v = rng.standard_normal(w.D)
        v /= np.linalg.norm(v)
        eps = 1e-4
        lps, _ = m.log_prob_grad(np.stack([th + eps * v, th - eps * v]))
        fd = (lps[0] - lps[1]) / (2 * eps)
        assert abs(fd - g @ v) <= 1e-6 * max(abs(g @ v), 1.0) + 1e-9 * abs(lp0) / eps * 1e-3, (fd, g @ v)


def test_shard_additivity(full, built_lib):
    from ppcseq_b200 import NBModel, _lib
    from ppcseq_b200 import dist as pdist
    from ppcseq_b200 import synthetic
    w, m = full
    L = _lib.lib()
    th = synthetic.random_thetas(w, 1, seed=6)[0]
    lp_full, g_full = m.log_prob_grad(th)
    excl = np.zeros((w.G, w.S), bool)
    excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
    W = 3
    total = np.zeros(8)
    g = np.zeros_like(th)
    last = None
    for r in range(W):
        g0, g1 = pdist.shard_range(w.G, r, W)
        ms = NBModel(w.counts[g0:g1], w.X, w.exposure, w.K, shard=(w.G, g0))
        ms.set_exclusion(np.argwhere(excl[g0:g1]))
        thl = np.ascontiguousarray(pdist.local_theta(th, w.G, w.K, w.C, g0, g1))
        nb = thl.nbytes
        d_th, d_gr, d_pt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        for p_, n_ in ((d_th, nb), (d_gr, nb), (d_pt, 64)):
            _lib.check(L.ppcseq_device_alloc(0, n_, ctypes.byref(p_)))
        _lib.check(L.ppcseq_memcpy_h2d(d_th, thl.ctypes.data_as(ctypes.c_void_p), nb, None))
        _lib.check(L.ppcseq_log_prob_grad_partial_device(ms.handle, 1, d_th, 1, d_pt, d_gr, None))
        _lib.check(L.ppcseq_stream_sync(ms.handle, None))
        part, gl = np.empty(8), np.empty_like(thl)
        _lib.check(L.ppcseq_memcpy_d2h(part.ctypes.data_as(ctypes.c_void_p), d_pt, 64, None))
        _lib.check(L.ppcseq_memcpy_d2h(gl.ctypes.data_as(ctypes.c_void_p), d_gr, nb, None))
        total += part
        pdist.scatter_local_grad(g, gl, w.G, w.K, w.C, g0, g1, write_hyper=False)
        last = (ms, d_th, d_gr, thl, nb)
        if r < W - 1:
            for p_ in (d_th, d_gr, d_pt):
                L.ppcseq_device_free(0, p_)
            ms.close()
    # finalise on the last shard's handle: hyper-priors, constraints, Jacobians from the summed partials
    ms, d_th, d_gr, thl, nb = last
    d_tot, d_lp = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.ppcseq_device_alloc(0, 64, ctypes.byref(d_tot)))
    _lib.check(L.ppcseq_device_alloc(0, 8, ctypes.byref(d_lp)))
    _lib.check(L.ppcseq_memcpy_h2d(d_tot, total.ctypes.data_as(ctypes.c_void_p), 64, None))
    _lib.check(L.ppcseq_finalize_hyper_device(ms.handle, 1, d_th, d_tot, 1, 1, d_lp, d_gr, None))
    _lib.check(L.ppcseq_stream_sync(ms.handle, None))
    lp, gl = np.empty(1), np.empty_like(thl)
    _lib.check(L.ppcseq_memcpy_d2h(lp.ctypes.data_as(ctypes.c_void_p), d_lp, 8, None))
    _lib.check(L.ppcseq_memcpy_d2h(gl.ctypes.data_as(ctypes.c_void_p), d_gr, nb, None))
    g[:3] = gl[:3]
    g[-3:] = gl[-3:]
    assert rel(lp[0], lp_full) < 1e-12
    lay = m.layout
    assert np.array_equal(g[3:lay.o_tail], g_full[3:lay.o_tail])            # gene blocks do not depend on the sharding
    assert np.allclose(g[:3], g_full[:3], rtol=1e-11, atol=0) and np.allclose(g[-3:], g_full[-3:], rtol=1e-11, atol=0)


def test_random_slice_against_the_oracle(full, built_lib):
    from oracle import c_oracle, model_np
    from ppcseq_b200 import NBModel
    w, m = full
    rng = np.random.default_rng(8)
    idx = np.sort(rng.choice(w.G, 1500, replace=False))
    excl = np.zeros((w.G, w.S), bool)
    excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
    d = model_np.ModelData(w.counts[idx], w.X, w.exposure, len(idx), exclude=excl[idx])
    ms = NBModel(d.counts, d.X, d.exposure, d.K)
    ms.set_exclusion(np.argwhere(d.exclude))
    th = rng.uniform(-2, 2, model_np.dim(len(idx), len(idx), w.C))
    lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=4)
    lp, g = ms.log_prob_grad(th)
    assert rel(lp, lp_ref) < 1e-10 and grad_err(g, g_ref) < 1e-10
