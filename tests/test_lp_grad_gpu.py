"""GPU parity: the CUDA log_prob/grad (through the C ABI) against the oracle.

Tolerance (north_star: 1e-9 relative in fp64): |lp - ref| <= 1e-9 |ref| and, per gradient
component, |g_i - ref_i| <= 1e-9 max(|ref_i|, 1e-3 ||ref||_inf)  (SURVEY.md 7.2).  The asserts
below use 1e-10 so there is a 10x margin.
"""
import numpy as np
import pytest

from oracle import c_oracle, model_mp, model_np
from tests.helpers import grad_err, rel, small_problem

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _model(d, **kw):
    from ppcseq_b200 import NBModel
    m = NBModel(d.counts, d.X, d.exposure, d.K, lambda_mu_mu=d.lambda_mu_mu, **kw)
    if d.exclude is not None:
        m.set_exclusion(np.argwhere(d.exclude))
    return m


CASES = [  # G, S, C, K, exclude_frac, continuous
    (5, 6, 3, 3, 0.1, False),
    (7, 4, 1, 2, 0.0, False),
    (33, 21, 2, 33, 0.05, False),
    (70, 45, 4, 40, 0.1, True),
    (40, 70, 8, 17, 0.02, True),
    (64, 33, 3, 0, 0.0, False),
]


@pytest.mark.parametrize("G,S,C,K,ef,cont", CASES)
@pytest.mark.parametrize("propto,jac", [(True, True), (False, True), (True, False)])
def test_against_mpmath_truth(G, S, C, K, ef, cont, propto, jac, built_lib):
    d = small_problem(G, S, C, K, seed=G * 100 + S, exclude_frac=ef, big=True, continuous=cont)
    th = np.random.default_rng(7).uniform(-2, 2, model_np.dim(G, K, C))
    if G * S <= 1000:
        lp_ref, g_ref = model_mp.to_float(*model_mp.log_prob_grad(d, th, propto, jac))
    else:
        lp_ref, g_ref = model_np.log_prob_grad(d, th, propto, jac)
    m = _model(d)
    for mode in ([1, 2, 3] if not cont else [1]):   # general, per-element categorical, Chebyshev-moment categorical
        m.set_design_path(mode)
        lp, g = m.log_prob_grad(th, propto, jac)
        assert rel(lp, lp_ref) < TOL, (mode, lp, lp_ref)
        assert grad_err(g, g_ref) < TOL, mode


def test_adversarial_values(built_lib):
    G, S, C, K = 6, 4, 2, 6
    d = small_problem(G, S, C, K, seed=11, big=True)
    lay = model_np.Layout(G, K, C)
    th = np.zeros(lay.D)
    th[lay.o_intercept:lay.o_intercept + G] = [-20, 20, 0, 5, -20, 20]
    th[lay.o_sigma_raw:lay.o_sigma_raw + G] = [np.log(1e3), np.log(1e3), -np.log(1e5), -np.log(1e5), 0, 0]
    lp_ref, g_ref = model_mp.to_float(*model_mp.log_prob_grad(d, th))
    m = _model(d)
    for mode in (1, 2, 3):
        m.set_design_path(mode)
        lp, g = m.log_prob_grad(th)
        assert rel(lp, lp_ref) < 1e-9 and grad_err(g, g_ref) < 1e-9, mode


def test_batch_and_determinism(built_lib):
    d = small_problem(300, 64, 3, 150, seed=5, exclude_frac=0.01)
    ths = np.random.default_rng(3).uniform(-2, 2, (5, model_np.dim(300, 150, 3)))
    m = _model(d)
    lps, gs = m.log_prob_grad(ths)
    for b in range(5):
        lp_ref, g_ref = c_oracle.log_prob_grad(d, ths[b])
        assert rel(lps[b], lp_ref) < TOL and grad_err(gs[b], g_ref) < TOL
    lps2, gs2 = m.log_prob_grad(ths)
    assert np.array_equal(lps, lps2) and np.array_equal(gs, gs2)      # bitwise reproducible


def test_medium_synthetic_config(built_lib):
    """A slice of BASELINE config 3's generator (2,000 x 500, C = 3, pass-2 mask) against the C oracle."""
    from ppcseq_b200 import synthetic
    w = synthetic.make(G=2000, S=500, C=3, mask=True, seed=20242)
    excl = np.zeros((w.G, w.S), bool)
    excl[w.exclude_pairs[:, 0], w.exclude_pairs[:, 1]] = True
    d = model_np.ModelData(w.counts, w.X, w.exposure, w.K, exclude=excl)
    m = _model(d)
    for th in (w.theta_true, synthetic.random_thetas(w, 1)[0]):
        lp_ref, g_ref = c_oracle.log_prob_grad(d, th, n_shards=4)
        for mode in (2, 3):
            m.set_design_path(mode)
            lp, g = m.log_prob_grad(th)
            assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL, mode


def test_exclusion_roundtrip(built_lib):
    """set_exclusion(n) then clearing it restores the pass-1 value bit for bit."""
    d = small_problem(50, 21, 2, 50, seed=9)
    th = np.random.default_rng(2).uniform(-2, 2, model_np.dim(50, 50, 2))
    m = _model(d)
    lp0, g0 = m.log_prob_grad(th)
    m.set_exclusion([[3, 4], [10, 20], [49, 0]])
    lp1, _ = m.log_prob_grad(th)
    assert lp1 != lp0
    m.set_exclusion(np.empty((0, 2), np.int32))
    lp2, g2 = m.log_prob_grad(th)
    assert lp2 == lp0 and np.array_equal(g0, g2)


def test_heavy_exclusion_lists(built_lib):
    """30 % of the points excluded, every pair listed twice and in random order, genes with all, all-but-one and none
    of their samples excluded; theta with large phi so that some genes stream their rows: the moment path's
    per-gene correction list (two lanes taking alternate points) against the C oracle and the per-element path."""
    rng = np.random.default_rng(31)
    G, S, C, K = 75, 90, 3, 40
    d = small_problem(G, S, C, K, seed=77, exclude_frac=0.3, big=True)
    d.exclude[3, :] = True
    d.exclude[4, :] = True; d.exclude[4, 17] = False
    d.exclude[5, :] = False
    d2 = model_np.ModelData(d.counts, d.X, d.exposure, d.K, exclude=d.exclude)
    pairs = np.argwhere(d.exclude)
    pairs = np.vstack([pairs, pairs])[rng.permutation(2 * len(pairs))]
    from ppcseq_b200 import NBModel
    m = NBModel(d.counts, d.X, d.exposure, d.K, lambda_mu_mu=d.lambda_mu_mu)
    m.set_exclusion(pairs)
    lay = model_np.Layout(G, K, C)
    for big_phi in (False, True):
        th = rng.uniform(-2, 2, lay.D)
        if big_phi:
            th[lay.o_sigma_raw:lay.o_sigma_raw + G] = rng.uniform(-6, -3, G)        # phi = 20 .. 400
        lp_ref, g_ref = c_oracle.log_prob_grad(d2, th)
        for mode in (2, 3):
            m.set_design_path(mode)
            lp, g = m.log_prob_grad(th)
            assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL, (mode, big_phi)


@pytest.mark.parametrize("C,S,spread", [(2, 30, 3.0), (3, 200, 2.5), (2, 500, 1.0), (1, 64, 4.0), (4, 96, 2.0)])
def test_wide_exposure_range_uses_exposure_bins(C, S, spread, built_lib):
    """Exposure from TMM on arbitrary libraries is unbounded (R/methods.R:222-238; pseudobulk spans 10-100x and more):
    the moment path cuts every design row into exposure bins with their own Chebyshev centre / half-width (piecewise
    series), so exposure rates spanning e^-spread..e^spread (ratios 7 .. 3000) stay on the data-only path, with an
    exclusion list, against the oracle."""
    G, K = 45, 20
    d = small_problem(G, S, C, K, seed=6 + S, exclude_frac=0.03, big=True)
    rng = np.random.default_rng(S)
    d.exposure[:] = rng.permutation(np.linspace(-spread, spread, S))
    d2 = model_np.ModelData(d.counts, d.X, d.exposure, d.K, exclude=d.exclude)
    m = _model(d2)
    m.set_design_path(3)                                 # must be available
    for seed in (4, 5):
        th = np.random.default_rng(seed).uniform(-2, 2, model_np.dim(G, K, C))
        lp_ref, g_ref = c_oracle.log_prob_grad(d2, th)
        lp, g = m.log_prob_grad(th)
        assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL, (C, S, spread, seed)
    m.set_design_path(2)                                 # the per-element path sees the same (exposure-sorted) layout
    lp, g = m.log_prob_grad(th)
    assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL


def test_absurd_exposure_range_falls_back(built_lib):
    """e^-12..e^12: even 8 bins per design row would need more than 48 terms -- auto uses the per-element path."""
    from ppcseq_b200 import PpcseqError
    d = small_problem(20, 30, 2, 10, seed=6)
    d.exposure[:] = np.linspace(-12.0, 12.0, 30)
    d.counts[:] = np.minimum(d.counts, 50)               # keep exp(eta) finite at the far end
    th = np.random.default_rng(4).uniform(-1, 1, model_np.dim(20, 10, 2))
    lp_ref, g_ref = model_np.log_prob_grad(d, th)
    m = _model(d)
    lp, g = m.log_prob_grad(th)
    assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL
    with pytest.raises(PpcseqError):
        m.set_design_path(3)


def test_moment_path_shapes(built_lib):
    """Design-row counts 1, 2, 3, 5, 8 (lanes per gene 1, 2, 4, 8, 8), small-count-only genes, gene counts that do
    not fill a warp tile, and an exclusion list, against the C oracle."""
    rng = np.random.default_rng(12)
    for C, G, S in [(1, 37, 40), (2, 70, 64), (3, 33, 50), (4, 21, 90), (4, 9, 300)]:
        d = small_problem(G, S, C, G // 2, seed=C * 7 + G, exclude_frac=0.03, big=True)
        if C == 4:                                  # 5 or 8 distinct rows instead of 2^(C-1)
            d.X[:, 3] = d.X[:, 1] * d.X[:, 2] if G == 21 else d.X[:, 3]
        d.counts[5 % G, :] = rng.integers(0, 31, S)                 # a gene with small counts only
        d2 = model_np.ModelData(d.counts, d.X, d.exposure, d.K, exclude=d.exclude)
        th = rng.uniform(-2, 2, model_np.dim(G, d.K, C))
        lp_ref, g_ref = c_oracle.log_prob_grad(d2, th)
        m = _model(d2)
        m.set_design_path(3)
        lp, g = m.log_prob_grad(th)
        assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL, (C, G, S)


def test_bad_arguments(built_lib):
    from ppcseq_b200 import NBModel, PpcseqError
    d = small_problem(5, 6, 2, 3, seed=1)
    with pytest.raises(PpcseqError):
        NBModel(-np.abs(d.counts) - 1, d.X, d.exposure, d.K)
    with pytest.raises(PpcseqError):
        NBModel(d.counts, d.X, d.exposure, 99)
    m = _model(d)
    with pytest.raises(ValueError):
        m.log_prob_grad(np.zeros(3))
    with pytest.raises(PpcseqError):
        m.set_exclusion([[99, 0]])


def test_ring_depth_variants_are_bitwise_identical(built_lib):
    """The moment kernel runs with a deep record ring when the launch leaves the SMs mostly empty (small gene shards)
    and with the two-stage ring otherwise.  G = 2,000 genes = 32 CTAs: a single theta takes the deep ring, a batch of 12
    thetas (384 CTAs) the shallow one -- same arithmetic, so lp and every gradient entry must be bitwise the same."""
    G, S, C, K = 2000, 64, 3, 1200
    d = small_problem(G, S, C, K, seed=77, exclude_frac=0.02, big=True)
    m = _model(d)
    m.set_design_path(3)
    ths = np.random.default_rng(2).uniform(-2, 2, (12, model_np.dim(G, K, C)))
    lpb, gb = m.log_prob_grad(ths)          # host batches are pipelined as single-theta launches: use the device entry
    import ctypes
    from ppcseq_b200 import _lib
    L = _lib.lib()
    nb = ths.nbytes
    d_th, d_gr, d_lp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    for p_, n_ in ((d_th, nb), (d_gr, nb), (d_lp, 8 * 12)):
        _lib.check(L.ppcseq_device_alloc(0, n_, ctypes.byref(p_)))
    _lib.check(L.ppcseq_memcpy_h2d(d_th, ths.ctypes.data_as(ctypes.c_void_p), nb, None))
    _lib.check(L.ppcseq_log_prob_grad_device(m.handle, 12, d_th, 1, 1, d_lp, d_gr, None))      # ONE launch, grid.y = 12
    _lib.check(L.ppcseq_stream_sync(m.handle, None))
    lp12, g12 = np.empty(12), np.empty_like(ths)
    _lib.check(L.ppcseq_memcpy_d2h(lp12.ctypes.data_as(ctypes.c_void_p), d_lp, 96, None))
    _lib.check(L.ppcseq_memcpy_d2h(g12.ctypes.data_as(ctypes.c_void_p), d_gr, nb, None))
    for p_ in (d_th, d_gr, d_lp):
        L.ppcseq_device_free(0, p_)
    assert np.array_equal(lp12, lpb) and np.array_equal(g12, gb)
    lp_ref, g_ref = c_oracle.log_prob_grad(d, ths[3], n_shards=2)
    assert rel(lp12[3], lp_ref) < TOL and grad_err(g12[3], g_ref) < TOL


@pytest.mark.parametrize("C,continuous", [(1, False), (3, False), (2, True)])
def test_optional_exposure_gradient_output(C, continuous, built_lib):
    """OPTIONAL output (outside every parity claim: exposure_rate is data in the reference,
    negBinomial_MPI.stan:167-168): d log_prob / d exposure_rate[s].  Checked against central differences of the C
    oracle's log_prob in the exposure vector and against the closed form; categorical and general designs, with an
    exclusion list; the multi-GPU parent handle returns the same vector."""
    from ppcseq_b200 import NBModel
    G, S, K = 150, 70, 60
    d = small_problem(G, S, C, K, seed=5 + C, exclude_frac=0.04, big=True, continuous=continuous)
    th = np.random.default_rng(3).uniform(-1.5, 1.5, model_np.dim(G, K, C))
    m = _model(d)
    got = m.exposure_grad(th)
    # closed form: sum_g w phi (n - mu) / (mu + phi)
    p = model_np.unpack(th, G, K, C)
    alpha = model_np.alpha_matrix(p, G, K, C)
    mu = np.exp((d.X @ alpha).T + d.exposure[None, :])
    phi = np.exp(-p["sigma_raw"])[:, None]
    w = 1.0 if d.exclude is None else ~d.exclude
    ref = (w * phi * (d.counts - mu) / (mu + phi)).sum(axis=0)
    assert np.allclose(got, ref, rtol=1e-11, atol=1e-9 * np.abs(ref).max())
    # central differences of the oracle's lp in exposure_s
    for s in (0, S // 2, S - 1):
        e = 1e-5
        lps = []
        for sign in (1, -1):
            ex = d.exposure.copy()
            ex[s] += sign * e
            lps.append(c_oracle.log_prob_grad(model_np.ModelData(d.counts, d.X, ex, d.K, exclude=d.exclude), th)[0])
        fd = (lps[0] - lps[1]) / (2 * e)
        assert abs(fd - got[s]) <= 2e-6 * max(abs(got[s]), 1.0) + 1e-6 * abs(lps[0]) * 1e-6 / e, (s, fd, got[s])
    if not continuous:
        m.set_design_path(1)                             # the general formulation of the same kernel
        assert np.allclose(m.exposure_grad(th), ref, rtol=1e-11, atol=1e-9 * np.abs(ref).max())
    mul = NBModel(d.counts, d.X, d.exposure, d.K, devices=[0])
    mul.set_exclusion(np.argwhere(d.exclude))
    assert np.allclose(mul.exposure_grad(th), ref, rtol=1e-11, atol=1e-9 * np.abs(ref).max())


@pytest.mark.parametrize("C,S,spread", [(2, 300, 0.3), (3, 260, 2.0), (1, 200, 0.2)])
def test_heavy_exclusion_lists_use_excluded_point_moments(C, S, spread, built_lib):
    """More than 24 excluded points per gene on average (pass 2 at S = 5,000 excludes > 100 per gene): the moment kernel
    takes the excluded points off through per-gene T_j moments stored next to the count moments instead of a per-point
    list.  Parity against the oracle in that mode, and through the switches heavy -> light -> none -> heavy (the record
    changes size each time); with exposure bins as well."""
    G, K = 50, 30
    rng = np.random.default_rng(S + C)
    d = small_problem(G, S, C, K, seed=40 + C, big=True)
    d.exposure[:] = rng.permutation(np.linspace(-spread, spread, S))
    heavy = rng.random((G, S)) < 0.3
    heavy[3, :] = True                                    # a fully excluded gene
    heavy[4, 1:] = True                                   # ... and one with a single point left
    light = rng.random((G, S)) < 0.02
    m = _model(model_np.ModelData(d.counts, d.X, d.exposure, d.K))
    m.set_design_path(3)
    th = rng.uniform(-2, 2, model_np.dim(G, K, C))
    for excl in (heavy, light, None, heavy):
        m.set_exclusion(np.empty((0, 2), np.int32) if excl is None else np.argwhere(excl))
        dd = model_np.ModelData(d.counts, d.X, d.exposure, d.K, exclude=excl)
        lp_ref, g_ref = c_oracle.log_prob_grad(dd, th)
        lp, g = m.log_prob_grad(th)
        assert rel(lp, lp_ref) < TOL and grad_err(g, g_ref) < TOL, (C, S, None if excl is None else int(excl.sum()))
        m.set_design_path(2)                              # the per-element path on the same handle agrees
        lp2, g2 = m.log_prob_grad(th)
        m.set_design_path(3)
        assert rel(lp2, lp_ref) < TOL and grad_err(g2, g_ref) < TOL
