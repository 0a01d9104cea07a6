"""Small end-to-end pass over every kernel family (all likelihood paths, exclusion modes, exposure bins, samplers,
PPC, flags), sized for a run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/tools/sanitize.py
    compute-sanitizer --tool racecheck python profiles/tools/sanitize.py lp
(`lp` restricts the pass to the likelihood kernels.)  compute-sanitizer is closed on this pool's boxes (gpurun refuses
it), so in round 2 this only ran plain, as a smoke pass."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic

what = sys.argv[1] if len(sys.argv) > 1 else "all"

for (G, S, C) in [(700, 64, 3), (333, 40, 2), (50, 21, 1), (130, 300, 4)]:
    w = synthetic.make(G=G, S=S, C=C, mask=True, seed=5)
    m = P.NBModel(w.counts, w.X, w.exposure, w.K)
    m.set_exclusion(w.exclude_pairs)
    th = np.vstack([w.theta_true, synthetic.random_thetas(w, 2, seed=3)])
    th[0, m.layout.o_sigma_raw:m.layout.o_sigma_raw + G] -= 4.0        # large phi: some genes stream their rows
    for mode in (3, 2, 1):
        m.set_design_path(mode)
        lp, g = m.log_prob_grad(th)
        lp1, g1 = m.log_prob_grad(th[1])
        assert np.isfinite(lp).all()
    m.set_design_path(0)
    lp, _ = m.log_prob_grad(th)
    # heavy exclusion list: the excluded-point moment rows (xm mode) and back
    rng = np.random.default_rng(1)
    heavy = np.stack([rng.integers(0, G, 40 * G), rng.integers(0, S, 40 * G)], 1).astype(np.int32)
    m.set_exclusion(heavy)
    lp2, _ = m.log_prob_grad(th)
    m.set_exclusion(w.exclude_pairs)
    lp3, _ = m.log_prob_grad(th)
    assert np.array_equal(lp3, lp), (lp3, lp)
    eg = m.exposure_grad(th[1])
    assert np.isfinite(eg).all() and np.isfinite(lp2).all()
    print("ok lp", G, S, C, lp, flush=True)

# wide exposure range: several exposure bins per design row
w = synthetic.make(G=200, S=96, C=2, mask=False, seed=9)
ex = w.exposure + np.linspace(-2.5, 2.5, w.exposure.size)
m = P.NBModel(w.counts, w.X, ex, w.K)
lp, g = m.log_prob_grad(w.theta_true)
assert np.isfinite(lp).all()
print("ok bins", lp, flush=True)

if what == "all":
    from ppcseq_b200 import inference, ppc
    w = synthetic.make(G=120, S=24, C=2, mask=False, seed=3)
    m = P.NBModel(w.counts, w.X, w.exposure, w.K)
    for threads in (0, -1):
        fit = inference.sample_nuts(m, chains=2, iter=25, warmup=15, seed=4, threads=threads)
        print("ok nuts", threads, fit.info(4), flush=True)
    fit = inference.advi(m, output_samples=50, iter=300, tol_rel_obj=0.05, seed=2)
    print("ok advi", flush=True)
    lo, up, mean, sd = fit.ppc_summary(0.05, exact=True, seed=7)
    lo2, up2, mean2, sd2 = fit.ppc_summary(0.01, exact=False, n_draws=4000, seed=7)
    d = fit.ppc_draws(seed=7)
    s4 = ppc.summarise_draws(d.reshape(d.shape[0], -1), 0.05)
    fl = ppc.flags(m, lo, up, mean, fit.slope())
    print("ok ppc", lo.shape, d.shape, flush=True)
print("sanitize pass done")
