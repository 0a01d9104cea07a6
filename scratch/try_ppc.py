import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic, Fit
w = synthetic.make(G=int(os.environ.get("PPC_G", "6000")), S=500, C=3, mask=False, seed=20242)
m = P.NBModel(w.counts, w.X, w.exposure, w.K)
rng = np.random.default_rng(0)
n_post = 1000
draws = w.theta_true[None, :] + 0.05 * rng.standard_normal((n_post, w.D))
fit = Fit.from_draws(m, draws)
for exact, n, p in [(True, 0, 0.05), (False, 25000, 4e-4), (True, 0, 0.05)]:
    t = time.perf_counter()
    lo, up, mean, sd = fit.ppc_summary(p, exact=exact, n_draws=n, seed=1)
    dt = time.perf_counter() - t
    nd = (n_post if exact else n) * m.K * m.S
    print(f"exact={exact} n={n_post if exact else n} pairs={m.K*m.S} time={dt:.3f}s draws/s={nd/dt:.3e}  mean upper={up.mean():.1f}")
