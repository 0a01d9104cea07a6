import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
w = synthetic.make("cfg3_60kx500")
m = P.NBModel(w.counts, w.X, w.exposure, w.K)
m.set_exclusion(w.exclude_pairs)
lay = m.layout
B = 6
ths = synthetic.random_thetas(w, B, seed=9)
ths[0] = w.theta_true
m.set_design_path(3)
one = [m.log_prob_grad(ths[i]) for i in range(B)]
tot = 0
for it in range(30):
    lp, g = m.log_prob_grad(ths)
    for b in range(B):
        bad = np.nonzero(g[b] != one[b][1])[0]
        if len(bad):
            tot += 1
            parts = {"icpt": ((bad >= lay.o_intercept) & (bad < lay.o_intercept + w.G)).sum(),
                     "a1": ((bad >= lay.o_alpha1) & (bad < lay.o_alpha1 + w.K)).sum(),
                     "sig": ((bad >= lay.o_sigma_raw) & (bad < lay.o_sigma_raw + w.G)).sum()}
            sig = bad[(bad >= lay.o_sigma_raw) & (bad < lay.o_sigma_raw + w.G)] - lay.o_sigma_raw
            gg = sig[0] if len(sig) else -1
            print("it", it, "b", b, "nbad", len(bad), parts, "supertiles", sorted(set((sig // 32).tolist()))[:8],
                  "d_sigma", g[b][lay.o_sigma_raw + gg], "vs", one[b][1][lay.o_sigma_raw + gg], "lp diff", lp[b] - one[b][0], flush=True)
print("total bad (b, iteration) pairs", tot, "of", 30 * B)
