import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
w = synthetic.make("cfg3_60kx500")
m = P.NBModel(w.counts, w.X, w.exposure, w.K)
m.set_exclusion(w.exclude_pairs)
ths = np.vstack([w.theta_true, synthetic.random_thetas(w, 2, seed=9)])
lay = m.layout
for i in range(3):
    m.set_design_path(2); lp2, g2 = m.log_prob_grad(ths[i])
    m.set_design_path(3); lp3, g3 = m.log_prob_grad(ths[i])
    d = np.abs(g3 - g2) / np.maximum(np.abs(g2), 1e-3 * np.abs(g2).max())
    bad = np.nonzero(~(d < 1e-9))[0]
    print(i, "lp2", lp2, "lp3", lp3, "n_bad", len(bad), bad[:12])
    sig = bad[(bad >= lay.o_sigma_raw) & (bad < lay.o_sigma_raw + w.G)] - lay.o_sigma_raw
    if len(sig):
        print("  bad genes (sigma_raw idx)", sig[:20], " mod32", (sig % 32)[:20], "supertile", (sig // 32)[:20])
        gg = sig[0]
        print("  gene", gg, "counts min/max", w.counts[gg].min(), w.counts[gg].max(), "theta sr", ths[i][lay.o_sigma_raw + gg],
              "g2", g2[lay.o_sigma_raw + gg], "g3", g3[lay.o_sigma_raw + gg])
