import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
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
    print("ok", G, S, C, lp)
