"""small PPC workload for ncu: exact (1000 draws) and supersampled (50k draws) on 300 genes x 500 samples"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import ppcseq_b200
from ppcseq_b200 import synthetic, Fit
w = synthetic.make("cfg3_60kx500")
Gp = 300
m = ppcseq_b200.NBModel(w.counts[:Gp], w.X, w.exposure, Gp)
lay = m.layout; full = ppcseq_b200.layout(w.G, w.K, w.C)
th = np.zeros(lay.D); th[:3] = w.theta_true[:3]; th[-3:] = w.theta_true[-3:]
for a, b, n in ((lay.o_intercept, full.o_intercept, Gp), (lay.o_sigma_raw, full.o_sigma_raw, Gp), (lay.o_alpha1, full.o_alpha1, Gp), (lay.o_alpha2, full.o_alpha2, Gp)):
    th[a:a + n] = w.theta_true[b:b + n]
draws = th[None, :] + 0.05 * np.random.default_rng(5).standard_normal((1000, lay.D))
fit = Fit.from_draws(m, draws)
fit.ppc_summary(0.05, exact=True, seed=1)
fit.ppc_summary(4e-5, exact=False, n_draws=50000, truncation_compensation=0.7352941, seed=2)
