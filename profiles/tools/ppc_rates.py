import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
import ppcseq_b200
from ppcseq_b200 import synthetic, Fit
import bench
w = synthetic.make("cfg3_60kx500")
r = bench.ppc_bench(w, 0)
print("exact", r["value"]/1e9, "G draws/s")
# supersampled: 250k draws per pair from 1000 posterior draws, 400 genes
Gp = 400
m = ppcseq_b200.NBModel(w.counts[:Gp], w.X, w.exposure, Gp)
lay = m.layout; full = ppcseq_b200.layout(w.G, w.K, w.C)
th = np.zeros(lay.D); th[:3] = w.theta_true[:3]; th[-3:] = w.theta_true[-3:]
th[lay.o_intercept:lay.o_intercept+Gp] = w.theta_true[full.o_intercept:full.o_intercept+Gp]
th[lay.o_sigma_raw:lay.o_sigma_raw+Gp] = w.theta_true[full.o_sigma_raw:full.o_sigma_raw+Gp]
th[lay.o_alpha1:lay.o_alpha1+Gp] = w.theta_true[full.o_alpha1:full.o_alpha1+Gp]
th[lay.o_alpha2:lay.o_alpha2+Gp] = w.theta_true[full.o_alpha2:full.o_alpha2+Gp]
draws = th[None,:] + 0.05*np.random.default_rng(5).standard_normal((1000, lay.D))
fit = Fit.from_draws(m, draws)
for nd in (25000, 250000):
    fit.ppc_summary(4e-5, exact=False, n_draws=2000, truncation_compensation=0.7352941, seed=1)
    t0=time.perf_counter(); fit.ppc_summary(4e-5, exact=False, n_draws=nd, truncation_compensation=0.7352941, seed=2); dt=time.perf_counter()-t0
    print("supersampled n_draws", nd, nd*Gp*w.S/dt/1e9, "G draws/s")
