import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ppcseq_b200 import NBModel, inference
from oracle import c_oracle, model_np
from tests.helpers import grad_err, rel
g = np.load("tests/golden/nuts_golden.npz")
m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
m.set_design_path(2)
fit = inference.sample_nuts(m, chains=4, iter=650, warmup=150, seed=11)
dr = fit.draws(0, m.D)
d = model_np.ModelData(g["counts"], g["X"], g["exposure"], int(g["K"]))
worst = (0, 0)
for mode in (2, 3):
    m.set_design_path(mode)
    lps, grs = m.log_prob_grad(dr[::4])
    el, eg = [], []
    for i, th in enumerate(dr[::4]):
        lr, gr_ = c_oracle.log_prob_grad(d, th)
        el.append(rel(lps[i], lr)); eg.append(grad_err(grs[i], gr_))
    el, eg = np.array(el), np.array(eg)
    print("mode", mode, "max rel lp err", el.max(), "max grad err", eg.max(), "at", eg.argmax())
    if mode == 3:
        i = eg.argmax(); th = dr[::4][i]
        lr, gr_ = c_oracle.log_prob_grad(d, th)
        sc = np.maximum(np.abs(gr_), 1e-3 * np.abs(gr_).max())
        k = np.argmax(np.abs(grs[i] - gr_) / sc)
        lay = m.layout
        print("worst comp", k, grs[i][k], gr_[k], "sigma_raw range", th[lay.o_sigma_raw:lay.o_sigma_raw+m.G].min(), th[lay.o_sigma_raw:lay.o_sigma_raw+m.G].max())
