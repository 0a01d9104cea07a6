import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ppcseq_b200 import NBModel, inference
g = np.load("tests/golden/nuts_golden.npz")
for mode in (2, 3):
    for seed in (11, 12, 13):
        m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]))
        m.set_design_path(mode)
        fit = inference.sample_nuts(m, chains=4, iter=1150, warmup=150, seed=seed)
        dr = fit.draws(0, m.D)
        z = np.abs(dr.mean(0) - g["mean"]) / g["sd"]
        top = np.argsort(z)[-3:][::-1]
        print("mode", mode, "seed", seed, "zmax", z.max().round(3), "args", top, z[top].round(3), "p90", np.percentile(z, 90).round(3), "info", fit.info(8)[[3, 5, 6, 7]].round(3))
print("oracle chain means spread (z):", (np.abs(g["chain_means"] - g["mean"]) / g["sd"]).max(axis=0)[[0,1,2,63,64,65]].round(3))
