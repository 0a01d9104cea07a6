import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
def check(w, tag):
    m = P.NBModel(w.counts, w.X, w.exposure, w.K)
    if len(w.exclude_pairs): m.set_exclusion(w.exclude_pairs)
    ths = synthetic.random_thetas(w, 4, seed=9)
    m.set_design_path(3)
    one = [m.log_prob_grad(ths[i]) for i in range(4)]
    for B in (2, 3, 4):
        lp, g = m.log_prob_grad(ths[:B])
        for b in range(B):
            nbad = int((g[b] != one[b][1]).sum())
            bad = np.nonzero(g[b] != one[b][1])[0]
            print(tag, "B", B, "b", b, "lp equal", lp[b] == one[b][0], "grad mismatches", nbad, bad[:6], flush=True)
    m.set_design_path(2)
    lp2, g2 = m.log_prob_grad(ths[0])
    print(tag, "path2 vs path3 (B=1)", lp2, one[0][0])
check(synthetic.make("cfg3_60kx500"), "cfg3")
check(synthetic.make(G=130000, S=21, C=2, mask=False, seed=3), "130kx21")
