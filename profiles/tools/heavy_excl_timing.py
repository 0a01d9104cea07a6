"""log_prob+grad time with a heavy pass-2 exclusion list (config 5 shape, 3.5 % of the points excluded = 175 per gene)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import ppcseq_b200
from ppcseq_b200 import synthetic
from ppcseq_b200._lib import check
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_60kx5000"
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.035
w = synthetic.make(name)
L = ppcseq_b200.lib()
m = ppcseq_b200.NBModel(w.counts, w.X, w.exposure, w.K)
rng = np.random.default_rng(3)
dev = torch.device("cuda", 0)
th = torch.from_numpy(np.ascontiguousarray(synthetic.random_thetas(w, 2))).to(dev)
lp = torch.zeros(1, dtype=torch.float64, device=dev); gr = torch.zeros(m.D, dtype=torch.float64, device=dev)
def t():
    ms = (ctypes.c_float * 20)()
    check(L.ppcseq_time_log_prob_grad_device(m.handle, 1, th[0].data_ptr(), 1, 1, lp.data_ptr(), gr.data_ptr(), None, 20, 1, ms))
    return float(np.median(np.array(ms[5:])) * 1e3)
print(name, "no exclusions", round(t(), 1), "us")
n = int(frac * w.G * w.S)
flat = rng.choice(w.G * w.S, n, replace=False)
pairs = np.stack([flat // w.S, flat % w.S], 1).astype(np.int32)
for sub, lab in ((pairs[: w.G * 12], "12 per gene (list)"), (pairs, f"{n // w.G} per gene")):
    m.set_exclusion(sub)
    print(name, lab, round(t(), 1), "us")
