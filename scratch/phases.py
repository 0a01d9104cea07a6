import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
w = synthetic.make(sys.argv[1] if len(sys.argv) > 1 else "cfg2_20kx21")
m = P.NBModel(w.counts, w.X, w.exposure, w.K)
if len(w.exclude_pairs): m.set_exclusion(w.exclude_pairs)
th = synthetic.random_thetas(w, 2)
for i in range(3): m.log_prob_grad(th[i % 2])
L = P.lib()
n = 8 * 2048
buf = (ctypes.c_longlong * n)()
L.ppcseq_debug_read.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
print("rc", L.ppcseq_debug_read(buf, n))
a = np.array(buf[:]).reshape(-1, 8)
a = a[a[:, 0] > 0][:1024]
d = np.diff(a[:, :7], axis=1)
names = ["table","A+sync","B:issue+loads","B:compute","B:flush","B:rest"]
print("CTAs", len(a), "total cycles median", np.median(a[:, 6] - a[:, 0]))
for k, nme in enumerate(names):
    print(f"{nme:12s} median {np.median(d[:, k]):9.0f}  p90 {np.percentile(d[:, k], 90):9.0f}")
print("span first start -> last end (cycles):", a[:, 6].max() - a[:, 0].min())
