import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import ppcseq_b200 as P
from ppcseq_b200 import synthetic
w = synthetic.make(sys.argv[1] if len(sys.argv) > 1 else "cfg3_60kx500")
m = P.NBModel(w.counts, w.X, w.exposure, w.K)
if len(w.exclude_pairs) and os.environ.get("NOMASK") != "1": m.set_exclusion(w.exclude_pairs)
th = synthetic.random_thetas(w, 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(3):
    if os.environ.get("NOFLUSH") != "1": flush.zero_()
    torch.cuda.synchronize()
    m.log_prob_grad(th[i % 2])
L = P.lib()
n = 8 * 4096
buf = (ctypes.c_longlong * n)()
L.ppcseq_debug_read.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
print("rc", L.ppcseq_debug_read(buf, n))
a = np.array(buf[:]).reshape(-1, 8)
a = a[a[:, 0] > 0]
t0 = a[:, 0].min()
a = (a - t0) / 1000.0          # us since the first warp started (globaltimer)
names = ["start", "after prologue", "after A", "after B1", "after B2", "after M", "after C", "end"]
print("warps", len(a))
for k, nme in enumerate(names):
    c = a[:, k]
    print(f"{nme:16s} min {c.min():7.2f} p10 {np.percentile(c,10):7.2f} p50 {np.median(c):7.2f} p90 {np.percentile(c,90):7.2f} max {c.max():7.2f}  us")
d = np.diff(a, axis=1)
for k in range(7):
    print(f"  phase {names[k]:>15s} -> {names[k+1]:15s} median {np.median(d[:,k]):6.2f} p90 {np.percentile(d[:,k],90):6.2f} us")
raw = np.array(buf[:]).reshape(-1, 8)
r = (raw[4095, :8] - t0) / 1000.0
print("last warp: group entry %.2f, group summed %.2f | top entry %.2f, hyper ready %.2f, loads done %.2f, sums in smem %.2f, finalize done %.2f us"
      % (r[5], r[6], r[0], r[1], r[4], r[2], r[3]))
