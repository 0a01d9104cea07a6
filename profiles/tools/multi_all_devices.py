"""One process, EVERY GPU of the box: the checks of tests/test_multi_gpu.py (lp / gradients against the oracle and the
unsharded handle, PPC and flags bitwise, fit queries) with devices = [0 .. n-1], then NUTS, ADVI and
identify_outliers() on that handle.  Run on the GPU box: python profiles/tools/multi_all_devices.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

n = torch.cuda.device_count()
devs = list(range(n))
from tests import test_multi_gpu as T
t0 = time.perf_counter()
T._check_against_single(devs)
print("ok parity on", devs, round(time.perf_counter() - t0, 1), "s", flush=True)

from ppcseq_b200 import NBModel, inference
g = np.load(os.path.join(T.GOLD, "nuts_golden.npz"))
m = NBModel(g["counts"], g["X"], g["exposure"], int(g["K"]), devices=devs)
t0 = time.perf_counter()
fit = inference.sample_nuts(m, chains=4, iter=150 + 600, warmup=150, seed=21)
dr = fit.draws(0, m.D)
z = np.abs(dr.mean(axis=0) - g["mean"]) / g["sd"]
assert np.delete(z, len(z) - 1).max() < 0.35 and np.percentile(z, 90) < 0.2, float(z.max())
print("ok nuts", round(time.perf_counter() - t0, 1), "s; max z", float(np.delete(z, len(z) - 1).max()), flush=True)
vb = inference.advi(m, output_samples=1000, iter=20000, tol_rel_obj=0.005, seed=4)
lay = m.layout
dv = vb.draws(lay.o_intercept, m.G)
ref_m, ref_s = g["mean"][lay.o_intercept:lay.o_intercept + m.G], g["sd"][lay.o_intercept:lay.o_intercept + m.G]
assert (np.abs(dv.mean(axis=0) - ref_m) / ref_s).max() < 1.0
print("ok advi", flush=True)
m.close()

from ppcseq_b200.api import identify_outliers
from tests.test_inference_gpu import _tidy
z, df = _tidy("bundled_test53.npz")
for vbi in (True, False):
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value", significance="PValue",
                            do_check="is_significant", percent_false_positive_genes=1, approximate_posterior_inference=vbi,
                            how_many_negative_controls=50, cores=4, seed=7, devices=devs)
    assert list(res["symbol"]) == ["SLC16A12", "CYP1A1", "ART3"]
    assert list(res["tot_deleterious_outliers"].astype(int)) == [0, 1, 0], list(res["tot_deleterious_outliers"])
    print("ok identify_outliers vb =", vbi, flush=True)
print("all devices pass done:", n, "GPUs")
