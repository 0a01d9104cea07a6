import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_inference_gpu import _tidy
from ppcseq_b200.api import identify_outliers
z, df = _tidy("bundled_readme515.npz")
cnt = collections.Counter()
N = 40
for seed in range(100, 100 + N):
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=5, seed=seed)
    failed = res["ppc_samples_failed"].to_numpy()
    for g, f, (e0, e1) in zip(z["expected_genes"], failed, z["expected"]):
        if f != e0: cnt[str(g)] += 1
print("mismatch frequency over", N, "seeds:", dict(cnt))
