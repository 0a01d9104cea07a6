import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_inference_gpu import _tidy
from ppcseq_b200.api import identify_outliers
z, df = _tidy("bundled_readme515.npz")
for seed in (21, 22, 23, 24, 25):
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=5, seed=seed)
    failed, dele = res["ppc_samples_failed"].to_numpy(), res["tot_deleterious_outliers"].to_numpy()
    bad = [(g, int(f), int(d), int(e0), int(e1)) for g, f, d, (e0, e1) in zip(z["expected_genes"], failed, dele, z["expected"]) if f != e0 or d != e1]
    print("seed", seed, "mismatches (gene, failed, dele, exp_failed, exp_dele):", bad)
