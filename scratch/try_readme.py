import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_inference_gpu import _tidy
from ppcseq_b200.api import identify_outliers
z, df = _tidy("bundled_readme515.npz")
for seed, vb in [(21, True), (22, True), (23, True), (24, True), (25, False)]:
    res = identify_outliers(df, "~ Label", sample="sample", transcript="symbol", abundance="value",
                            significance="PValue", do_check="is_significant", percent_false_positive_genes=5, seed=seed,
                            approximate_posterior_inference=vb, cores=4)
    print(seed, vb, res["ppc_samples_failed"].to_numpy(), res["tot_deleterious_outliers"].to_numpy(), res.attrs["fit 2 info"][:5])
    sw = res["sample_wise_data"].iloc[14]
    bad = sw[~sw["posterior_predictive_check_succeded"]]
    print(bad[["value", ".lower", ".upper", "deleterious_outliers", "Label", "slope_after_outlier_filtering"]])
    print(sw[["value", ".lower", ".upper"]].T.to_string())
