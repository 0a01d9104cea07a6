#!/usr/bin/env python
"""identify_outliers() wall-clock at BASELINE scale, both passes, split into prep / upload / pass 1 / pass 2 / result.

  python profiles/tools/e2e_identify_outliers.py cfg3_60kx500 [--inference vb|nuts] [--devices 0,1] [--pfp 1]
        [--just-discovery] [--format failing]

The tidy input table (one row per gene x sample) is built from the synthetic workload; transcripts and samples are
integer ids (hashing 3e7..3e8 Python strings is a property of the table, not of the path).  Every gene is checked
(K = G, BASELINE "all genes checked").  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--inference", default="vb", choices=["vb", "nuts"])
    ap.add_argument("--devices", default="0")
    ap.add_argument("--pfp", type=float, default=1.0)
    ap.add_argument("--just-discovery", action="store_true")
    ap.add_argument("--format", default="failing")
    ap.add_argument("--genes", type=int, default=0, help="use only the first N genes of the workload")
    ap.add_argument("--cores", type=int, default=4)
    ap.add_argument("--p2", type=float, default=0.0, help="adj_prob_theshold_2 override (e.g. 1.25e-3 => 8,000 draws)")
    ap.add_argument("--exact-analysis", action="store_true", help="approximate_posterior_analysis = FALSE in pass 2")
    a = ap.parse_args()
    import pandas as pd

    from ppcseq_b200 import synthetic
    from ppcseq_b200.api import identify_outliers
    t0 = time.perf_counter()
    w = synthetic.make(a.workload)
    G = a.genes or w.G
    S = w.S
    cols = {"symbol": np.repeat(np.arange(G, dtype=np.int64), S), "sample": np.tile(np.arange(S, dtype=np.int64), G),
            "value": w.counts[:G].reshape(-1), "PValue": np.repeat(np.linspace(1e-9, 1e-3, G), S),
            "do_check": np.ones(G * S, bool)}
    formula = "~ Label"
    cols["Label"] = np.tile(np.where(w.X[:, 1] > 0, "B", "A"), G)
    if w.C >= 3:
        cols["batch"] = np.tile(np.where(w.X[:, 2] > 0, "y", "x"), G)
        formula = "~ Label + batch"
    df = pd.DataFrame(cols)
    t_table = time.perf_counter() - t0
    devices = [int(d) for d in a.devices.split(",")]
    tm = {}
    t0 = time.perf_counter()
    res = identify_outliers(df, formula, sample="sample", transcript="symbol", abundance="value", significance="PValue",
                            do_check="do_check", percent_false_positive_genes=a.pfp, how_many_negative_controls=0,
                            approximate_posterior_inference=(a.inference == "vb"), cores=a.cores, seed=11,
                            devices=devices if len(devices) > 1 else None, device=devices[0], return_format=a.format,
                            just_discovery=a.just_discovery, timings=tm,
                            adj_prob_theshold_2=a.p2 if a.p2 > 0 else None,
                            approximate_posterior_analysis=False if a.exact_analysis else True)
    wall = time.perf_counter() - t0
    tot = res.attrs.get("gene_totals")
    i1, i2 = tm.pop("pass1_info"), tm.pop("pass2_info")
    out = {"workload": a.workload, "G": G, "S": S, "C": w.C, "inference": a.inference, "devices": devices, "pfp": a.pfp, "p2": a.p2,
           "exact_analysis": a.exact_analysis,
           "wall_s": wall, "split_s": tm, "make_table_s": t_table, "rows": int(G) * int(S),
           "pass1": {"evals": i1[1], "sampler_s": i1[2], "detail": [float(x) for x in i1[3:9]]},
           "pass2": {"evals": i2[1], "sampler_s": i2[2], "detail": [float(x) for x in i2[3:9]]},
           "result_rows": int(len(res)), "genes_with_failed_samples": None if tot is None else int((tot["ppc_samples_failed"] > 0).sum()),
           "pass2_ppc_draws": float(res.attrs["total_draws"]), "just_discovery": a.just_discovery}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
