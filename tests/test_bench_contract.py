"""The bench line the driver parses: the committed result of the last GPU run (profiles/) must carry every key of the
contract (task statement, section 4), with the roofline arithmetic consistent, and the reference arm its own keys."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_b200_arm_line():
    d = _load("r2r_bench_cfg3_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["config"]["workload"] == "cfg3_60kx500" and d["dtype"] == "f64" and d["scaling"] == "strong"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "frac_traffic", "fp64_pipe_frac", "binding"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per launch / measured time per launch
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["unit"] == d["unit"]
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_multi_gpu_lines_carry_parity():
    """N > 1 (strong scaling on the fixed BASELINE problem): the line carries the cross-rank parity block."""
    for name, n in (("r2p_bench_cfg3_n2_strong.json", 2), ("r2s_bench_cfg3_n4_strong.json", 4),
                    ("r2s_bench_cfg3_n8_strong.json", 8)):
        d = _load(name)
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["G"] == 60000
        p = d["parity"]
        assert p["ok"] is True and p["ranks_bitwise_identical"] is True and p["all_finite"] is True
        assert p["vs_partial_allreduce_finalize_max_rel"] <= 1e-12
        v = p["vs_single_gpu"]
        assert v["lp_max_rel"] < 1e-12 and v["hyper_grad_max_rel"] < 1e-11 and v["gene_block_gradients_bitwise"] is True
        assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]            # one step = the WHOLE problem
        assert "weak" in d and d["weak"]["value"] > d["value"]


def test_both_arms_print_the_same_config():
    a, b = _load("r2r_bench_cfg3_n1.json"), _load("r2q_bench_cfg3_reference_arm.json")
    assert a["config"] == b["config"]


def test_reference_arm_line():
    d = _load("r2q_bench_cfg3_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "log_prob+grad evals/sec" and d["unit"] == "evals/s"
    assert d["config"]["workload"] == "cfg3_60kx500"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
