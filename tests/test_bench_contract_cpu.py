"""The committed bench line (profiles/r1_bench_tc_n1.json, written by `python bench.py` on a B200) carries every
key of the driver's contract; guards the JSON shape against regressions without needing a GPU."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_tc_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "e2e", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 1344 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]      # samples per step / step time
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["unit"] in ("GB/s", "TFLOP/s") and r["traffic"] is not None
    for name, k in r["kernels"].items():
        assert 0 < k["frac"] < 1, name                      # nothing above its roofline
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    ck = d["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                          "sw_thermal_slowdown"}
