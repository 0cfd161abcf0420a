"""The committed bench lines (profiles/r2_bench_*.json, written by `python bench.py` on B200 boxes) carry every key of
the driver's contract; guards the JSON shape against regressions without needing a GPU."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_committed_bench_line_has_the_contract_keys():
    d = _load("r2_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "e2e", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "strong" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    c = d["config"]
    assert c["subjects_total"] == 42 and c["subjects_per_gpu"] == [42] and c["batch_sizes"] == [32] * 8 + [24]
    # value = training samples of the timed steps / their time
    assert abs(d["value"] - c["train_samples_in_timed_region"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["unit"] in ("GB/s", "TFLOP/s") and r["traffic"] is not None
    for name, k in r["kernels"].items():
        assert 0 < k["frac"] < 1, name                      # nothing above its roofline
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    ck = d["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                          "sw_thermal_slowdown"}
    p = d["preprocess"]
    assert p["roofline"]["bound"] == "hbm" and p["roofline"]["traffic"] and 0 < p["roofline"]["frac"] < 1


def test_committed_8gpu_line_is_the_42_subject_workload():
    d = _load("r2_bench_n8.json")
    assert d["n_gpus"] == 8 and d["scaling"] == "strong"
    assert d["config"]["subjects_total"] == 42 and d["config"]["subjects_per_gpu"] == [6, 6, 5, 5, 5, 5, 5, 5]
    assert d["dp_parity"]["ok"] is True
    one = _load("r2_bench_n1.json")
    assert 4.5 < d["value"] / one["value"] <= 7.0            # strong scaling of 42 subjects: ceiling 42 / 6 = 7.0


def test_committed_scaling_lines_are_monotonic():
    vals = [_load(f"r2_bench_n{n}.json") for n in (1, 2, 4, 8)]
    assert [v["n_gpus"] for v in vals] == [1, 2, 4, 8]
    assert all(v["config"]["subjects_total"] == 42 and sum(v["config"]["subjects_per_gpu"]) == 42 for v in vals)
    assert all(b["value"] > 1.5 * a["value"] for a, b in zip(vals, vals[1:]))      # every doubling pays at least 1.5x
    assert all(v["dp_parity"]["ok"] for v in vals[1:])


def test_reference_arm_line():
    d = _load("r2_bench_ref.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_cpu_arm_survives_a_multi_gpu_box(monkeypatch):
    """The stock Trainer_uni wraps its model in nn.DataParallel when it sees more than one GPU (EEGNet_tor.py:85-87); the CPU arm
    of bench.py (cpu_baseline at N=1, --impl reference) must not, or an N=1 run on an 8-GPU box dies in the baseline leg."""
    import importlib.util
    import sys

    import pytest
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not present (oracle/_ref or /root/reference)")
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 8)
    monkeypatch.setattr(bench, "_host_threads", lambda: 4)          # one thread-count candidate: keeps the test short
    r = bench.cpu_train_reference(1, 1)
    assert r["kind"] == "reference" and r["value"] > 0
    assert torch.cuda.device_count() == 8                            # the patch inside was undone
