"""The JSON line bench.py prints is a contract with the driver: the recorded lines of this round (profiles/r01_bench_line_*.json,
copied from real B200 runs) must carry every key the contract names, with consistent values. CPU only."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_line_*.json")))


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_recorded_bench_line_has_the_contract_keys(path):
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "infer_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and d["warmup"] >= 3
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.001
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1 and r["traffic"] is None or r["traffic"] > 0
    # whole-job value = frames of all ranks / time
    b = d["config"]["batch_per_gpu"]
    assert abs(d["value"] - d["n_gpus"] * b / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    if "cpu_baseline" in d:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == d["unit"] and cb["sample"]
    if "train" in d:
        t = d["train"]
        assert t["metric"] == "train_samples_per_sec" and t["scaling"] == "strong" and t["gpu_launches"] > 0
        assert abs(t["value"] - t["global_batch"] / (t["ms_per_step"] / 1e3)) < 1e-6 * t["value"]


def test_a_single_gpu_line_with_cpu_baseline_is_recorded():
    assert any("cpu_baseline" in json.load(open(p)) and json.load(open(p))["n_gpus"] == 1 for p in LINES)


R02 = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_line_*.json")))


@pytest.mark.parametrize("path", R02, ids=[os.path.basename(p) for p in R02])
def test_recorded_round2_bench_line_has_the_contract_keys(path):
    """Round 2: the training step (BASELINE configs[2]) is the primary metric, strong scaling over the ranks."""
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_hbm"):
        assert k in d, k
    assert d["metric"] == "train_samples_per_sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and d["warmup"] >= 3
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
    assert d["gpu_launches"] > 0 and d["config"]["launch_mode"] == "cuda_graph"
    c = d["clocks"]
    assert c["sm_mhz"] > 0 and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    for r, bound, unit in ((d["roofline"], "tensor", "TFLOP/s"), (d["roofline_hbm"], "hbm", "GB/s")):
        assert r["bound"] == bound and r["unit"] == unit and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
        assert r["traffic"] is None or r["traffic"] > 0
    if os.path.basename(path) == "r02_bench_line_n1.json":
        assert "cpu_baseline" in d and "infer" in d      # the driver-shaped single-GPU line carries both
    if "cpu_baseline" in d:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["unit"] == d["unit"] and cb["sample"]
    if "infer" in d:
        assert d["infer"]["metric"] == "infer_frames_per_sec" and d["infer"]["e2e"]["d2h_bytes_per_step"] > 0
