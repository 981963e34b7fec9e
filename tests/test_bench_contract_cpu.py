"""CPU-only guard of the bench.py output contract: the most recent committed bench lines of both arms
(profiles/r02_bench*.json, produced on a B200) must carry every key the driver and the judge read."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, f"no committed bench line matches {pattern}"
    with open(files[-1]) as f:
        return json.load(f), files[-1]


def test_own_arm_line_has_the_contract_keys():
    line, path = _latest("r02_bench.json")
    missing = (BASE_KEYS | {"gpu_launches", "clocks", "roofline", "ops", "scenes", "stability"}) - set(line)
    assert not missing, f"{path}: missing {sorted(missing)}"
    assert line["metric"].startswith("scans/sec") and line["unit"] == "scans/s" and line["higher_is_better"] is True
    assert line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["steps"] >= 1 and line["warmup"] >= 3
    assert abs(line["value"] - line["n_gpus"] * 1e3 / line["ms_per_step"]) <= 1e-6 * line["value"]
    e2e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e2e)
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] != line["value"]
    assert line["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert roof["bound"] in ("hbm", "tensor") and roof["unit"] in ("GB/s", "TFLOP/s")
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) <= 1e-9
    cpu = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cpu) and cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1
    # BASELINE configs[2..4] ride in the same line: every scene and every operator entry with a recomputable fraction
    assert {s["scene"] for s in line["scenes"]} == {"kitti", "scannet"} and all("error" not in s and s["fwd_bwd_ms"] > 0 for s in line["scenes"])
    ops = line["ops"]
    assert ops["tf32_peak_TFLOPs_measured_here"] > 100 and len(ops["entries"]) >= 12
    for e in ops["entries"]:
        if "GBps" in e:
            assert abs(e["hbm_frac"] - e["GBps"] / ops["hbm_peak_GBps"]) <= 1e-6
        if "TFLOPs" in e:
            assert abs(e["tf32_frac"] - e["TFLOPs"] / ops["tf32_peak_TFLOPs_measured_here"]) <= 1e-6


def test_reference_arm_line_has_the_contract_keys():
    line, path = _latest("r02_bench_reference.json")
    missing = (BASE_KEYS | {"impl"}) - set(line)
    assert not missing, f"{path}: missing {sorted(missing)}"
    assert line["impl"] == "reference" and line["unit"] == "scans/s"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]
    assert line["cpu_baseline"]["kind"] == "reference"
    assert "reference Python modules unmodified" in line["config"]["arm"] and len(line["scenes"]) == 2


def test_multi_gpu_lines_scale_and_stay_in_sync():
    one, _ = _latest("r02_bench.json")
    for pattern, n in (("r02_bench_2gpu.json", 2), ("r02_bench_8gpu.json", 8)):
        line, path = _latest(pattern)
        assert line["n_gpus"] == n and line["scaling"] == "weak"
        assert line["config"]["replicas_bit_identical_after_run"] is True, path
        assert line["value"] > 0.85 * n * one["value"], f"{path}: {line['value']:.0f} scans/s on {n} GPUs vs {one['value']:.0f} on one"
