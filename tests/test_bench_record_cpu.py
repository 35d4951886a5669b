"""The committed bench record (profiles/r02_bench_default.json: the one JSON line `python bench.py`
printed on a B200) satisfies the driver's contract and is arithmetically self-consistent."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        lines = [l for l in f.read().splitlines() if l.strip()]
    assert len(lines) == 1, "one JSON line"
    return json.loads(lines[0])


def test_default_record_contract():
    d = _line("r02_bench_default.json")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
                "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["unit"] == base.get("unit", d["unit"])
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == d["steps"]                                # one fused launch per step
    # value = poses of all steps / device time
    poses = d["run"]["batch_per_gpu_per_step"] * d["n_gpus"]
    assert poses == 1 << 26                                              # BASELINE configs[1] chunk
    assert abs(d["value"] - poses / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.85 < r["frac"] < 1.0
    assert r["traffic"] is not None and abs(r["traffic"] / (r["algorithmic_bytes_per_unit"] * poses) - 1) < 0.05
    assert r["ms_per_launch"] <= d["ms_per_step"] * 1.001                 # the kernel fits in the step
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert 0.85 < e["link"]["frac"] <= 1.1                                # at the link roofline
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    clk = d["clocks"]
    assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the other BASELINE configs ride in the same line
    for name in ("ik", "traj", "calib"):
        w = d["workloads"][name]
        assert w["ms_per_step"] > 0 and "roofline" in w and "cpu_baseline" in w
    assert d["workloads"]["calib"]["roofline"]["bound"] == "fp64"
    assert d["workloads"]["calib"]["ms_per_step"] < 100.0                 # north_star: < 100 ms per 1e8


def test_reference_arm_record_contract():
    d = _line("r02_bench_reference_arm.json")
    mine = _line("r02_bench_default.json")
    assert d["impl"] == "reference" and d["metric"] == mine["metric"] and d["unit"] == mine["unit"]
    assert d["config"] == mine["config"]                                  # same workload description
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"]
    ref = d["cpu_baseline_reference"]                                     # the UNMODIFIED reference beside it
    assert ref["kind"] == "reference" and 10 < ref["value"] < 1e4
