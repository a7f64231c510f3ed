"""The bench.py JSON contract (what the driver parses), checked on the CPU for the reference arm and
on the GPU for the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, p.stdout          # exactly ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "2", "--warmup", "1"], 300)
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["metric"] == "NV12->I420 frames/s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "nv12_to_i420_1080p_x300_pitch2048" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["cpu_model"] and "-O2" in cb["compiler_flags"]                # CPU model and compiler flags beside the CPU figure
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 100


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_product_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-extras"], 600)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["gpu_launches"] == 3                                   # one launch per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_launch"] == 300 * 6220800       # 3*w*h per frame
    assert r["achieved"] > 0.7 * r["peak"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 300 * 1920 * 1620 and e["d2h_bytes_per_step"] == 300 * 3110400
    assert 0 < e["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_max_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert d["verified_bit_exact"] is True


def test_both_arms_carry_the_same_config_and_traffic_is_tied_to_the_sources(tmp_path, monkeypatch):
    """`config` comes from one function for both arms (the driver compares the two objects), and roofline.traffic is
    reported only while the recorded ncu capture belongs to the kernel sources in the tree."""
    sys.path.insert(0, ROOT)
    import bench
    a = bench.run_config("nv12_to_i420_1080p_x300_pitch2048", 4)
    assert a == bench.run_config("nv12_to_i420_1080p_x300_pitch2048", 4)
    assert a["workload"] == "nv12_to_i420_1080p_x300_pitch2048" and a["frames_per_step_per_gpu"] == 300 and "L2" in a["l2"]
    assert "4 GPU" in a["parallelism"] and "model" not in a
    h = bench.kernel_sources_sha256()
    assert len(h) == 64
    t, prov = bench.recorded_traffic("nv12_to_i420_1080p_x300_pitch2048")
    assert prov["status"] in ("current", "stale: kernel sources changed since the capture", "no ncu capture recorded for this workload")
    assert (t is not None) == (prov["status"] == "current")
    # a capture taken from other sources is never reported
    fake = {"nv12_to_i420_1080p_x300_pitch2048": {"traffic": 123, "kernel": "k", "capture": "c", "git": "g", "when": "w", "sources_sha256": "0" * 64}}
    (tmp_path / "profiles").mkdir()
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps(fake))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_sources_sha256", lambda: h)
    t, prov = bench.recorded_traffic("nv12_to_i420_1080p_x300_pitch2048")
    assert t is None and prov["status"].startswith("stale")
    fake["nv12_to_i420_1080p_x300_pitch2048"]["sources_sha256"] = h
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps(fake))
    assert bench.recorded_traffic("nv12_to_i420_1080p_x300_pitch2048")[0] == 123


def test_ratios_against_the_cpu_baseline_of_the_same_run():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.ratios_vs_cpu({"value": 20000.0, "cores": 16, "kind": "reference"}, 1.1e6, 15500.0, {"value": 17400.0})
    assert abs(r["e2e"] - 0.775) < 1e-9 and abs(r["e2e_device_resident_input"] - 0.87) < 1e-9 and abs(r["device_only"] - 55.0) < 1e-9
    assert "16 threads" in r["basis"]
    assert bench.ratios_vs_cpu({"value": 20000.0}, 1.0, 1.0, None)["e2e_device_resident_input"] is None
    assert bench.ratios_vs_cpu({"value": 0}, 1.0, 1.0, None) is None and bench.ratios_vs_cpu({}, 1.0, 1.0, None) is None
