"""CPU: the parts of bench.py's contract that need no GPU - the reference arm's JSON line and the clock sampler's
behaviour without nvidia-smi / with rows outside the load window."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    r = json.loads(lines[0])
    assert r["impl"] == "reference" and r["metric"].startswith("Mpix/s encode+decode") and r["unit"] == "Mpix/s"
    assert r["higher_is_better"] is True and r["n_gpus"] == 1 and r["steps"] == 1 and r["value"] > 0
    assert r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] >= 1 and r["cpu_baseline"]["value"] == r["value"]
    assert r["e2e"] == {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in r["config"] and "model" not in r["config"]


def test_clock_sampler_window_and_missing_tool():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]          # never started (ranks > 0, or no tool)
    s = bench.ClockSampler(0)
    s.proc = subprocess.Popen([sys.executable, "-c", "pass"])           # stands in for an nvidia-smi that already exited
    t0 = time.perf_counter()
    s.rows = [(t0 - 1.0, ["300", "1965", "100", "Not Active", "Not Active", "Not Active", "Not Active"]),   # idle, before the load
              (t0 + 0.1, ["1900", "1965", "900", "Not Active", "Not Active", "Not Active", "Active"]),
              (t0 + 0.2, ["1950", "1965", "900", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (t0 + 0.3, ["1965", "1965", "900", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (t0 + 9.0, ["200", "1965", "90", "Active", "Not Active", "Not Active", "Not Active"])]        # after the load
    assert s.count_since(t0) == 4
    c = s.stop(t0, t0 + 1.0)
    assert c["samples"] == 3 and c["sm_mhz"] == 1950.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
