"""bench.py on the CPU: the reference arm prints ONE JSON line with the contract's keys, and its `config` object is the
one the GPU arm uses (the driver compares them).  The GPU arm itself needs a B200 (driver-run)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config("c1")
    assert bench.bench_config("c3")["workload"] == "c3" and "l2_policy" in bench.bench_config("c3")
    # the default workload is the largest single-GPU configuration (BASELINE.json configs[2])
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--workload", default="c3"' in src
