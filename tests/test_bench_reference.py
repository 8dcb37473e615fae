"""The reference arm of bench.py (`--impl reference`) runs on the host alone: the CPU oracle re-host of
apps/run_euroc_single_thread.cpp, one process per core.  Checks the JSON contract of its line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, MSKF_BENCH_NO_CLOCKS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "frames/s"
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
