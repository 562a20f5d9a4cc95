"""``bench.py --impl reference`` (the driver's reference arm) on the "tiny" configuration: runs the UNMODIFIED reference's
``gen_image`` on the host cores, prints ONE JSON line with the contract keys, and bounds its sample (calibration step +
time budget) whatever --steps / --warmup are passed. Needs the reference (dev container, or oracle/_ref on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.reference


def run(*extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny", "--height", "256",
                        "--width", "256", *extra], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    line = run("--steps", "2", "--warmup", "1")
    assert "unavailable" not in line, line
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and "gen_image" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["ar_steps_run"] == {"warmup": 1, "timed": 2}


def test_reference_sample_is_bounded(monkeypatch):
    """a slow host (calibration step >> budget) shrinks the sample to 1 warm-up + 1 timed step and says so"""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import ref_runner as rr
    calls = []
    real = rr.run_bounded

    def slow(pipe, info, *, n_ar, **kw):
        calls.append(n_ar)
        out = real(pipe, info, n_ar=n_ar, **kw)
        if len(calls) == 1:
            out["ar_s"] = [1000.0]          # pretend the cold calibration step took 1000 s
        return out

    monkeypatch.setattr(rr, "run_bounded", slow)
    r = bench.reference_sample("tiny", "cpu", 3, 5, 3, 3.0, 256, 1, threads=2, with_decode=False, budget_s=150.0)
    assert calls == [1, 2] and (r["n_warm"], r["n_timed"]) == (1, 1) and "bounded" in r["sample"]
    calls.clear()
    monkeypatch.setattr(rr, "run_bounded", lambda pipe, info, *, n_ar, **kw: (calls.append(n_ar), real(pipe, info, n_ar=n_ar, **kw))[1])
    r = bench.reference_sample("tiny", "cpu", 3, 5, 3, 3.0, 256, 1, threads=2, with_decode=False, budget_s=150.0)
    assert calls == [1, 8] and (r["n_warm"], r["n_timed"]) == (3, 5) and "bounded" not in r["sample"]
