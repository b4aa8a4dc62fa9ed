"""CPU: the bench.py contract pieces that run without a GPU - the reference arm prints exactly one JSON line with the keys the
driver reads (`impl: reference`, metric / unit / value, `cpu_baseline`, zero-copy `e2e`), under torchrun ranks > 0 stay silent,
and the argument surface of the train mode exists (`--mode train --train-gemm tf32|fp32`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "dinounet_s", "--size", "128",
                        "--batch", "1", "--steps", "1", "--warmup", "0", *extra], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_with_the_driver_keys():
    lines = _run([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patches/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("2D patches/sec") and d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_non_zero_ranks():
    assert _run([], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


def test_train_mode_arguments_exist():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "--train-gemm" in r.stdout and "--mode" in r.stdout and "--pdl" in r.stdout
