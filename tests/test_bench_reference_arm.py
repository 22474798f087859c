"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one tiny step through the subprocess the driver would
launch, checked against the JSON contract; and under a torchrun-style environment only rank 0 runs and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "1", "--size", "64", "--steps", "1", "--warmup", "0", "--ref-seconds", "5"]


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run(CMD, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref")):        # staged by build(): then it is the reference's own modules that ran
        assert cb["kind"] == "reference"


def test_reference_arm_other_ranks_do_nothing():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
