"""CPU-only: the reference arm of bench.py (`--impl reference`: the reference's numba loops staged under oracle/_ref when
present, else the oracle port, on the host cores) prints one JSON line with the contract's keys; under torchrun
semantics only rank 0 works."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(env_extra=None, args=("--steps", "1", "--warmup", "0", "--cpu-sample-times", "2")):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


@pytest.mark.parametrize("engine", ["port", "auto"])
def test_reference_arm_line(engine):
    lines = _run(args=("--steps", "1", "--warmup", "0", "--cpu-sample-times", "2", "--cpu-engine", engine))
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "visibilities gridded/sec" and d["unit"] == "vis/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    staged = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "ngcasa", "imaging", "_imaging_utils", "_standard_grid.py")) \
        or os.path.isdir("/root/reference")
    want = "port" if engine == "port" or not staged else "reference"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "vis/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_arm_sizes_its_sample_from_calibration_steps():
    """Without --cpu-sample-times the sample is sized so that steps + warmup fit the time budget, capped at the whole
    workload (a small one here: every integration is taken) and the line says what it took."""
    lines = _run(args=("--steps", "2", "--warmup", "1", "--n-time", "48", "--n-chan", "8", "--n-uv", "512"))
    d = json.loads(lines[0])
    assert d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["sample"].startswith("48 of 48 integrations") and "calibration steps of 8 and 40" in d["cpu_baseline"]["sample"]
