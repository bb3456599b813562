"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the contract's keys, the
reference's own binary behind it when oracle/_ref is built, and only rank 0 does the work under a multi-rank launch."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None):
    env = {**os.environ, **(extra_env or {})}
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--reads", "3000", "--cpu-sample", "3000",
                        "-k", "20", "--steps", "2", "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_reference_arm_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "query_reads_per_s" and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2
    # steps / warmup are what the arm really ran (timed passes of the fastest process count / exploratory passes)
    assert d["steps_requested"] == 2 and d["warmup_requested"] == 0 and 1 <= d["steps"] <= 2 and 0 <= d["warmup"] <= 2
    assert d["sample_of"]["reads_per_set"] == 3000 and d["sample_of"]["sample_reads_per_set"] == 3000
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C2") and d["config"]["k"] == 20 and d["config"]["reads_per_set"] == 3000
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert str(cb["cores"]) in cb["tried_reads_per_s"] and max(cb["tried_reads_per_s"].values()) > 0
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    from oracle import oracle
    if oracle.have_ref():
        assert cb["kind"] == "reference"


def test_reference_arm_other_ranks_do_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""
