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


def test_roofline_traffic_file_is_what_the_committed_capture_says(tmp_path):
    """bench.py reads roofline.traffic from profiles/ncu_traffic.json: that file must be exactly what
    scripts/ncu_traffic.py derives from the committed ncu --set full summary, and carry the search kernel of the
    bench workload (bench.ncu_traffic returns it for C2 and nothing for another workload)."""
    dst = tmp_path / "t.json"
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_traffic.py"), str(ROOT / "profiles" / "r02_ncu_full_summary.csv"),
                        str(dst)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    want = json.loads(dst.read_text())
    got = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
    assert got["kernels"] == want["kernels"] and got["workload"] == want["workload"]
    sys.path.insert(0, str(ROOT))
    import bench
    c2 = {"reads_per_set": 10_000_000, "read_len": 100, "k": 33, "t": 2}
    s = got["kernels"]["k_search"]
    assert bench.ncu_traffic("k_search", c2) == s["dram_bytes_per_launch"] and abs(s["dram_bytes_per_launch"] - s["dram_read"] - s["dram_write"]) <= 2
    assert bench.ncu_traffic("k_search", {**c2, "k": 27}) is None
    # 64-byte fills: between one and two 32-byte sectors per reference-semantics bit test of the committed bench line
    line = json.loads((ROOT / "profiles" / "r02_bench_n1.json").read_text())
    tests_ = line["kernels"]["n_probes"]
    assert 32 * tests_ < s["dram_bytes_per_launch"] < 64 * tests_


def test_committed_launch_summary_is_what_the_committed_launch_list_says():
    """profiles/r02_launches_summary.txt = scripts/launch_summary.py over profiles/r02_launches.csv (the ncu launch list of
    bench.py --steps 2 --warmup 1): the kernel shares DESIGN.md quotes come from the committed list"""
    r = subprocess.run([sys.executable, str(ROOT / "scripts" / "launch_summary.py"), str(ROOT / "profiles" / "r02_launches.csv")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == (ROOT / "profiles" / "r02_launches_summary.txt").read_text().strip()
    assert "k_search<0, 2, 5, 0>" in r.stdout and "k_bin_apply2<2048, 1>" in r.stdout and "k_bin_scatter2<96>" in r.stdout
