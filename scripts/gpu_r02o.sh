#!/bin/bash
# round 2, call o: level-synchronous verification in the search (parity, A/B at k=33 and k=27), k_stage_filter2 with 2 blocks/SM, tool trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tools.py -q -m gpu -x -k "search or chunk or probe or selection or filter_reads or vs_reference or commet_py or full_mode or known_answer" > gpurun_out/r02o_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02o_tests.txt
for lv in 0 1; do
  COMMET_B200_SEARCH_LEVEL=$lv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r02o_bench_k33_level$lv.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02o_bench_k33_level$lv.json'));print('k33 level $lv', d['ms_per_step'], d['kernels']['search_ms'], d['kernels']['index_ms'], d['e2e']['ms_per_step'])"
  COMMET_B200_SEARCH_LEVEL=$lv timeout 300 python bench.py -k 27 --steps 3 --warmup 1 --no-cpu --no-extra > gpurun_out/r02o_bench_k27_level$lv.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02o_bench_k27_level$lv.json'));print('k27 level $lv', d['ms_per_step'], d['kernels']['search_ms'], d['roofline']['frac_of_random_sector_ceiling'])"
  COMMET_B200_SEARCH_LEVEL=$lv timeout 600 python scripts/bench_c4.py --ref-reads 100000000 --out gpurun_out/r02o_c4_fifth_level$lv.json > /dev/null 2>&1
  python -c "import json;d=json.load(open('gpurun_out/r02o_c4_fifth_level$lv.json'));print('c4/5 level $lv', d['seconds'], d['phases_s_rank0'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_stage_filter|k_filter|k_encode' -c 14 --csv --log-file gpurun_out/r02o_c5_launches.csv \
    python scripts/sweep_c5.py --batches 1 --reps 1 > /dev/null 2>&1; grep -o '"[^"]*k_[a-z_0-9<>, ]*[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","gpu__time[^"]*","[^"]*","[^"]*"' gpurun_out/r02o_c5_launches.csv | awk -F'","' '{print $1, $NF}' | cut -c1-120
COMMET_B200_TRACE=1 python - > gpurun_out/r02o_tool_trace.txt 2>&1 <<'PY'
import sys, time, subprocess, numpy as np, pathlib, tempfile, os
sys.path.insert(0, os.getcwd())
import torch, bench
from commet_b200 import build
build.build_all()
n, L = 10_000_000, 100
dev = torch.device("cuda", 0)
ref, qry, offs = bench.make_sets_torch(n, L, 0, dev)
td = pathlib.Path(tempfile.mkdtemp(dir="/dev/shm"))
for name, arr in (("ref", ref), ("qry", qry)):
    a = arr.view(n, L).cpu().numpy()
    rows = np.empty((n, 1 + 8 + 1 + L + 1), dtype=np.uint8); rows[:, 0] = ord(">")
    idx = np.arange(n, dtype=np.int64)
    for d in range(8): rows[:, 1 + d] = (idx // 10 ** (7 - d)) % 10 + 48
    rows[:, 9] = 10; rows[:, 10:10 + L] = a; rows[:, -1] = 10
    rows.tofile(td / f"{name}.fa"); (td / f"{name}.txt").write_text(f"{name}:{td}/{name}.fa\n")
del ref, qry; torch.cuda.empty_cache()
for i in range(3):
    t0 = time.perf_counter()
    r = subprocess.run([str(build.BIN / "index_and_search"), "-i", str(td / "ref.txt"), "-s", str(td / "qry.txt"), "-o", str(td / "out"), "-l", str(td / "out"), "-k", "33"], capture_output=True, text=True)
    print("tool wall", round(time.perf_counter() - t0, 3), "rc", r.returncode)
    print("\n".join(l for l in r.stderr.split("\n") if "commet tool" in l))
PY
tail -24 gpurun_out/r02o_tool_trace.txt
