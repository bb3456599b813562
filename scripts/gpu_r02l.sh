#!/bin/bash
# round 2, call l: balanced region ownership: multi-rank parity on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r02l_multi_tests.txt 2>&1; echo "multi tests rc=$?"; tail -8 gpurun_out/r02l_multi_tests.txt
COMMET_B200_TRACE=1 COMMET_B200_DEVICES=0 bash -c 'cd /tmp && python - <<PY
import sys, time, subprocess, numpy as np
sys.path.insert(0, "'$PWD'")
import torch, bench
from commet_b200 import build
n, L = 10_000_000, 100
dev = torch.device("cuda", 0)
ref, qry, offs = bench.make_sets_torch(n, L, 0, dev)
import pathlib, tempfile
td = pathlib.Path(tempfile.mkdtemp(dir="/dev/shm"))
for name, arr in (("ref", ref), ("qry", qry)):
    a = arr.view(n, L).cpu().numpy()
    rows = np.empty((n, 1 + 8 + 1 + L + 1), dtype=np.uint8); rows[:, 0] = ord(">")
    idx = np.arange(n, dtype=np.int64)
    for d in range(8): rows[:, 1 + d] = (idx // 10 ** (7 - d)) % 10 + 48
    rows[:, 9] = 10; rows[:, 10:10 + L] = a; rows[:, -1] = 10
    rows.tofile(td / f"{name}.fa"); (td / f"{name}.txt").write_text(f"{name}:{td}/{name}.fa\n")
del ref, qry; torch.cuda.empty_cache()
for i in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([str(build.BIN / "index_and_search"), "-i", str(td / "ref.txt"), "-s", str(td / "qry.txt"), "-o", str(td / "out"), "-l", str(td / "out"), "-k", "33"], capture_output=True, text=True)
    print("tool wall", round(time.perf_counter() - t0, 3), "rc", r.returncode)
    print(r.stderr[-3000:])
PY' > gpurun_out/r02l_tool_trace.txt 2>&1; echo "tool trace rc=$?"; tail -40 gpurun_out/r02l_tool_trace.txt
