#!/bin/bash
# round 2, call n: block-span fused staging+selection (parity, sweep, kernel times), k=27 search capture, tool-level trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_tools.py -q -m gpu -x -k "filter_reads or c5 or several_ranks or full_mode" > gpurun_out/r02n_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02n_tests.txt
timeout 600 python scripts/sweep_c5.py > gpurun_out/r02n_c5_sweep.json 2> gpurun_out/r02n_c5_sweep.err; echo "c5 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_c5_sweep.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'ms' in v or (isinstance(v,dict) and 'stage_ms' in v): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note',) and not isinstance(b,dict)})
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_stage_filter|k_filter|k_encode|k_popcount|k_bvop' -c 40 --csv --log-file gpurun_out/r02n_c5_launches.csv \
    python scripts/sweep_c5.py --batches 1 --reps 1 > /dev/null 2>&1; echo "c5 launch list rc=$?"
python - <<'PY'
import csv,re,collections
rows=list(csv.reader(l for l in open('gpurun_out/r02n_c5_launches.csv') if l.startswith('"')))
h=rows[0]; ki,vi,ui,mi,ii=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit"),h.index("Metric Name"),h.index("ID")
acc=collections.OrderedDict()
for r in rows[1:]:
    acc.setdefault((r[ii], re.sub(r'\(.*','',r[ki])[:40]),{})[r[mi]]=(float(r[vi].replace(",","")), r[ui])
for (i,k),m in acc.items():
    t=m['gpu__time_duration.sum']; ms=t[0]/1e6 if t[1].startswith('n') else t[0]/1e3 if t[1].startswith('u') else t[0]
    def gb(x):
        v,u=m[x]; return v/1e9 if u=='byte' else v/1e3 if u.startswith('M') else v if u.startswith('G') else v/1e6 if u.startswith('K') else v
    print(f"{ms:9.3f} ms  rd {gb('dram__bytes_read.sum'):7.3f} GB  wr {gb('dram__bytes_write.sum'):7.3f} GB  {k}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search -s 40 -c 1 -f -o gpurun_out/r02n_search_k27 \
    python bench.py -k 27 --steps 1 --warmup 0 --no-cpu --no-extra > gpurun_out/r02n_search_k27.log 2>&1; echo "k27 capture rc=$?"; ls -la gpurun_out/r02n_search_k27.ncu-rep
COMMET_B200_TRACE=1 python - > gpurun_out/r02n_tool_trace.txt 2>&1 <<'PY'
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
for i in range(2):
    t0 = time.perf_counter()
    r = subprocess.run([str(build.BIN / "index_and_search"), "-i", str(td / "ref.txt"), "-s", str(td / "qry.txt"), "-o", str(td / "out"), "-l", str(td / "out"), "-k", "33"], capture_output=True, text=True)
    print("tool wall", round(time.perf_counter() - t0, 3), "rc", r.returncode)
    print("\n".join(l for l in r.stderr.split("\n") if "commet tool" in l))
PY
echo "tool trace rc=$?"; cat gpurun_out/r02n_tool_trace.txt | tail -24
