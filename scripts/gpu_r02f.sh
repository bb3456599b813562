#!/bin/bash
# round 2, call f: filter_reads tests (fused staging+selection, border overflow), C5 sweep + its launch list, insert A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_tools.py -q -m gpu -x -k "filter_reads or c5 or bvop or l2_blocked" > gpurun_out/r02f_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02f_tests.txt
timeout 600 python scripts/sweep_c5.py > gpurun_out/r02f_c5_sweep.json 2> gpurun_out/r02f_c5_sweep.err; echo "c5 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_c5_sweep.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note',) and not isinstance(b,dict)})
print({k:round(v['frac_of_hbm_peak'],3) for k,v in d['bvop'].items() if isinstance(v,dict)})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_stage_filter|k_filter|k_encode|k_popcount|k_bvop' -c 60 --csv --log-file gpurun_out/r02f_c5_launches.csv \
    python scripts/sweep_c5.py --batches 1 --reps 1 > /dev/null 2>&1; echo "c5 launch list rc=$?"
python - <<'PY'
import csv,re,collections
rows=list(csv.reader(l for l in open('gpurun_out/r02f_c5_launches.csv') if l.startswith('"')))
h=rows[0]; ki,vi,ui=h.index("Kernel Name"),h.index("Metric Value"),h.index("Metric Unit")
for r in rows[1:]:
    v=float(r[vi].replace(",","")); v=v/1e6 if r[ui].startswith("n") else v/1e3 if r[ui].startswith("u") else v
    print(f"{v:9.3f} ms  {re.sub(r'\(.*','',r[ki])[:60]}")
PY
timeout 600 python scripts/ab_index.py COMMET_B200_S2_TW=128 COMMET_B200_S2_TW=96 COMMET_B200_S2_TW=64 COMMET_B200_S2_TW=96,COMMET_B200_SCATTER_BPS=2 > gpurun_out/r02f_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02f_ab.txt
