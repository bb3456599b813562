#!/bin/bash
# round 2, call aa: the encode rewritten (left shifts on the FMA pipe, funnel-shift gather): parity subset, step time, C5 sweep, launch list of the sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "index_filter_bit_exact or filter_reads or kmer_counts or upload_async or selection" 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r02aa_bench.json 2> /dev/null
python -c "import json;d=json.load(open('gpurun_out/r02aa_bench.json'));print('step', round(d['ms_per_step'],2), d['kernels']['index_ms'], d['kernels']['search_ms'], round(d['e2e']['ms_per_step'],2))"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'k_stage_filter|k_filter|k_encode|k_bvop|k_popcount' -c 40 --csv --log-file gpurun_out/r02aa_c5_launches.csv \
    python scripts/sweep_c5.py --batches 1 --reps 1 > gpurun_out/r02aa_c5_ncu.log 2>&1; echo "c5 launch list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02aa_c5_launches.csv')) if len(r)>10]
h=rows[0]; iN=h.index('Kernel Name'); iM=h.index('Metric Name'); iV=h.index('Metric Value'); iI=h.index('ID'); iG=h.index('Grid Size')
d={}
for r in rows[1:]:
    d.setdefault(r[iI],{'k':r[iN].split('(')[0],'g':r[iG]})[r[iM]]=r[iV]
for k,v in d.items(): print(k, v)
PY
timeout 600 python scripts/sweep_c5.py > gpurun_out/r02aa_c5_sweep.json 2> gpurun_out/r02aa_c5_sweep.err; echo "c5 sweep rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02aa_c5_sweep.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a!='note' and not isinstance(b,(dict,list))})
PY
