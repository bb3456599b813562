#!/bin/bash
# round 2, call ab: fused staging + selection kernel with 4 blocks of 256 reads per SM against 2 blocks of 512 (parity subset under both, C5 timings)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for th in 256 512; do
  COMMET_B200_SF_THREADS=$th timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "filter_reads" 2>&1 | tail -1
  COMMET_B200_SF_THREADS=$th timeout 300 python scripts/sweep_c5.py --batches 2 --reps 5 > gpurun_out/r02ab_c5_$th.json 2> gpurun_out/r02ab_c5_$th.err; echo "sweep $th rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02ab_c5_$th.json'))
for k in ('filter_reads_selection_only','filter_reads_fused_with_staging','filter_reads_two_kernels'):
    v=d[k]; print('$th', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','stage_ms','filter_ms','algorithmic_GBps','frac_of_hbm_peak','selected')})
PY
done
