#!/bin/bash
# round 2, call v: 32-bit-key search variants (k <= 30): parity under each, then k=27 timing; default (2 positions, 5 blocks) at k=33
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
COMMET_B200_SEARCH_VARIANT=325 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "search or chunk or probe or selection or upload" 2>&1 | tail -2
COMMET_B200_SEARCH_VARIANT=346 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "search or chunk or probe" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "search or chunk or probe or selection or upload" 2>&1 | tail -2
for v in 0 325 326 345 346 44; do
  COMMET_B200_SEARCH_VARIANT=$v timeout 300 python bench.py -k 27 --steps 3 --warmup 1 --no-cpu --no-extra > gpurun_out/r02v_k27_v$v.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02v_k27_v$v.json'));print('k27 variant $v', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2), round(d['roofline']['frac_of_random_sector_ceiling'],3))"
done
for v in 0 44; do
  COMMET_B200_SEARCH_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r02v_k33_v$v.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02v_k33_v$v.json'));print('k33 variant $v', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2), round(d['e2e']['ms_per_step'],2))"
done
