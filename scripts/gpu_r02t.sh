#!/bin/bash
# round 2, call t: smaller probe batches with more resident warps, at k=27 (L2-resident) and k=33
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for v in 25 24 16 18; do
  COMMET_B200_SEARCH_VARIANT=$v timeout 300 python bench.py -k 27 --steps 3 --warmup 1 --no-cpu --no-extra > gpurun_out/r02u_k27_v$v.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02u_k27_v$v.json'));print('k27 variant $v', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2), round(d['roofline']['frac_of_random_sector_ceiling'],3))"
  COMMET_B200_SEARCH_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r02u_k33_v$v.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02u_k33_v$v.json'));print('k33 variant $v', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2))"
done
