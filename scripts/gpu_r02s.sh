#!/bin/bash
# round 2, call s: k_search compiled for 3 / 4 / 5 / 6 resident blocks per SM, at k=33 and k=27
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for mb in 1 4 5 6; do
  COMMET_B200_SEARCH_MINB=$mb timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r02s_k33_minb$mb.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02s_k33_minb$mb.json'));print('k33 minb $mb', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2))"
  COMMET_B200_SEARCH_MINB=$mb timeout 300 python bench.py -k 27 --steps 3 --warmup 1 --no-cpu --no-extra > gpurun_out/r02s_k27_minb$mb.json 2> /dev/null
  python -c "import json;d=json.load(open('gpurun_out/r02s_k27_minb$mb.json'));print('k27 minb $mb', round(d['ms_per_step'],2), round(d['kernels']['search_ms'],2), round(d['roofline']['frac_of_random_sector_ceiling'],3))"
done
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "search or chunk_loop or popcount_of_several" 2>&1 | tail -2
