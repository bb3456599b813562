#!/bin/bash
# A/B of cudaLimitMaxL2FetchGranularity on the random-sector ceilings and on the search kernel. GPU box only.
for g in 128 64 32; do
  echo "== COMMET_B200_L2_FETCH=$g"
  COMMET_B200_L2_FETCH=$g python scripts/microbench.py
  COMMET_B200_L2_FETCH=$g python bench.py --steps 3 --warmup 3 --no-cpu
  COMMET_B200_L2_FETCH=$g python bench.py --steps 3 --warmup 3 --no-cpu --direct-index
done
