#!/bin/bash
# A/B of the dynamic read hand-out search (COMMET_B200_SEARCH_DYNAMIC=3|4) against the default, k=33 and k=27,
# after its parity tests.  usage: gpurun --timeout 500 -- 'bash scripts/gpu_call_c.sh r01g'
tag=${1:-r01g}
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
for d in 3; do
  COMMET_B200_SEARCH_DYNAMIC=$d timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x \
      -k "search_against or small_and_large or chunk_loop or selection or k33 or empty" > $out/${tag}_tests_dyn$d.txt 2>&1
  echo "tests dyn=$d rc=$?"; tail -2 $out/${tag}_tests_dyn$d.txt
done
for kd in "27 3" "27 4" "33 4"; do
  set -- $kd; k=$1; d=$2
  for once in 1; do
    COMMET_B200_SEARCH_DYNAMIC=$d timeout 120 python bench.py -k $k --steps 2 --warmup 1 --no-cpu > $out/${tag}_bench_k${k}_dyn$d.json 2> $out/${tag}_bench_k${k}_dyn$d.err
    python - $out/${tag}_bench_k${k}_dyn$d.json $k $d <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(f"k={sys.argv[2]} dyn={sys.argv[3]}: step {d['ms_per_step']:.2f} ms  search {d['kernels']['search_ms']:.2f} ms  index {d['kernels']['index_ms']:.2f} ms  shared {d['shared_reads']}")
except Exception as e:
    print("failed", sys.argv[1:], e)
PY
  done
done
