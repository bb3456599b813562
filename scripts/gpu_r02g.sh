#!/bin/bash
# round 2, call g: compacted-list search: parity (chunk loops, probe counts, C4 twin), then k=27 / k=33 A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py tests/test_gpu_fullsize.py -q -m gpu -x -k "chunk or probe or selection or c4 or group or distributed or upload_async" > gpurun_out/r02g_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02g_tests.txt
for c in 0 1; do
  COMMET_B200_SEARCH_COMPACT=$c timeout 300 python bench.py -k 27 --steps 3 --warmup 1 --no-cpu > gpurun_out/r02g_bench_k27_compact$c.json 2> gpurun_out/r02g_bench_k27_compact$c.err; echo "k27 compact=$c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02g_bench_k27_compact$c.json'));print(d['ms_per_step'], d['kernels'], d['roofline']['frac_of_random_sector_ceiling'])"
done
for c in 0 1; do
  COMMET_B200_SEARCH_COMPACT=$c timeout 600 python scripts/bench_c4.py --ref-reads 100000000 --out gpurun_out/r02g_c4_fifth_compact$c.json > /dev/null 2> gpurun_out/r02g_c4_fifth_compact$c.err; echo "c4/5 compact=$c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r02g_c4_fifth_compact$c.json'));print(d['seconds'], d['chunks'], d['phases_s_rank0'], list(d['shared_per_set'].values())[:3])"
done
