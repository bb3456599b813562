#!/bin/bash
# 2-GPU call: bench N=2 (distributed placement), the 2-GPU parity tests, commet_nxn on 2 GPUs.
# usage: gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_call_b.sh r01d'
tag=${1:-r01d}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1200 $out/${tag}_bench_n2.json; tail -5 $out/${tag}_bench_n2.err
timeout 200 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_fullsize.py::test_c2_filter_sorted_insert_equals_direct_and_oracle_keys" \
    "tests/test_gpu_tools.py::test_commet_nxn_multi_gpu_matches_single" -q > $out/${tag}_tests_2gpu.txt 2>&1
echo "tests rc=$?"; tail -5 $out/${tag}_tests_2gpu.txt
