#!/bin/bash
# 1-GPU validation of the new tests + secondary bench lines.  usage: gpurun --timeout 900 -- 'bash scripts/gpu_call_a.sh r01c'
tag=${1:-r01c}
out=gpurun_out
mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 420 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_multi.py "tests/test_gpu_kernels.py::test_upload_async_streams_behave_like_uploaded_ones" -q --durations=12 > $out/${tag}_new_tests.txt 2>&1
echo "new tests rc=$?"; tail -25 $out/${tag}_new_tests.txt
timeout 200 python bench.py -k 27 --steps 2 --warmup 1 --no-cpu > $out/${tag}_bench_k27.json 2> $out/${tag}_bench_k27.err
echo "k27 rc=$?"; tail -c 400 $out/${tag}_bench_k27.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference rc=$?"; tail -c 900 $out/${tag}_bench_reference.json
