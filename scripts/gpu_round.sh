#!/bin/bash
# One gpurun call that refreshes everything the round's profiles/ cite, most important first:
#   bench line (N=1), GPU parity tests, ncu launch list, ncu --set full of the hot kernels, reference arm, smoke.
# usage (here):  gpurun --timeout 1100 -- 'bash scripts/gpu_round.sh r01b'
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1

timeout 400 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"; tail -c 600 $out/${tag}_bench_n1.json

timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.txt 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_gpu_tests.txt

timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/${tag}_launches_bench.log 2>&1
echo "launch list rc=$?"

timeout 420 ncu --set full --clock-control none --import-source on -k "regex:k_search|k_bin_apply|k_bin_scatter|k_bin_count" -c 4 \
    -f -o $out/${tag}_full python bench.py --steps 1 --warmup 0 --no-cpu > $out/${tag}_full_bench.log 2>&1
echo "ncu full rc=$?"; ls -la $out/${tag}_full.ncu-rep

timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
echo "reference arm rc=$?"; tail -c 300 $out/${tag}_bench_reference.json

timeout 200 python __graft_entry__.py smoke > $out/${tag}_smoke.txt 2>&1
echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.txt
