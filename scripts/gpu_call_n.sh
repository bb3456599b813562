#!/bin/bash
# N-GPU bench only.  usage: gpurun --gpus N --timeout 400 -- 'bash scripts/gpu_call_n.sh N tag'
n=${1:-4}; tag=${2:-r01e}
out=gpurun_out; mkdir -p $out
python -c "import __graft_entry__ as g; g.build()" > $out/${tag}_build.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 3 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
echo "bench n$n rc=$?"; tail -c 1300 $out/${tag}_bench_n$n.json; grep -v "OMP_NUM\|\*\*\*\*" $out/${tag}_bench_n$n.err | tail -8
