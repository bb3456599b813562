#!/bin/bash
# bench.py at N GPUs, launched as the driver launches it.  usage (on the box): bash scripts/gpu_bench_only.sh N tag
N=${1:-2}; tag=${2:-r02}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
out=gpurun_out; mkdir -p $out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
echo "bench n$N rc=$?"; python - <<PY
import json
try:
    d=json.load(open('$out/${tag}_bench_n$N.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','dist_mode','phases_ms_rank0','clocks')}, d['e2e'])
except Exception as e: print('no line', e); print(open('$out/${tag}_bench_n$N.err').read()[-1500:])
PY
