#!/bin/bash
# round 2, call y (N GPUs): host-to-device bandwidth alone / together, unbound and bound to the GPU's node; then bench.py as the driver launches it
N=${1:-8}; tag=${2:-r02y}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 200 $TR 29521 scripts/h2d_concurrent.py --bind 0 > $out/${tag}_h2d_n$N.json 2> $out/${tag}_h2d_n$N.err; echo "h2d rc=$?"; cat $out/${tag}_h2d_n$N.json
timeout 200 $TR 29522 scripts/h2d_concurrent.py --bind 1 > $out/${tag}_h2d_bound_n$N.json 2> $out/${tag}_h2d_bound_n$N.err; echo "h2d bound rc=$?"; cat $out/${tag}_h2d_bound_n$N.json
nvidia-smi topo -m > $out/${tag}_topo_n$N.txt 2>&1; head -14 $out/${tag}_topo_n$N.txt
bash scripts/gpu_bench_only.sh $N $tag
