#!/bin/bash
# multi-GPU call: usage (on the box): bash scripts/gpu_multi.sh N tag [c3]
#   bench.py at N GPUs (driver-style launch), C4 at full size over N GPUs, the real-N-GPU parity tests (N=2), C3 at full
#   size with commet_nxn --gpus N when the third argument is "c3"
N=${1:-2}; tag=${2:-r02}; c3=${3:-}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 400 $TR 29511 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err
echo "bench n$N rc=$?"; python - <<PY
import json
try:
    d=json.load(open('$out/${tag}_bench_n$N.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','dist_mode','phases_ms_rank0')}, d['e2e'])
except Exception as e: print('no line', e); print(open('$out/${tag}_bench_n$N.err').read()[-1500:])
PY
if [ "$N" = "8" ]; then
  COMMET_B200_DIST_MODE=merge timeout 400 $TR 29513 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_n${N}_mergemode.json 2> $out/${tag}_bench_n${N}_mergemode.err
  python - <<PY
import json
for f in ('mergemode',):
    try:
        d=json.load(open('$out/${tag}_bench_n${N}_%s.json' % f)); print(f, d['ms_per_step'], d.get('dist_mode'), d.get('phases_ms_rank0'), d['e2e']['ms_per_step'])
    except Exception as e: print(f, 'no line', e)
PY
fi
timeout 900 $TR 29514 scripts/bench_c4.py --out $out/${tag}_c4_full_n$N.json > $out/${tag}_c4_full_n$N.log 2>&1
echo "c4 n$N rc=$?"; tail -c 1500 $out/${tag}_c4_full_n$N.log
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_tools.py::test_commet_nxn_multi_gpu_matches_single" -q -m gpu > $out/${tag}_tests_${N}gpu.txt 2>&1
  echo "tests rc=$?"; tail -5 $out/${tag}_tests_${N}gpu.txt
fi
if [ "$c3" = "c3" ]; then
  timeout 900 python scripts/bench_nxn.py --sets 10 --reads 20000000 --gpus $N --reps 1 --out $out/${tag}_nxn_c3_full_${N}gpu.json > $out/${tag}_nxn_c3_${N}gpu.log 2>&1
  echo "c3 n$N rc=$?"; python -c "import json;d=json.load(open('$out/${tag}_nxn_c3_full_${N}gpu.json'));print(d['commet_nxn_run0']['wall_s'], d['commet_nxn_run0']['seconds_at_end_of'], d['generate_s'])" || tail -5 $out/${tag}_nxn_c3_${N}gpu.log
fi
if [ "$N" = "2" ] || [ "$N" = "8" ]; then
  timeout 300 python scripts/group_bench.py --gpus $N > $out/${tag}_group_n$N.json 2> $out/${tag}_group_n$N.err; echo "group n$N rc=$?"; cat $out/${tag}_group_n$N.json
fi
if [ "$N" = "2" ]; then
  # single-pass metrics (no kernel replay: the kernels write peer memory) of the NVLink kernels in the one-process group
  timeout 600 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:'k_owner_apply|k_gather_slices|k_owner_tiles|k_owner_fills|k_bin_scatter2' -c 24 --csv --log-file $out/${tag}_ncu_nvlink_n$N.csv \
      python scripts/group_bench.py --gpus $N --steps 1 --warmup 1 > /dev/null 2> $out/${tag}_ncu_nvlink_n$N.err; echo "ncu nvlink rc=$?"; tail -12 $out/${tag}_ncu_nvlink_n$N.csv | cut -c1-300
  COMMET_B200_DIST_MODE=merge timeout 600 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum --clock-control none \
      -k regex:'k_merge_peers' -c 4 --csv --log-file $out/${tag}_ncu_merge_n$N.csv \
      python scripts/group_bench.py --gpus $N --steps 1 --warmup 1 > /dev/null 2> $out/${tag}_ncu_merge_n$N.err; echo "ncu merge rc=$?"; tail -6 $out/${tag}_ncu_merge_n$N.csv | cut -c1-300
fi
