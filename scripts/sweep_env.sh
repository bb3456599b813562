#!/bin/bash
# A/B sweep of the launch-shape knobs the library reads from the environment (no rebuild): one bench.py run per
# setting, one line per run.  GPU box only.   usage: gpurun --timeout 900 -- 'bash scripts/sweep_env.sh [k]'
k=${1:-33}
out=gpurun_out; mkdir -p $out
run() {   # name, env assignments...
  name=$1; shift
  env "$@" timeout 120 python bench.py -k $k --steps 3 --warmup 2 --no-cpu > $out/sweep_$name.json 2> $out/sweep_$name.err
  python - $out/sweep_$name.json "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    kk = d.get("kernels", {})
    print(f"{sys.argv[2]:28s} step {d['ms_per_step']:7.2f} ms  index {kk.get('index_ms', 0):6.2f}  search {kk.get('search_ms', 0):6.2f}  e2e {d['e2e']['ms_per_step']:7.2f}")
except Exception as e:
    print(f"{sys.argv[2]:28s} failed: {e}")
PY
}
run default
for v in 1 2 3; do run scatter_bps_$v COMMET_B200_SCATTER_BPS=$v; done
for v in 2 3 4 6; do run count_bps_$v COMMET_B200_COUNT_BPS=$v; done
for v in 2048 4096 8192; do run apply_tile_$v COMMET_B200_APPLY_TILE=$v; done
for v in 4 6 8 12 16; do run apply_bps_$v COMMET_B200_APPLY_BPS=$v; done
run apply_noprefetch COMMET_B200_APPLY_PREFETCH=0
for v in 64 128 256 512 1024; do run search_bps_$v COMMET_B200_SEARCH_BPS=$v; done
for v in 0 2 3 4 6 8; do run search_both_$v COMMET_B200_SEARCH_BOTH=$v; done
for v in 3 4 8; do run search_dynamic_$v COMMET_B200_SEARCH_DYNAMIC=$v; done
