#!/bin/bash
# round 2, call z2: the N=1 line once more with the final bench.py (roofline_index against the L2 RED.OR ceiling), and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02z_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','roofline_index','clocks')}); print(d['e2e']); print(d['cpu_baseline'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/r02z_bench_reference.json | cut -c1-1500
