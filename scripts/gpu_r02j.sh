#!/bin/bash
# round 2, call j: after the plan / query-part changes: parity subset, then the N=1 bench line with its extra legs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py -q -m gpu -x -k "uploaded_parts or chunk_loop or distributed or group or upload_async or empty" > gpurun_out/r02j_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02j_tests.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02j_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','kernels')}); print(d['e2e']); print(d.get('roofline')); print(json.dumps(d.get('extra'), indent=1)[:3000]); print(d.get('cpu_baseline'))
PY
