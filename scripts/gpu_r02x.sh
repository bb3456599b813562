#!/bin/bash
# round 2, call x: final state -- the whole GPU suite, smoke, the N=1 line, launch list and full captures of the hot kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r02x_gpu_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/r02x_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02x_smoke.txt
timeout 900 python bench.py > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02x_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','kernels','clocks')}); print(d['e2e']); print(d['extra']['tool_e2e_c2'], d['extra']['k27_l2_resident'], d['extra']['c4_small'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02x_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/r02x_launches_bench.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/r02x_launches.csv > gpurun_out/r02x_launches_summary.txt 2>&1; cat gpurun_out/r02x_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter2|k_bin_apply2|k_search|k_encode' -s 4 -c 5 -f -o gpurun_out/r02x_full \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02x_full_bench.log 2>&1; echo "full capture rc=$?"; ls -la gpurun_out/r02x_full.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_search' -s 1 -c 1 -f -o gpurun_out/r02x_full_k27 \
    python bench.py -k 27 --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/r02x_full_k27_bench.log 2>&1; echo "k27 capture rc=$?"; ls -la gpurun_out/r02x_full_k27.ncu-rep
