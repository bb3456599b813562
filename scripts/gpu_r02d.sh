#!/bin/bash
# round 2, call d: launch list + full ncu capture of the insert kernels (second form) on C2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02d_launches_bench.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py gpurun_out/r02d_launches.csv > gpurun_out/r02d_launches_summary.txt 2>&1; cat gpurun_out/r02d_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_scatter2|k_bin_apply2' -s 2 -c 2 -f -o gpurun_out/r02d_insert \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r02d_full_bench.log 2>&1; echo "full capture rc=$?"; ls -la gpurun_out/r02d_insert.ncu-rep
