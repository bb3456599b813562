#!/bin/bash
# round 2, call z: C4 at full size on 1 GPU with the final search kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python scripts/bench_c4.py --out gpurun_out/r02z_c4_full_n1.json > gpurun_out/r02z_c4_full_n1.log 2>&1; echo "c4 n1 rc=$?"; tail -c 1200 gpurun_out/r02z_c4_full_n1.log
