#!/bin/bash
# round 2, call c: insert A/B on C2 + group / dist tests on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python scripts/ab_index.py COMMET_B200_INSERT=1 COMMET_B200_INSERT=2,COMMET_B200_S2_TW=128 COMMET_B200_S2_TW=96 COMMET_B200_S2_TW=64 COMMET_B200_S2_TW=128,COMMET_B200_APPLY_TILE=4096 COMMET_B200_APPLY_TILE=2048,COMMET_B200_APPLY_BPS=6 COMMET_B200_APPLY_BPS=8,COMMET_B200_APPLY_PREFETCH=0 > gpurun_out/r02c_ab.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r02c_ab.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r02c_multi_tests.txt 2>&1; echo "multi tests rc=$?"; tail -15 gpurun_out/r02c_multi_tests.txt
timeout 600 python -m pytest tests/test_gpu_tools.py tests/test_gpu_kernels.py -q -m gpu -x -k "several_ranks or small_and_large_k" > gpurun_out/r02c_tests.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02c_tests.txt
