#!/bin/bash
# round 2, call h: owner-applied insert: multi-rank parity on one GPU (IPC and in-process), both modes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r02h_multi_tests.txt 2>&1; echo "multi tests rc=$?"; tail -15 gpurun_out/r02h_multi_tests.txt
COMMET_B200_DIST_MODE=merge timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "29 or 30 or 28 or 31" > gpurun_out/r02h_multi_tests_merge.txt 2>&1; echo "merge-mode tests rc=$?"; tail -3 gpurun_out/r02h_multi_tests_merge.txt
