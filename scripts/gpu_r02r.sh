#!/bin/bash
# round 2, call r: after the last host-side changes: tools and multi-rank parity again
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 2000 python -m pytest tests/test_gpu_tools.py tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r02r_tests.txt 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02r_tests.txt
