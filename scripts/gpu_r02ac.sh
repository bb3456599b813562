#!/bin/bash
# round 2, call ac: the tests that go through filter_reads with the fused kernel's new default (256 reads per block)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tools.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -q -m gpu -x -k "filter_reads or commet_flow or nxn_bit_exact or c5 or unmodified" 2>&1 | tail -3
