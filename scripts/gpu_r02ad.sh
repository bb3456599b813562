#!/bin/bash
# round 2, call ad: the multi-rank loop after its chunk plan was split into a counts-source template (device / host): thread groups, then processes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "group_of_one_process" 2>&1 | tail -2
timeout 150 python -m pytest tests/test_gpu_multi.py -q -m gpu -x -k "distributed_reference_set_ranks" 2>&1 | tail -2
