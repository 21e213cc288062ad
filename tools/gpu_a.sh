#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/a_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/a_gpu_tests.log | tail -20
