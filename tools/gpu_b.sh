#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
VGS_G8_PROF=1 timeout 120 python tools/i8_tune.py 1
timeout 120 python tools/i8_tune.py 10
} > gpurun_out/b_tune.log 2>&1
grep -E "g8 prof|knobs|passed|failed|rror" gpurun_out/b_tune.log | cut -c1-400 | tail -14
