#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "tensor_core" 2>&1 | tail -15
timeout 300 python tools/pair_tune.py trees 1000 6 500
timeout 300 python tools/pair_tune.py chains 1000 6
timeout 300 python tools/pair_tune.py chains 400 7
timeout 300 python tools/pair_tune.py trees 1000 7 500
} > gpurun_out/b_tune.log 2>&1
cat gpurun_out/b_tune.log | cut -c1-300
