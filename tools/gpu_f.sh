#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_edges.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 python tools/i8_tune.py 10 2>&1 | tail -1
TUNE_L=6000 timeout 300 python tools/i8_tune.py 10 2>&1 | tail -1
for W in 8 2; do timeout 300 python tools/shard_tune.py $W 10 2>&1 | tail -1; done
echo "== laps"
VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_laps.py 2>&1 | grep -A7 "stream drained" | grep -B7 "'nll_only': True}  " | tail -16
} > gpurun_out/f_val.log 2>&1
cat gpurun_out/f_val.log | cut -c1-400
