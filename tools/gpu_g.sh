#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -5
for rep in 1 2; do
VGS_PDL=0 timeout 120 python tools/i8_tune.py 20 2>&1 | tail -1
timeout 120 python tools/i8_tune.py 20 2>&1 | tail -1
done
for W in 8 2; do VGS_PDL=0 timeout 300 python tools/shard_tune.py $W 10 2>&1 | tail -1; timeout 300 python tools/shard_tune.py $W 10 2>&1 | tail -1; done
} > gpurun_out/g_val.log 2>&1
cat gpurun_out/g_val.log | cut -c1-300
