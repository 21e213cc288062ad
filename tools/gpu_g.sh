#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -5
timeout 300 python tools/i8_tune.py 10 2>&1 | tail -1
VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_laps.py 2>&1 | grep -v "batch" | grep "True}  \|'device_log': True}\|^{}" | tail -12
} > gpurun_out/g_val.log 2>&1
cat gpurun_out/g_val.log | cut -c1-400
