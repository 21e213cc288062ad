#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
for K in "VGS_LOAD_NT=1" "VGS_LOAD_SPIN=1 VGS_LOAD_NT=1" "VGS_LOAD_SPIN=1 VGS_LOAD_BATCH=64" "VGS_LOAD_SPIN=1 VGS_LOAD_BATCH=16" "VGS_LOAD_SPIN=1 VGS_LOAD_NT=1 VGS_LOAD_BATCH=16"; do
echo "== $K"
env $K VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_laps.py 2>&1 | grep -B40 "'nll_only': True}  " | grep "nll_only\|ln p computed\|drained" | tail -6
done
} > gpurun_out/g_laps.log 2>&1
cat gpurun_out/g_laps.log | cut -c1-300
