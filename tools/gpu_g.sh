#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 120 python tools/i8_tune.py 10 2>&1 | tail -1
VGS_G8_TSA=1 timeout 120 python tools/i8_tune.py 10 2>&1 | tail -3
VGS_G8_TSA=1 VGS_G8_PROF=1 timeout 120 python tools/i8_tune.py 2 2>&1 | grep -v "pair [0-9]*: main" | tail -3
} > gpurun_out/g_val.log 2>&1
cat gpurun_out/g_val.log | cut -c1-700
