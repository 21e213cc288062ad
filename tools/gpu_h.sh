#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== default L=10000"; VGS_G8_PROF=1 timeout 300 python tools/i8_tune.py 1 2>&1 | tail -24
echo "== L=6000 (no frequent k-mers?)"; TUNE_L=6000 VGS_G8_PROF=1 timeout 300 python tools/i8_tune.py 1 2>&1 | tail -6
echo "== skip functor"; VGS_G8_DEBUG=1 VGS_G8_PROF=1 timeout 300 python tools/i8_tune.py 1 2>&1 | tail -6
} > gpurun_out/h_epi.log 2>&1
cat gpurun_out/h_epi.log | cut -c1-700
