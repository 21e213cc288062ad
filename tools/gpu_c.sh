#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_breakdown.py > gpurun_out/c_e2e.log 2>&1
grep -v "^\[vgs_load" gpurun_out/c_e2e.log | tail -12; grep "^\[vgs_load" gpurun_out/c_e2e.log | tail -5; nproc; lscpu | grep -E "Model name|^CPU\(s\)|MHz" | head -4
