#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_laps.py > gpurun_out/b_laps.log 2>&1
tail -60 gpurun_out/b_laps.log
