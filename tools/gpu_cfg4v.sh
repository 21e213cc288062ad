#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python tests/large_cfg4.py --verify --steps 2 > gpurun_out/r2_cfg4_n20000_verify_1gpu.json 2> gpurun_out/cfg4v.err ) 2> gpurun_out/cfg4v_time.txt
echo "cfg4 verify exit $?"; cat gpurun_out/cfg4v_time.txt | tail -4; tail -3 gpurun_out/cfg4v.err | cut -c1-300; cat gpurun_out/r2_cfg4_n20000_verify_1gpu.json | cut -c1-3000
