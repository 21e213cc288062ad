#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo skip tests > gpurun_out/a_gpu_tests.log
echo "gpu tests exit $?" >> gpurun_out/a_gpu_tests.log
tail -8 gpurun_out/a_gpu_tests.log
( time timeout 900 python bench.py ) 2>> gpurun_out/a_time.txt > gpurun_out/a_bench_nll.json 2> gpurun_out/a_bench_nll.err
echo "bench exit $?"; cat gpurun_out/a_time.txt; tail -3 gpurun_out/a_bench_nll.err | cut -c1-300; head -c 4500 gpurun_out/a_bench_nll.json
