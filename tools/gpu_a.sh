#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/a_gpu_tests.log
tail -8 gpurun_out/a_gpu_tests.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/a_bench_nll.json 2> gpurun_out/a_bench_nll.err
echo "bench exit $?"; grep -E "Elapsed|Maximum resident" gpurun_out/a_bench_nll.err; tail -3 gpurun_out/a_bench_nll.err | cut -c1-300; head -c 4500 gpurun_out/a_bench_nll.json
