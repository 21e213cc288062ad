#!/bin/bash
# first GPU pass of round 2: new CTA-pair int8 GEMM + sharded NLL (emulated) + probe, then the whole GPU suite, then bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -s -k "probe" > gpurun_out/a_round2_first.log 2>&1
echo "round2-first exit $?" >> gpurun_out/a_round2_first.log
tail -15 gpurun_out/a_round2_first.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/a_gpu_tests.log
tail -15 gpurun_out/a_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --probe-models 0 > gpurun_out/a_bench_nll.json 2> gpurun_out/a_bench_nll.err
echo "bench exit $?"; tail -3 gpurun_out/a_bench_nll.err; head -c 3000 gpurun_out/a_bench_nll.json
