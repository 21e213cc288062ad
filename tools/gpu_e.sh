#!/bin/bash
# validation of a GEMM / contraction change: the parity tests that touch it, then timing with and without in-kernel counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -15
echo "== tune"
timeout 300 python tools/i8_tune.py 10 2>&1 | tail -2
VGS_I8_FRAC_BITS=51 timeout 300 python tools/i8_tune.py 10 2>&1 | tail -2
VGS_G8_PROF=1 timeout 300 python tools/i8_tune.py 2 2>&1 | grep -v "pair [0-9]*: main" | tail -8
for W in 8 4 2; do timeout 300 python tools/shard_tune.py $W 10 2>&1 | tail -1; done
} > gpurun_out/e_val.log 2>&1
cat gpurun_out/e_val.log | cut -c1-900
