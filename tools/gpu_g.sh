#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mode-sweep --probe-models 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_nll_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
} > gpurun_out/g_val.log 2>&1
cat gpurun_out/g_val.log | cut -c1-300
