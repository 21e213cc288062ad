#!/bin/bash
# ncu evidence for the cfg-3 step: launch list (device time per launch) + one full capture of the int8 GEMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mode-sweep --probe-models 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_nll_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_g8_gemm -s 3 -c 2 -o gpurun_out/r2_g8_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_nll_cfg3.csv
tail -2 gpurun_out/ncu_plain.log | cut -c1-600
