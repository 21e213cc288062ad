#!/bin/bash
# ncu launch list (device time per launch) of the cfg-3 bench command, after the same command has exited 0 without ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mode-sweep --probe-models 0"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_nll_cfg3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
tail -1 gpurun_out/ncu_plain.log | cut -c1-300
