#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python tools/i8_tune.py 3"
$CMD > gpurun_out/ncu3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_i8_hist|k_nll_combine_full" -s 6 -c 4 -o gpurun_out/r2_hist_combine $CMD > gpurun_out/ncu3.log 2>&1
echo "exit $?"; tail -1 gpurun_out/ncu3_plain.log
