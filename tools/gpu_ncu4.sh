#!/bin/bash
# DRAM traffic per launch of the dominant CUDA-core kernels (profiles/traffic.json entries)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for W in frobenius_union estimate frobenius_deep nll_deep; do
  K=k_pair_partials; [ $W = estimate ] && K=k_est_pairs; [ $W = nll_deep ] && K=k_nll_scores
  CMD="python bench.py --workload $W --probe-models 0 --no-cpu-baseline --steps 3 --warmup 3"
  $CMD > gpurun_out/ncu4_plain_$W.log 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:$K -s 4 -c 2 --csv --log-file gpurun_out/r2_traffic_$W.csv $CMD > /dev/null 2>&1
  echo "$W exit $?"
done
tail -n 20 gpurun_out/r2_traffic_*.csv | cut -c1-260
