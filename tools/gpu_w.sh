#!/bin/bash
# bench lines of the secondary workloads (profiles/r2_bench_<workload>.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
for W in frobenius frobenius_union projection estimate nll_deep frobenius_deep; do
  timeout 600 python bench.py --workload $W --probe-models 0 --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_$W.json 2> gpurun_out/r2_bench_$W.err || tail -5 gpurun_out/r2_bench_$W.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_$W.json').read())
r=d['roofline']
print('$W','step ms %.3f'%d['ms_per_step'],'e2e ms %.3f'%d['e2e']['ms_per_step'],'| kernel',r['kernel'][:40],'ms %.3f'%r['kernel_ms_per_launch'],'bound',r['bound'],'frac',r['frac'])
"
done
timeout 600 python bench.py --mode sequential --probe-models 0 --no-cpu-baseline --no-mode-sweep --steps 5 > gpurun_out/r2_bench_nll_sequential.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_nll_sequential.json').read()); print('nll sequential', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'].get('smem_gather'), d['roofline']['frac'])"
} > gpurun_out/w_bench.log 2>&1
cat gpurun_out/w_bench.log | cut -c1-400
