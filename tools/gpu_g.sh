#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -12
timeout 900 python bench.py --probe-models 0 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench exit $?"; tail -3 gpurun_out/g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g_bench.json').read()); r=d['roofline']
print('step',d['ms_per_step'],'value %.3e'%d['value'],'e2e ms',d['e2e']['ms_per_step'],'e2e value %.3e'%d['e2e']['value'],'frac',r['frac'],'gemm ms',r['kernel_ms_per_launch'],'nd',r.get('digit_rows_per_model'),'launches',d['gpu_launches'], d['config'].get('ms_per_step_by_nll_mode'))
PY
} > gpurun_out/g_val.log 2>&1
cat gpurun_out/g_val.log | cut -c1-400
