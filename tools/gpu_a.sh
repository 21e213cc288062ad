#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/a_smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/a_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/a_gpu_tests.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/a_gpu_tests.log | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_cfg3_n1.json 2> gpurun_out/a_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2>> gpurun_out/a_bench.err; echo "ref exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_cfg3_n1.json').read()); r=d['roofline']
print('step',d['ms_per_step'],'value %.3e'%d['value'],'e2e ms',d['e2e']['ms_per_step'],'e2e value %.3e'%d['e2e']['value'],'frac',r['frac'],'frac2xbf16',r['frac_of_2x_bf16_burst'],'gemm ms',r['kernel_ms_per_launch'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print(d['scale_probe']); print(d['cpu_baseline'])
ref=json.loads(open('gpurun_out/r2_bench_reference.json').read()); print('reference', ref['value'], ref['cpu_baseline']['cores'], 'ratio e2e %.0f'%(d['e2e']['value']/ref['value']))
PY
