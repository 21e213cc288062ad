#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python tests/large_cfg5.py > gpurun_out/r2_cfg5_sweep_1gpu.jsonl 2> gpurun_out/cfg5.err
echo "cfg5 exit $?"; tail -3 gpurun_out/cfg5.err; python - <<'PY'
import json
for l in open('gpurun_out/r2_cfg5_sweep_1gpu.jsonl'):
    d=json.loads(l)
    ft=d.get('frobenius_tensor')
    print(d['arm'],d['depth'],d['n_models'],'ctx %.0f'%d['mean_contexts'],'fill %.3f'%d['universe_fill'],'| frob cuda %.3f ms ok=%s'%(d['frobenius_cuda']['ms'],d['frobenius_cuda']['corner_within_bound_of_c_oracle']),'| tensor',('%.3f ms ok=%s'%(ft['ms'],ft['corner_within_bound_of_c_oracle']) if ft else None),'| picks',d['dispatcher_picks'],'| nll %.3f ms %s ok=%s'%(d['nll']['ms'],d['nll']['kernel'][:6],d['nll']['within_1e-9']))
PY
