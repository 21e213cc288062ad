#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 120 python tools/i8_tune.py 10
} > gpurun_out/b_tune.log 2>&1
cat gpurun_out/b_tune.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/b_launches.csv python tools/i8_tune.py 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/b_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]; ki=cols.index('Kernel Name'); vi=cols.index('Metric Value')
for r in rows[hdr+1:][-8:]: print('%-60s %8.1f us'%(r[ki][:60], float(r[vi].replace(',',''))/1000))
PY
