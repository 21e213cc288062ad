#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for W in 8 4 2; do
python tools/shard_tune.py $W 10
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/b_launches.csv python tools/shard_tune.py $W 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/b_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cols=rows[hdr]; ki=cols.index('Kernel Name'); vi=cols.index('Metric Value'); gi=cols.index('Grid Size')
for r in rows[hdr+1:][-5:]: print('   %-56s %-16s %8.1f us'%(r[ki][:56], r[gi], float(r[vi].replace(',',''))/1000))
PY
done
