#!/bin/bash
# multi-GPU pass: parity on real hardware (tests/mgpu_check.py) + bench at this GPU count
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_smi.txt
[ -z "$SKIP_CHECK" ] && timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/m${N}_check.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/m${N}_check.log
grep -E "mgpu|exit|rror" gpurun_out/m${N}_check.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/m${N}_bench.err | cut -c1-400; grep "^{" gpurun_out/m${N}_bench.json | head -c 6000
