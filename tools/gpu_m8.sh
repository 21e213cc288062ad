#!/bin/bash
# all GPUs of the box: cfg 4 (N = 20 000) timing + checksums, multi-GPU parity check, bench line at this GPU count
N=${1:-8}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m${N}_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/large_cfg4.py > gpurun_out/r2_cfg4_n20000_${N}gpu.json 2> gpurun_out/m${N}_cfg4.err
echo "cfg4 exit $?"; tail -2 gpurun_out/m${N}_cfg4.err | cut -c1-300; grep "^{" gpurun_out/r2_cfg4_n20000_${N}gpu.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > gpurun_out/m${N}_check.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/m${N}_check.log
grep -E "mgpu|exit|rror" gpurun_out/m${N}_check.log | tail -16
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "bench exit $?"; tail -2 gpurun_out/m${N}_bench.err | cut -c1-300; grep "^{" gpurun_out/m${N}_bench.json | head -c 5000
