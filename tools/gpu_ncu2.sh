#!/bin/bash
# ncu evidence for the tensor-core vs CUDA-core choice of the pair distances: both kernels on either side of the crossover
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for SHAPE in "trees 1000 6 500" "trees 1000 7 500"; do
  TAG=$(echo $SHAPE | tr ' ' '_')
  python tools/pair_tune.py $SHAPE > gpurun_out/ncu2_plain_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k_g8_gemm|k_pair_partials" -s 6 -c 6 -o gpurun_out/r2_pairs_$TAG python tools/pair_tune.py $SHAPE > gpurun_out/ncu2_$TAG.log 2>&1
  echo "$TAG exit $?"
  tail -2 gpurun_out/ncu2_plain_$TAG.log | cut -c1-250
done
ls -la gpurun_out/r2_pairs_*.ncu-rep
