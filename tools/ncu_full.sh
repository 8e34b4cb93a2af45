#!/bin/bash
# ncu --set full capture of selected kernels from the bench step (one GPU; SNB200_NO_GRAPH=1: same kernels, launched one by one).
# usage: tools/ncu_full.sh <tag> <kernel-regex> [count] [skip]
mkdir -p gpurun_out
TAG=$1; RE=$2; CNT=${3:-4}
SNB200_NO_GRAPH=1 timeout -k 10 1200 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:$RE -s ${4:-6} -c $CNT -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out/${TAG}_prof.ncu-rep
