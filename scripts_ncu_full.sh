#!/bin/bash
# ncu --set full capture of selected kernels from the bench step (one GPU).  usage: scripts_ncu_full.sh <tag> <kernel-regex> [count]
mkdir -p gpurun_out
TAG=$1; RE=$2; CNT=${3:-4}
timeout -k 10 1200 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:$RE -s 40 -c $CNT -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out/${TAG}_prof.ncu-rep
