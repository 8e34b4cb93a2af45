#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2k}
export SNB200_CONV_STACK=v2 SNB200_GENERATOR_BACKWARD=cuda
timeout -k 10 200 python tools/diag_bwd_layers.py 64 512 64 bnc > gpurun_out/${TAG}_diag_layers.txt 2>&1; echo "diag rc=$?"
timeout -k 10 200 python tools/diag_bwd_layers.py 32 1024 64 bnc >> gpurun_out/${TAG}_diag_layers.txt 2>&1; echo "diag rc=$?"
cat gpurun_out/${TAG}_diag_layers.txt | cut -c1-200
