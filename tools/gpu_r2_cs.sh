#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2b}
timeout -k 10 200 python tools/check_conv_stack.py > gpurun_out/${TAG}_check_cs.txt 2>&1; echo "check rc=$?" | tee -a gpurun_out/${TAG}_check_cs.txt
timeout -k 10 200 python tools/check_generator_bwd.py > gpurun_out/${TAG}_check_bwd.txt 2>&1; echo "check bwd rc=$?" | tee -a gpurun_out/${TAG}_check_bwd.txt
cat gpurun_out/${TAG}_check_cs.txt gpurun_out/${TAG}_check_bwd.txt
