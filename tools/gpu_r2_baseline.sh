#!/bin/bash
# Round-2 GPU session A: GPU tests, conv-stack timeline, compute-sanitizer logs.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout -k 10 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout -k 10 120 python tools/conv_stack_timestamps.py > gpurun_out/${TAG}_cs_timeline.txt 2>&1
for tool in memcheck synccheck racecheck initcheck; do
  timeout -k 10 420 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_ops.py > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/${TAG}_sanitizer_${tool}.log
done
tail -15 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_cs_timeline.txt; for t in memcheck synccheck racecheck initcheck; do echo "== $t"; tail -6 gpurun_out/${TAG}_sanitizer_${t}.log; done
