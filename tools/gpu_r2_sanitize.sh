#!/bin/bash
# compute-sanitizer over one small invocation of every kernel family (tools/sanitize_ops.py); logs -> gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r2t}
for tool in memcheck synccheck initcheck racecheck; do
  FAMS=""; if [ $tool != racecheck ]; then FAMS="chamfer softproj tail generator emd matching group train progressive multislice"; fi
  timeout -k 10 500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_ops.py $FAMS > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/${TAG}_sanitizer_${tool}.log
  grep -c "ERROR SUMMARY\|=========" gpurun_out/${TAG}_sanitizer_${tool}.log; grep "ERROR SUMMARY\|ok$\|done\|Error\|error" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -14
done
