#!/bin/bash
# multi-slice conv stack: correctness on all shapes, timelines at B=32 / 128, quick tests, bench
mkdir -p gpurun_out
TAG=${1:-r2mu}
timeout -k 10 300 python tools/check_conv_stack.py > gpurun_out/${TAG}_check_cs.txt 2>&1; echo "check rc=$?"
grep "MISMATCH\|ALL OK\|FAILED\|Error\|error" gpurun_out/${TAG}_check_cs.txt | head
CS_QUICK=1 CS_TL_B=128 timeout -k 10 200 python tools/check_conv_stack.py > gpurun_out/${TAG}_check_cs_b128.txt 2>&1
grep "head: start\|FC. input staged\|FC4 stored\|generator" gpurun_out/${TAG}_check_cs_b128.txt
timeout -k 10 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "samplenet or generator or conv_stack or graph or tf_variant or backward" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["clocks"])
print({k:d["kernel_us"][k] for k in ("generator_us","conv_stack_us","tail_fused_us","sat_generator_us")}, d["roofline"]["saturated_B"])
PY
