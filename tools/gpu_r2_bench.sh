#!/bin/bash
# bench (both arms) + secondary configurations + ncu launch list
mkdir -p gpurun_out
TAG=${1:-r2o}
timeout -k 10 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout -k 10 900 python tools/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","launches_per_step")}, d["e2e"]["value"], d["clocks"])
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","frac_of_3xtf32_ceiling","us_per_launch")}, d["roofline"].get("conv_phase_only"), d["roofline"].get("saturated_B"))
print("kernel_us", d["kernel_us"])
print("cpu", d["cpu_baseline"])
print("gpu_reference", d["gpu_reference"])
PY
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-600; tail -3 gpurun_out/${TAG}_bench.err
cut -c1-400 gpurun_out/${TAG}_configs.jsonl; tail -5 gpurun_out/${TAG}_configs.err
