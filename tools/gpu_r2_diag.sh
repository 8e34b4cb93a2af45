#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2dg}
timeout -k 10 300 python tools/diag_multislice.py > gpurun_out/${TAG}_diag.txt 2>&1; echo "diag rc=$?"; cat gpurun_out/${TAG}_diag.txt | cut -c1-220
timeout -k 10 600 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "range_guard or host_pipeline or conv_stack_kernel_vs" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-200
timeout -k 10 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
