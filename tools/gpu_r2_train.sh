#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2s}
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "backward or graphed or progressive or trainers or train" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python tools/bench_configs.py --only progressive > gpurun_out/${TAG}_prog.jsonl 2> gpurun_out/${TAG}_prog.err; cut -c1-520 gpurun_out/${TAG}_prog.jsonl | tail -1
