#!/bin/bash
# One GPU-box session: smoke, GPU parity tests, bench (both arms), ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout -k 10 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/${TAG}_smoke.log; tail -40 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
# ncu --set full of the two step kernels (same bench command, eager launches) and the secondary configurations
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:conv_stack|tail_fused' -s 8 -c 4 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 10 600 python tools/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
timeout -k 10 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:approxmatch|matchcost' -s 5 -c 5 -o gpurun_out/${TAG}_emd -f python tools/run_emd.py > gpurun_out/${TAG}_ncu_emd.log 2>&1; echo "ncu emd rc=$?"

