#!/bin/bash
# Round-2 evidence session on one B200 box: smoke, full GPU test suite, bench (both arms), conv-stack timeline, secondary configurations,
# ncu launch lists (bench step and training step) and ncu --set full captures (step kernels, backward kernels).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r2fin}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout -k 10 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 200 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout -k 10 200 python tools/check_conv_stack.py > gpurun_out/${TAG}_check_cs.txt 2>&1; echo "check rc=$?"
timeout -k 10 900 python tools/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; echo "configs rc=$?"
# launch lists (device time of every launch; cold caches, serialised: shares only)
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_launches.csv python tools/profile_train_step.py > gpurun_out/${TAG}_train_ncu.log 2>&1; echo "ncu train rc=$?"
# full-set captures
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:conv_stack_kernel|tail_fused' -s 8 -c 4 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 10 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:conv_bwd_kernel|conv1_bwd|pool_bwd|fc_bwd|reduce_partials|progressive' -s 11 -c 12 -o gpurun_out/${TAG}_prof_bwd -f python tools/profile_train_step.py > gpurun_out/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
tail -3 gpurun_out/${TAG}_smoke.log; tail -15 gpurun_out/${TAG}_pytest.log; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
ls -la gpurun_out/${TAG}_*
