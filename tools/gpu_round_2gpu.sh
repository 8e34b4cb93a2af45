#!/bin/bash
# Two-GPU session: N=1 and N=2 bench lines (both arms at N=2 are launched exactly like the driver does).
mkdir -p gpurun_out
TAG=${1:-r1}
timeout -k 10 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 rc=$?"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "n2 rc=$?"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_ref_n2.json 2> gpurun_out/${TAG}_bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench_n[12].json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['n_gpus'], round(d['value']), 'clouds/s', round(d['ms_per_step']*1e3,1),'us/step e2e', round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/${TAG}_bench_n2.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_configs.py --only train > gpurun_out/${TAG}_train_n2.jsonl 2> gpurun_out/${TAG}_train_n2.err; echo "train n2 rc=$?"; cat gpurun_out/${TAG}_train_n2.jsonl | cut -c1-300

