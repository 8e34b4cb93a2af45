#!/bin/bash
# backward-pass session: gradient checks, the tests that touch the backward kernels, the training-step lines of bench_configs, a launch list of the training step
mkdir -p gpurun_out
TAG=${1:-r2bw}
timeout -k 10 300 python tools/check_generator_bwd.py > gpurun_out/${TAG}_check_bwd.txt 2>&1; echo "check bwd rc=$?"; grep -v "^$" gpurun_out/${TAG}_check_bwd.txt | tail -12 | cut -c1-250
timeout -k 10 900 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider -k "backward or graphed or trainers or train or chamfer or registration or samplenet" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python tools/bench_configs.py --only train > gpurun_out/${TAG}_train.jsonl 2> gpurun_out/${TAG}_train.err; cut -c1-330 gpurun_out/${TAG}_train.jsonl
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_launches.csv python tools/profile_train_step.py > gpurun_out/${TAG}_train_ncu.log 2>&1; echo "ncu train rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_train_launches.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=="ID"][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
d=rows[hdr+1:]; n=len(d); last=d[n*3//4:]
agg=collections.OrderedDict()
for r in last:
    v=float(r[vi].replace(",","")); v = v/1000 if r[ui]=="ns" else v
    k=r[ki][:60]
    if "snb::" in k: agg[k]=agg.get(k,[0,0]); agg[k][0]+=v; agg[k][1]+=1
for k,(v,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]): print("%8.1f us  x%-3d %s" % (v,c,k))
PY
