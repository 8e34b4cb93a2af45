#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r2p}
# every launch of 4 eager training steps with its device time
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_launches.csv python tools/profile_train_step.py > gpurun_out/${TAG}_train_ncu.log 2>&1; echo "ncu train rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_train_launches.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=="ID"][0]
h=rows[hdr]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
d=rows[hdr+1:]
n=len(d); last=d[n*3//4:]     # the last of the 4 steps
tot=0; agg=collections.OrderedDict()
for r in last:
    v=float(r[vi].replace(",","")); v = v/1000 if r[ui]=="ns" else v
    k=r[ki][:70]; agg[k]=agg.get(k,[0,0]); agg[k][0]+=v; agg[k][1]+=1; tot+=v
print("last step: %d launches, %.1f us of kernel time" % (len(last), tot))
for k,(v,c) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:30]:
    print("%8.1f us  x%-3d %s" % (v,c,k))
PY
# bench launch list (eager) and the full-set profile of the step kernels
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
SNB200_NO_GRAPH=1 timeout -k 10 900 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:conv_stack_kernel|tail_fused' -s 8 -c 4 -o gpurun_out/${TAG}_prof -f python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 10 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k 'regex:conv_bwd_kernel|pool_bwd|fc_bwd' -s 11 -c 10 -o gpurun_out/${TAG}_prof_bwd -f python tools/profile_train_step.py > gpurun_out/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
ls -la gpurun_out/${TAG}_*
