import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb
torch.manual_seed(0)
net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda().train()
x = torch.rand(32, 1024, 3, device="cuda") - 0.5
conv, fc = net._layer_specs()
with torch.no_grad():
    for _ in range(5):
        sb.ops.generator_forward(x, "bnc", conv, fc, True, 64, separate_head=True)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    sb._lib.lib().snb200_debug_head_timestamps(ctypes.addressof(buf))
ts = list(buf)
t0 = ts[0]
names = {0: "start", 1: "tma issued", 2: "phase0 done", 3: "sync0 done"}
for l in range(4):
    for i, n in enumerate(["pass start", "before sync A", "after sync A", "input staged", "compute done", "after sync B", "stored", "cluster.sync done"]):
        names[4 + l * 8 + i] = "L%d %s" % (l, n)
prev = t0
for i in range(36):
    print("%-28s %8d cycles  (+%d)" % (names.get(i, str(i)), ts[i] - t0, ts[i] - prev))
    prev = ts[i]
