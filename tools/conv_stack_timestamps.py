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
        sb.ops.generator_forward(x, "bnc", conv, fc, True, 64)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 64)()
    sb._lib.lib().snb200_debug_conv_stack_timestamps(ctypes.addressof(buf))
ts = list(buf); t0 = ts[0]
names = {0: "start", 1: "setup done", 2: "moments + barrier done"}
for l in range(4):
    for i, n in enumerate(["layer start", "W staged", "scale/shift", "main loop issued", "acc ready", "epilogue done", "grid barrier done"]):
        names[3 + l * 8 + i] = "L%d %s" % (l + 2, n)
names[36] = "head: start"; names[37] = "head: pooled"; names[38] = "head: barrier P done"
for l in range(4):
    for i, n in enumerate(["start", "input staged", "partials done", "combined", "BN scale/shift", "stored"]):
        names[39 + l * 6 + i] = "FC%d %s" % (l + 1, n)
prev = t0
for i in sorted(names):
    print("%-28s %8d cycles  (+%d)" % (names[i], ts[i] - t0, ts[i] - prev)); prev = ts[i]
