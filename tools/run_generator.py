"""Tiny driver for profilers: runs the generator (and optionally the whole eager step) a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb
torch.manual_seed(0)
net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda().train()
x = torch.rand(32, 1024, 3, device="cuda") - 0.5
conv, fc = net._layer_specs()
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        sb.ops.generator_forward(x, "bnc", conv, fc, True, 64)
        simp, proj = net(x)
        net.get_simplification_loss(x, simp, 64)
    torch.cuda.synchronize()
print("done", sb._lib.launch_count())
