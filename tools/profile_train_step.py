"""Eager training steps (forward + losses + CUDA backward + Adam) for the profilers: `ncu --metrics gpu__time_duration.sum ... python tools/profile_train_step.py`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb
torch.manual_seed(0)
net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
x = torch.rand(32, 1024, 3, device="cuda") - 0.5
for it in range(int(os.environ.get("STEPS", "4"))):
    torch.cuda.nvtx.range_push("step%d" % it)
    opt.zero_grad(set_to_none=False)
    simp, proj = net(x)
    loss = 0.01 * net.get_simplification_loss(x, simp, 64) + 0.01 * net.get_projection_loss() + proj.sum() * 0.0
    loss.backward()
    opt.step()
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done", float(loss))
