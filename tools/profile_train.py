"""Kernel-time breakdown of one eager training step (torch profiler, CUDA activities)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
x = torch.rand(32, 1024, 3, device="cuda") - 0.5
def step():
    opt.zero_grad()
    simp, proj = net(x)
    loss = net.get_simplification_loss(x, simp, 64) + net.get_projection_loss() + (proj * proj).mean()
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
