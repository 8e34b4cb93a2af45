"""Tiny driver for profilers: EMD approx_match / match_cost at the reconstruction AE size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb
b, n = (int(sys.argv[1]) if len(sys.argv) > 1 else 50), 2048
g = torch.Generator().manual_seed(0)
x = torch.rand(b, n, 3, generator=g).cuda(); y = torch.rand(b, n, 3, generator=g).cuda()
for _ in range(2):
    m = sb.ops.approx_match(x, y)
    c = sb.ops.match_cost_forward(x, y, m)
    g1, g2 = sb.ops.match_cost_grad(x, y, m)
torch.cuda.synchronize()
print("done", float(c.sum()))
