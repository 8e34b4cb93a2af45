"""Which generator path disagrees at a multi-slice batch?  Prints max |diff| of the pooled feature between the default launch (fused head),
the stand-alone conv stack + cluster head, the per-layer kernels and the exact-fp32 path, training and eval."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb

for (b, n) in ((64, 1024), (128, 1024), (41, 1999)):
    torch.manual_seed(b * 1000 + n)
    net = sb.SampleNet(64, 128, group_size=8, input_shape="bnc", output_shape="bnc").cuda()
    with torch.no_grad():
        for bn in [net.bn1, net.bn2, net.bn3, net.bn4, net.bn5]:
            bn.weight.copy_(1 + 0.3 * torch.randn_like(bn.weight)); bn.bias.copy_(0.2 * torch.randn_like(bn.bias))
            bn.running_mean.copy_(0.1 * torch.randn_like(bn.running_mean)); bn.running_var.copy_(0.5 + torch.rand_like(bn.running_var))
    print("negative BN scales:", [int((bn.weight < 0).sum()) for bn in [net.bn1, net.bn2, net.bn3, net.bn4, net.bn5]])
    x = torch.rand(b, n, 3, device="cuda") - 0.5
    conv, fc = net._layer_specs()
    for training in (True, False):
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        res = {}
        for name, kw in (("fused", dict()), ("fused again", dict()), ("separate_head", dict(separate_head=True)), ("per_layer", dict(per_layer_kernels=True)), ("exact", dict(exact_fp32=True))):
            net.load_state_dict(sd)
            out, feat = sb.ops.generator_forward(x, "bnc", conv, fc, training, 64, **kw)
            res[name] = (out.clone(), feat.clone())
        ref = res["exact"][1]
        for name in res:
            d = (res[name][1] - ref).abs()
            bad = (d > 3e-4 * ref.abs() + 3e-5)
            idx = bad.nonzero()
            print("b=%d n=%d train=%d  %-14s feat max|diff vs exact| %.3e  bad %d  clouds %s channels %s" %
                  (b, n, training, name, d.max().item(), int(bad.sum()), sorted(set(idx[:, 0].tolist()))[:8], sorted(set(idx[:, 1].tolist()))[:8]), flush=True)
