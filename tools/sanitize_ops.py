"""One small invocation of every kernel family through the public API -- the workload for compute-sanitizer
(memcheck / racecheck / synccheck / initcheck):   compute-sanitizer --tool racecheck python tools/sanitize_ops.py [families...]
Sizes are small: the sanitizers slow kernels 10-100x and the persistent kernels spin on grid barriers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import samplenet_b200 as sb
from samplenet_b200 import tf_ops

fam = set(sys.argv[1:]) or {"chamfer", "softproj", "tail", "generator", "emd", "matching", "group", "train", "progressive"}
torch.manual_seed(0)
dev = torch.device("cuda:0")
x = (torch.rand(4, 256, 3, device=dev) - 0.5)
q = (x[:, :32] + 0.02 * torch.randn(4, 32, 3, device=dev)).contiguous()
if "chamfer" in fam:
    a, b = q.clone().requires_grad_(True), x.clone().requires_grad_(True)
    d1, d2 = sb.ChamferDistance()(a, b)
    (d1.mean() + d2.mean()).backward()
    tf_ops.nn_distance(q, x)
    print("chamfer ok")
if "softproj" in fam:
    sp = sb.SoftProjection(8, 1.0).to(dev)
    pc, qc = x.permute(0, 2, 1).contiguous().requires_grad_(True), q.permute(0, 2, 1).contiguous().requires_grad_(True)
    feats = torch.rand(4, 5, 256, device=dev, requires_grad=True)
    pr, prop = sp(pc, qc, feats, action="project_and_propagate")
    (pr.sum() + prop.sum()).backward()
    tf_ops.knn_point(7, x, q)
    print("softproj ok")
if "group" in fam:
    _, idx = tf_ops.knn_point(4, x, q)
    pts = x.clone().requires_grad_(True)
    tf_ops.group_point(pts, idx).sum().backward()
    print("group ok")
if "generator" in fam or "tail" in fam or "train" in fam:
    net = sb.SampleNet(32, 128, group_size=8, input_shape="bnc", output_shape="bnc").to(dev).train()
    if "generator" in fam:
        with torch.no_grad():
            conv, fc = net._layer_specs()
            sb.ops.generator_forward(x, "bnc", conv, fc, True, 32)
            sb.ops.generator_forward(x, "bnc", conv, fc, True, 32, per_layer_kernels=True)
            sb.ops.generator_forward(x, "bnc", conv, fc, True, 32, exact_fp32=True)
            net.eval(); net(x); net.train()
        print("generator ok")
    if "tail" in fam:
        with torch.no_grad():
            simp, proj = net(x)
            net.get_simplification_loss(x, simp, 32)
        print("tail ok")
    if "train" in fam:
        simp, proj = net(x)
        loss = net.get_simplification_loss(x, simp, 32) + 0.01 * net.get_projection_loss() + proj.sum() * 0.0
        loss.backward()
        print("train ok")
if "emd" in fam:
    a = torch.rand(2, 96, 3, device=dev).requires_grad_(True)
    b = torch.rand(2, 64, 3, device=dev).requires_grad_(True)
    match = tf_ops.approx_match(a, b)
    tf_ops.match_cost(a, b, match).sum().backward()
    print("emd ok")
if "progressive" in fam:
    from samplenet_b200 import trainers
    so = torch.rand(4, 64, 3, device=dev, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    trainers.progressive_simplification_loss(xr, so, [2, 4, 8, 16, 32, 64]).backward()
    print("progressive ok")
if "emd" in fam:
    tf_ops.approx_match(torch.rand(2, 48, 3, device=dev), torch.rand(2, 32, 3, device=dev), exact=True)
    print("emd exact ok")
if "multislice" in fam:   # more than 256 points per SM: every CTA of the persistent generator kernel walks two slices per layer (not in the default set:
    xb = torch.rand(40, 1024, 3, device=dev) - 0.5   # 148 CTAs x 512 threads under racecheck take minutes)
    netb = sb.SampleNet(32, 128, group_size=8, input_shape="bnc", output_shape="bnc").to(dev).train()
    simp, proj = netb(xb)
    (netb.get_simplification_loss(xb, simp, 32) + proj.sum() * 0.0).backward()
    print("multislice ok")
if "matching" in fam:
    _, idx1, _, _ = sb.ops.nn_distance_forward(q, x)
    sb.sputils.nn_matching_cuda(x, idx1, 32)
    print("matching ok")
torch.cuda.synchronize()
print("sanitize_ops done")
