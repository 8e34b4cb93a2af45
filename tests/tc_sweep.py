"""Bring-up sweep for the tcgen05 layer kernel: tries descriptor encodings and reports the error of each against fp64.
Run on the GPU box:  python tests/tc_sweep.py > gpurun_out/tc_sweep.log"""
import itertools
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import samplenet_b200 as sb

torch.manual_seed(0)
rows, c_in, c_out = 128, 64, 64
A = torch.randn(rows, c_in, device="cuda"); W = torch.randn(c_out, c_in, device="cuda") / 8; bias = torch.zeros(c_out, device="cuda")
ref = A.double() @ W.double().T
def hi(sbo16, version, layout):
    return (sbo16 & 0x3fff) | (version << 14) | (layout << 29)
variants = []
for layout, swz in ((2, 1), (0, 0), (2, 0), (6, 1), (4, 1)):
    for sbo in (64, 8, 128, 1, 32):
        for ver in (1, 0):
            for kadv in (2, 1, 4):
                variants.append((hi(sbo, ver, layout), kadv, swz, "layout=%d swz=%d sbo16=%d ver=%d kadv16=%d" % (layout, swz, sbo, ver, kadv)))
best = []
for dh, kadv, swz, name in variants:
    try:
        D = sb.ops.debug_tc_gemm(A, W, bias, dh, kadv, swz)
        torch.cuda.synchronize()
        err = (D.double() - ref).abs().max().item()
    except Exception as e:  # noqa
        err = float("nan"); name += " EXC " + str(e)[:80]
    best.append((err, name))
    print("%-60s err=%.3e" % (name, err), flush=True)
best = [b for b in best if b[0] == b[0]]
best.sort()
print("BEST:", best[:5])
