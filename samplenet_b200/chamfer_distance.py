"""ChamferDistance -- drop-in for registration/src/chamfer_distance/chamfer_distance.py:14-66.

Same call convention (`ChamferDistance()(xyz1, xyz2) -> (dist1, dist2)`, BNC float32 tensors, squared distances to the
nearest neighbour in the other cloud, autograd through both inputs), one fused forward launch instead of two and a
deterministic backward instead of float atomics.  CUDA tensors only.
"""
import torch

from . import ops


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        dist1, idx1, dist2, idx2 = ops.nn_distance_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        gradxyz1, gradxyz2 = ops.nn_distance_backward(xyz1, xyz2, graddist1.contiguous(), idx1, graddist2.contiguous(), idx2)
        return gradxyz1, gradxyz2


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)
