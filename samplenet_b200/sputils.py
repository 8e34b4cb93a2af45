"""sputils -- drop-in for registration/src/sputils.py: the shared argparse flags (verbatim-compatible) and nn_matching.

`nn_matching` keeps the reference's numpy signature (full_pc (B,N,3), idx (B,M), k -> (B,k,3) float64 array) but runs
the order-preserving unique + farthest-point completion on the GPU (`ops.nn_matching`); `nn_matching_cuda` is the
tensor-in / tensor-out form used by SampleNet.forward in eval mode (no host round trip).
"""
import argparse

import numpy as np
import torch

from . import ops


def nn_matching_cuda(full_pc, idx, k, complete_fps=True):
    """full_pc (B,N,3) float32 CUDA, idx (B,M) int -> matched (B,k,3) float32 CUDA."""
    return ops.nn_matching(full_pc, idx.to(torch.int32).contiguous(), k, complete_fps)


def nn_matching(full_pc, idx, k, complete_fps=True):
    """numpy in / numpy out like registration/src/sputils.py:31-41; computed by the CUDA kernel."""
    dev = torch.device("cuda", torch.cuda.current_device())
    pc = torch.as_tensor(np.ascontiguousarray(full_pc, dtype=np.float32), device=dev)
    ii = torch.as_tensor(np.ascontiguousarray(idx).astype(np.int32), device=dev)
    out = nn_matching_cuda(pc, ii, k, complete_fps)
    return out.cpu().numpy().astype(np.float64)


def simple_projection_and_continued_fps(full_pc, gen_pc, idx):
    """reconstruction/src/samplenet_pointnet_ae.py:535-549 on the GPU: per cloud, order-preserving unique of the nearest-neighbour indices
    `idx` (B, k) of the generated points, completed to k points by farthest point sampling seeded with them
    (`fps_from_given_indices`, :513-533).  full_pc (B,N,3) float32 CUDA, gen_pc (B,k,3) (only its k is used, as in the reference),
    idx (B,k) int -> (out_pc (B,k,3), out_pc_idx (B,k) int32, n_unique_points (B,1))."""
    k = gen_pc.shape[1]
    ii = idx.to(torch.int32).contiguous()
    out_pc, out_idx = ops.nn_matching(full_pc, ii, k, complete_fps=True, return_idx=True)
    srt = torch.sort(ii.long(), dim=1)[0]
    n_unique = 1 + (srt[:, 1:] != srt[:, :-1]).sum(dim=1, keepdim=True)
    return out_pc, out_idx, n_unique


# Flag table: (short/long names, type or action, default, help).  Produces exactly the parser of
# registration/src/sputils.py:45-62 so that scripts written against the reference parse the same command lines.
_FLAGS = (
    (("--skip-projection",), "store_true", None, "Do not project points in training"),
    (("-in", "--num-in-points"), int, 1024, "Number of input Points [default: 1024]"),
    (("-out", "--num-out-points"), int, 64, "Number of output points [2, 1024] [default: 64]"),
    (("--bottleneck-size",), int, 128, "bottleneck size [default: 128]"),
    (("--alpha",), float, 0.01, "Simplification regularization loss weight [default: 0.01]"),
    (("--gamma",), float, 1, "Lb constant regularization loss weight [default: 1]"),
    (("--delta",), float, 0, "Lb linear regularization loss weight [default: 0]"),
    (("-gs", "--projection-group-size"), int, 8, "Neighborhood size in Soft Projection [default: 8]"),
    (("--lmbda",), float, 0.01, "Projection regularization loss weight [default: 0.01]"),
)


def get_parser():
    parser = argparse.ArgumentParser("SampleNet: Differentiable Point Cloud Sampling")
    for names, kind, default, text in _FLAGS:
        if kind == "store_true":
            parser.add_argument(*names, action="store_true", help=text)
        else:
            parser.add_argument(*names, type=kind, default=default, help=text)
    return parser
