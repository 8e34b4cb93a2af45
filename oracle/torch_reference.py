"""Plain-torch CPU restatement of the reference's SampleNet step (test infrastructure, NOT product):

    registration/src/samplenet.py:82-161   forward (train mode): conv/BN/ReLU x5 -> max -> FC/BN/ReLU x3 -> FC -> soft projection
    registration/src/samplenet.py:171-181  get_simplification_loss

The layer stack is stock torch (nn.Conv1d / BatchNorm1d / Linear run by torch's CPU kernels, exactly what the reference
executes on CPU); kNN + soft projection go through the C oracle (the reference's knn_cuda / pointnet2 are CUDA-only);
Chamfer goes through the reference's own CPU code in oracle/_ref when present, else the C oracle.
Used by bench.py (cpu_baseline and --impl reference) and by tests.  Only importable from those places.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import oracle as orc


class ReferenceGenerator(nn.Module):
    """Same parameter names / registration order as registration/src/samplenet.py:40-60."""

    def __init__(self, num_out_points, bottleneck_size):
        super().__init__()
        self.num_out_points = num_out_points
        w = [3, 64, 64, 64, 128, bottleneck_size]
        for i in range(5):
            setattr(self, "conv%d" % (i + 1), nn.Conv1d(w[i], w[i + 1], 1))
        for i in range(5):
            setattr(self, "bn%d" % (i + 1), nn.BatchNorm1d(w[i + 1]))
        f = [bottleneck_size, 256, 256, 256, 3 * num_out_points]
        for i in range(4):
            setattr(self, "fc%d" % (i + 1), nn.Linear(f[i], f[i + 1]))
        for i in range(3):
            setattr(self, "bn_fc%d" % (i + 1), nn.BatchNorm1d(256))

    def forward(self, x_bcn):
        y = x_bcn
        for i in range(1, 6):
            y = F.relu(getattr(self, "bn%d" % i)(getattr(self, "conv%d" % i)(y)))
        y = torch.max(y, 2)[0]
        for i in range(1, 4):
            y = F.relu(getattr(self, "bn_fc%d" % i)(getattr(self, "fc%d" % i)(y)))
        y = self.fc4(y)
        return y.view(-1, 3, self.num_out_points)


_POOL = {}


def _pool(workers):
    from concurrent.futures import ThreadPoolExecutor
    if workers not in _POOL:
        _POOL[workers] = ThreadPoolExecutor(max_workers=workers)
    return _POOL[workers]


def _over_clouds(fn, workers, *arrays):
    """Run fn(chunk of every array) over contiguous chunks of the batch on `workers` threads (the C calls release the GIL) and
    concatenate the results -- clouds are independent, so this is how the CPU path uses all host threads."""
    b = arrays[0].shape[0]
    if workers <= 1 or b <= 1:
        return fn(*arrays)
    w = min(workers, b)
    bounds = [(b * i) // w for i in range(w + 1)]
    futs = [_pool(w).submit(fn, *[np.ascontiguousarray(a[bounds[i]:bounds[i + 1]]) for a in arrays]) for i in range(w)]
    parts = [f.result() for f in futs]
    return tuple(np.concatenate([p[j] for p in parts], axis=0) for j in range(len(parts[0])))


def cpu_generator(gen, x_bnc):
    """The layer stack (torch CPU kernels, torch's intra-op threads): x (B,N,3) torch -> simp (B,M,3) numpy."""
    with torch.no_grad():
        return gen(x_bnc.permute(0, 2, 1)).permute(0, 2, 1).contiguous().numpy()


def cpu_pairwise(x_bnc, simp, k, sigma, gamma=1.0, delta=0.0, workers=1):
    """kNN + soft projection (C oracle) and Chamfer (the reference's CPU code when oracle/_ref exists) + the loss reductions,
    cloud-parallel on `workers` threads.  Returns (proj, loss)."""
    x = x_bnc.numpy()
    _, idx = _over_clouds(lambda a, q: orc.knn_point(k, a, q, contract=False, tie_mode=0), workers, x, simp)
    proj, _, _ = _over_clouds(lambda a, q, i: orc.soft_project(a, q, i, float(sigma)), workers, x, simp, idx)
    if orc.have_ref():
        c12, _, c21, _ = _over_clouds(lambda q, a: orc.ref_chamfer_forward(q, a), workers, simp, x)
    else:
        c12, _, c21, _ = _over_clouds(lambda q, a: orc.nn_distance(q, a), workers, simp, x)
    m = simp.shape[1]
    loss = np.float32(c12.mean(dtype=np.float32) + c12.max(axis=1).mean(dtype=np.float32) + np.float32(gamma + delta * m) * c21.mean(dtype=np.float32))
    return proj, loss


def cpu_step(gen, x_bnc, k, sigma, gamma=1.0, delta=0.0, workers=1):
    """One forward (train mode) + simplification loss on the CPU.  x_bnc: torch (B,N,3) float32.  Returns (simp, proj, loss).
    workers > 1: kNN / projection / Chamfer run cloud-parallel on that many threads (the layer stack uses torch's own threads)."""
    simp = cpu_generator(gen, x_bnc)
    proj, loss = cpu_pairwise(x_bnc, simp, k, sigma, gamma, delta, workers)
    return simp, proj, loss
