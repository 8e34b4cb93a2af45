"""SoftProjection -- drop-in for registration/src/soft_projection.py:22-152 (PyTorch flavour, BCN tensors).

The reference chains knn_cuda.KNN, pointnet2 grouping_operation and ~8 small torch kernels; here `project`,
`propagate` and `project_and_propagate` are ONE fused CUDA launch each (brute-force kNN with a warp-resident top-k,
temperature softmax, weighted gather), with a hand-written backward for point_cloud, query_cloud, point_features and the
temperature.  Constructor, attributes (`_temperature`, `_group_size`, `_min_sigma`), `sigma()` and error behaviour follow
the reference.
"""
import torch
import torch.nn as nn

from . import ops


def knn_point(group_size, point_cloud, query_cloud):
    """registration/src/soft_projection.py:11-14: (dist, idx) of knn_cuda.KNN(k, transpose_mode=False) for BCN clouds:
    dist (B, k, M) Euclidean distances ascending, idx (B, k, M) int64."""
    o = ops.knn_soft_project_forward(point_cloud, query_cloud, group_size, "bcn", want=("idx", "val"))
    dist = o["val"].sqrt().permute(0, 2, 1).contiguous()
    idx = o["idx"].long().permute(0, 2, 1).contiguous()
    return dist, idx


class SoftProjection(nn.Module):
    def __init__(self, group_size, initial_temperature=1.0, is_temperature_trainable=True, min_sigma=1e-4):
        """Computes a soft nearest neighbor point cloud (see the reference docstring, soft_projection.py:30-45).

        Inputs:  point_cloud (B, 3, N), query_cloud (B, 3, M), point_features (B, F, N) [optional],
                 action: 'project', 'propagate' or 'project_and_propagate'.
        Outputs: projected_points (B, 3, M) and/or propagated_features (B, F, M).
        """
        super().__init__()
        self._group_size = group_size
        self._temperature = torch.nn.Parameter(
            torch.tensor(initial_temperature, requires_grad=is_temperature_trainable, dtype=torch.float32)
        )
        self._min_sigma = torch.tensor(min_sigma, dtype=torch.float32)
        self._min_sigma_value = None
        self._layout = "bcn"

    def forward(self, point_cloud, query_cloud, point_features=None, action="project"):
        point_cloud = point_cloud.contiguous()
        query_cloud = query_cloud.contiguous()
        if action == "project":
            return self.project(point_cloud, query_cloud)
        elif action == "propagate":
            return self.propagate(point_cloud, point_features, query_cloud)
        elif action == "project_and_propagate":
            return self.project_and_propagate(point_cloud, point_features, query_cloud)
        else:
            raise ValueError("action should be one of the following: 'project', 'propagate', 'project_and_propagate'")

    def sigma(self):
        # max(T^2, min_sigma) as in the reference (soft_projection.py:97-99).  The bound enters as a Python scalar so that
        # no host->device copy happens per call (the reference's `.to(device)` would break CUDA-graph capture).
        if self._min_sigma_value is None:
            self._min_sigma_value = float(self._min_sigma)
        return torch.clamp(self._temperature ** 2, min=self._min_sigma_value)

    def _run(self, point_cloud, query_cloud, point_features, want_proj, want_prop, hard=False, layout=None):
        # the kernel evaluates sigma = max(T^2, min_sigma) itself (no elementwise launches on the step's critical path)
        if self._min_sigma_value is None:
            self._min_sigma_value = float(self._min_sigma)
        return ops.SoftProjectFunction.apply(point_cloud, query_cloud, self._temperature, point_features, self._group_size,
                                             layout or self._layout, hard, want_proj, want_prop, 1, self._min_sigma_value)

    def project_and_propagate(self, point_cloud, point_features, query_cloud):
        proj, prop, _, _, _ = self._run(point_cloud, query_cloud, point_features, True, True)
        return (proj, prop)

    def propagate(self, point_cloud, point_features, query_cloud):
        _, prop, _, _, _ = self._run(point_cloud, query_cloud, point_features, False, True)
        return prop

    def project(self, point_cloud, query_cloud, hard=False, layout=None):
        if hard:
            raise NotImplementedError  # as the reference (soft_projection.py:144-145)
        proj, _, _, _, _ = self._run(point_cloud, query_cloud, None, True, False, layout=layout)
        return proj
