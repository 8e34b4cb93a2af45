"""Task networks of the classification and reconstruction trainers, restated in stock torch ops.  They are CALLERS
of the hot path -- frozen while the sampler trains (train_samplenet.py:227-232, sampler/train_samplenet.py:100-118) -- so they stay plain
torch modules; only the layer stacks' structure, widths, BatchNorm placement and eps follow the reference:

    PointNetCls   classification/models/pointnet_cls_basic.py:55-136 (vanilla PointNet: 3-64-64-64-128-1024 1x1 convs with BN, max-pool,
                  fc 512 - fc 256 - dropout(keep 0.7) - fc 40; tf_util batch norm eps 1e-3); get_loss = mean sparse softmax cross-entropy
    PointNetAE    reconstruction/src/ae_templates.py:24-37 + encoders_decoders.py (encoder 64-128-128-256-bneck 1x1 convs with BN + ReLU and
                  max symmetry; decoder FC 256-256-n*3, ReLU between, no BN), output reshaped to (B, n, 3)
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class PointNetCls(nn.Module):
    def __init__(self, num_classes=40, bn_eps=1e-3):
        super().__init__()
        w = [3, 64, 64, 64, 128, 1024]
        self.convs = nn.ModuleList([nn.Conv1d(w[i], w[i + 1], 1) for i in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(w[i + 1], eps=bn_eps) for i in range(5)])
        self.fc1, self.bn_fc1 = nn.Linear(1024, 512), nn.BatchNorm1d(512, eps=bn_eps)
        self.fc2, self.bn_fc2 = nn.Linear(512, 256), nn.BatchNorm1d(256, eps=bn_eps)
        self.dp1 = nn.Dropout(p=0.3)
        self.fc3 = nn.Linear(256, num_classes)

    def forward(self, point_cloud):
        """point_cloud (B, N, 3) -> (logits (B, classes), end_points)."""
        y = point_cloud.permute(0, 2, 1)
        for conv, bn in zip(self.convs, self.bns):
            y = F.relu(bn(conv(y)))
        end_points = {"critical_set_idx": torch.argmax(y, dim=2)}
        y = torch.max(y, 2)[0]
        end_points["GFV"] = y
        y = F.relu(self.bn_fc1(self.fc1(y)))
        y = F.relu(self.bn_fc2(self.fc2(y)))
        return self.fc3(self.dp1(y)), end_points

    @staticmethod
    def get_loss(pred, label, end_points=None):
        return F.cross_entropy(pred, label.long())


class PointNetAE(nn.Module):
    def __init__(self, n_pc_points=2048, bneck_size=128, bn_eps=1e-3):
        super().__init__()
        w = [3, 64, 128, 128, 256, bneck_size]
        self.convs = nn.ModuleList([nn.Conv1d(w[i], w[i + 1], 1) for i in range(5)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(w[i + 1], eps=bn_eps) for i in range(5)])
        self.dec = nn.ModuleList([nn.Linear(bneck_size, 256), nn.Linear(256, 256), nn.Linear(256, n_pc_points * 3)])
        self.n_pc_points = n_pc_points

    def encode(self, x):
        y = x.permute(0, 2, 1)
        for conv, bn in zip(self.convs, self.bns):
            y = F.relu(bn(conv(y)))
        return torch.max(y, 2)[0]

    def decode(self, z):
        y = F.relu(self.dec[0](z))
        y = F.relu(self.dec[1](y))
        return self.dec[2](y).view(-1, self.n_pc_points, 3)

    def forward(self, x):
        return self.decode(self.encode(x))
