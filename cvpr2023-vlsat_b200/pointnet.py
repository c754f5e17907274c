"""Host-side mirror of ``PointNetfeat`` and ``PointNetRelClsMulti``
(src/model/model_utils/network_PointNet.py:88-176, 305-341): same constructor arguments, parameter
names (``conv1..3``, ``fc1..3``) and xavier-normal / zero-bias initialisation; forward on the fused
vlsat_b200 kernels."""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._cache import DerivedCache, require_inference


def _init_xavier_normal(module: nn.Module) -> None:
    """``BaseNetwork.init_weights('xavier_normal', 1)`` (networks_base.py:9-46): every sub-module with a
    weight gets xavier_normal_(gain=1), every bias 0; BatchNorm weights were set to 1 beforehand."""
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Linear)):
            nn.init.xavier_normal_(m.weight.data, gain=1)
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0.0)
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight.data, 1.0)   # the later xavier pass fails on 1-D weights in torch>=1.x
            nn.init.constant_(m.bias.data, 0.0)


class PointNetfeat(nn.Module):
    def __init__(self, global_feat=True, input_transform=True, feature_transform=False, point_size=3, out_size=1024,
                 batch_norm=True, init_weights=True, pointnet_str: str = None):
        super().__init__()
        if input_transform or feature_transform:
            raise NotImplementedError("STN input/feature transforms are disabled on the VL-SAT path "
                                      "(SGFN_MMG/model.py:51-57; feature_transform=false in mmgnet.json)")
        if not global_feat:
            raise NotImplementedError("global_feat=False (per-point features) is not on the VL-SAT path")
        self.name = 'pnetenc'
        self.use_batch_norm = batch_norm
        self.point_size, self.out_size = point_size, out_size
        self.global_feat, self.input_transform, self.feature_transform = True, False, False
        self.relu = nn.ReLU()
        self.conv1 = nn.Conv1d(point_size, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, out_size, 1)
        if batch_norm:
            # network_PointNet.py:142-143,155-156,161-162 call bnX(x) and DISCARD the result, so the
            # output does not depend on these modules; they exist only for state_dict compatibility.
            self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(out_size)
        if init_weights:
            _init_xavier_normal(self)
        self._cache = DerivedCache()

    def _weights(self):
        convs = (self.conv1, self.conv2, self.conv3)
        srcs = tuple(p for c in convs for p in (c.weight, c.bias))
        return self._cache.get("w", srcs, lambda: tuple(t for c in convs for t in (c.weight.squeeze(-1).contiguous(),
                                                                                    c.bias.contiguous())))

    def forward(self, x, return_meta=False):
        assert x.ndim > 2
        from . import train_path as T
        if T.differentiable(self):
            out = T.pointnet_feat(self, x)
            return (out, torch.zeros([1]), torch.zeros([1])) if return_meta else out
        w1, b1, w2, b2, w3, b3 = self._weights()
        if x.shape[2] == 1:
            # one "point" per row (the relationship encoders, SGFN_MMG/model.py:305-306): the max is the
            # identity, so the encoder is a 3-layer row MLP on the dense-projection kernels.
            out = ops.linear_chain(x.reshape(x.shape[0], x.shape[1]).contiguous(),
                                   [(w1, b1, ops.ACT_RELU), (w2, b2, ops.ACT_RELU), (w3, b3, ops.ACT_RELU)])
        else:
            out = ops.pointnet(x.contiguous(), w1, b1, w2, b2, w3, b3)
        if return_meta:
            return out, torch.zeros([1]), torch.zeros([1])
        return out


class PointNetRelClsMulti(nn.Module):
    def __init__(self, k=2, in_size=1024, batch_norm=True, drop_out=True, init_weights=True):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("PointNetRelClsMulti(batch_norm=True) is not on the VL-SAT path (WITH_BN=false)")
        self.name = 'pnetcls'
        self.in_size, self.use_bn, self.use_drop_out = in_size, False, drop_out
        self.fc1 = nn.Linear(in_size, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        if drop_out:
            self.dropout = nn.Dropout(p=0.3)
        self.relu = nn.ReLU()
        if init_weights:
            _init_xavier_normal(self)

    def forward(self, x, x_split=None):
        """``x_split``: (hi, lo) pair of x when its producer emitted one (inference path)."""
        from . import train_path as T
        if T.differentiable(self):
            return T.rel_classifier(self, x)
        return ops.linear_chain(x, x_split=x_split, layers=[(self.fc1.weight.detach(), self.fc1.bias.detach(), ops.ACT_RELU),
                                    (self.fc2.weight.detach(), self.fc2.bias.detach(), ops.ACT_RELU),
                                    (self.fc3.weight.detach(), self.fc3.bias.detach(), ops.ACT_SIGMOID)])
