"""Host-side mirror of ``Mmgnet`` (src/model/SGFN_MMG/model.py:20-335, forward only) and of the CLIP
``AdapterModel`` (clip_adapter/model.py:6-32): same sub-module names (hence the same ``state_dict``
keys and shapes), same config keys, same ``forward`` signature and outputs.

Out of scope here (callers of the path): losses, optimiser, CLIP text supervision,
checkpoint directories, metrics. ``accelerate_reference_model`` swaps the kernels in underneath an
instance of the reference's own ``Mmgnet`` so ``main.py`` keeps all of that unchanged.
"""
from __future__ import annotations

import json
import math
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import ops
from ._cache import DerivedCache, require_inference
from .mmg import MMG
from .pointnet import PointNetfeat, PointNetRelClsMulti


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    try:
        return getattr(cfg, key)
    except (AttributeError, RuntimeError):       # the reference Config raises RuntimeError on a missing key
        return default


def _need(cfg, key):
    v = _get(cfg, key, None)
    if v is None:
        raise RuntimeError('key', key, 'is not defined!')      # same error type as config.py:48-51
    return v


def load_model_config(path: str) -> dict:
    """Read an mmgnet.json-style file and enforce the ``_KEY`` enumerations (config.py:34-44)."""
    with open(path) as f:
        cfg = json.load(f)

    def check(d):
        for k, v in d.items():
            if isinstance(v, dict):
                check(v)
            elif "_" + k in d and v not in d["_" + k]:
                raise RuntimeError('value for', k, 'should be one of', d["_" + k], 'got', v)
    check(cfg)
    return cfg


DEFAULT_MODEL_CONFIG = dict(N_LAYERS=2, USE_SPATIAL=True, WITH_BN=False, USE_RGB=False, USE_NORMAL=False,
                            USE_GCN_EDGE=True, ATTENTION="fat", DROP_OUT_ATTEN=0.5, multi_rel_outputs=True,
                            feature_transform=False, clip_feat_dim=512, DIM_ATTEN=256, GCN_AGGR="max", NUM_HEADS=8)


class AdapterModel(nn.Module):
    def __init__(self, input_size=512, output_size=512, alpha=0.5):
        super().__init__()
        self.input_size, self.output_size, self.alpha = input_size, output_size, alpha
        self.obj_logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.fc1 = nn.Linear(input_size, 256)
        self.relu = nn.ReLU()
        self.fc2 = nn.Linear(256, output_size)
        for m in (self.fc1, self.fc2):
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        """alpha * fc2(relu(fc1(x))) + (1 - alpha) * x, the residual folded into the GEMM epilogue."""
        return ops.linear_chain(x, [(self.fc1.weight.detach(), self.fc1.bias.detach(), ops.ACT_RELU),
                                    (self.fc2.weight.detach(), self.fc2.bias.detach(), ops.ACT_NONE)],
                                residual=x, alpha=self.alpha, beta=1.0 - self.alpha)


class Mmgnet(nn.Module):
    def __init__(self, config, num_obj_class, num_rel_class, dim_descriptor=11,
                 obj_text_features: Optional[torch.Tensor] = None, adapter_path: Optional[str] = None):
        super().__init__()
        self.config = config
        self.mconfig = mconfig = _need(config, "MODEL")
        with_bn = _need(mconfig, "WITH_BN")
        if with_bn:
            raise NotImplementedError("MODEL.WITH_BN=true: the reference discards the PointNet BN outputs and "
                                      "BatchNorms the GAT MLPs; not built (mmgnet.json ships WITH_BN=false)")
        if _need(mconfig, "feature_transform"):
            raise NotImplementedError("MODEL.feature_transform=true is not supported")
        if not _need(mconfig, "multi_rel_outputs"):
            raise NotImplementedError("MODEL.multi_rel_outputs=false (PointNetRelCls/log_softmax head) is not built")
        dim_point = 3 + (3 if _need(mconfig, "USE_RGB") else 0) + (3 if _need(mconfig, "USE_NORMAL") else 0)
        self.dim_point, self.dim_edge = dim_point, dim_descriptor
        self.num_class, self.num_rel = num_obj_class, num_rel_class
        self.flow = 'target_to_source'
        self.clip_feat_dim = _need(mconfig, "clip_feat_dim")
        dim_point_feature = 768

        self.obj_encoder = PointNetfeat(global_feat=True, batch_norm=with_bn, point_size=dim_point,
                                        input_transform=False, feature_transform=False, out_size=dim_point_feature)
        self.rel_encoder_2d = PointNetfeat(global_feat=True, batch_norm=with_bn, point_size=dim_descriptor,
                                           input_transform=False, feature_transform=False, out_size=512)
        self.rel_encoder_3d = PointNetfeat(global_feat=True, batch_norm=with_bn, point_size=dim_descriptor,
                                           input_transform=False, feature_transform=False, out_size=512)
        self.mmg = MMG(dim_node=512, dim_edge=512, dim_atten=_need(mconfig, "DIM_ATTEN"),
                       depth=_need(mconfig, "N_LAYERS"), num_heads=_need(mconfig, "NUM_HEADS"),
                       aggr=_need(mconfig, "GCN_AGGR"), flow=self.flow, attention=_need(mconfig, "ATTENTION"),
                       use_edge=_need(mconfig, "USE_GCN_EDGE"), DROP_OUT_ATTEN=_need(mconfig, "DROP_OUT_ATTEN"))
        self.triplet_projector_3d = nn.Sequential(nn.Linear(512 * 3, 512 * 2), nn.Dropout(0.5), nn.ReLU(),
                                                  nn.Linear(512 * 2, 512))
        self.triplet_projector_2d = nn.Sequential(nn.Linear(512 * 3, 512 * 2), nn.Dropout(0.5), nn.ReLU(),
                                                  nn.Linear(512 * 2, 512))
        self.clip_adapter = AdapterModel(input_size=512, output_size=512, alpha=0.5)
        self.obj_logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.mlp_3d = nn.Sequential(nn.Linear(512 + 256, 512 - 8), nn.BatchNorm1d(512 - 8), nn.ReLU(), nn.Dropout(0.1))
        self.rel_predictor_3d = PointNetRelClsMulti(num_rel_class, in_size=512, batch_norm=with_bn, drop_out=True)
        self.rel_predictor_2d = PointNetRelClsMulti(num_rel_class, in_size=512, batch_norm=with_bn, drop_out=True)

        # init_weight (SGFN_MMG/model.py:161-187)
        nn.init.xavier_uniform_(self.mlp_3d[0].weight)
        for proj in (self.triplet_projector_3d, self.triplet_projector_2d):
            nn.init.xavier_uniform_(proj[0].weight)
            nn.init.xavier_uniform_(proj[-1].weight)
        self.obj_predictor_2d = nn.Linear(self.clip_feat_dim, num_obj_class)
        self.obj_predictor_3d = nn.Linear(self.clip_feat_dim, num_obj_class)
        if obj_text_features is None:
            # the reference copies L2-normalised CLIP text features here; CLIP weights are not available
            # offline, so default to seeded unit-norm rows (load a checkpoint / pass obj_text_features).
            gen = torch.Generator().manual_seed(1234)
            obj_text_features = torch.randn(num_obj_class, self.clip_feat_dim, generator=gen)
            obj_text_features = obj_text_features / obj_text_features.norm(dim=-1, keepdim=True)
        with torch.no_grad():
            self.obj_predictor_2d.weight.copy_(obj_text_features)
            self.obj_predictor_3d.weight.copy_(obj_text_features)
        if adapter_path is not None:
            self.clip_adapter.load_state_dict(torch.load(adapter_path, 'cpu'))
        for p in self.clip_adapter.parameters():
            p.requires_grad = False
        self._cache = DerivedCache()

    # ---- derived weights -------------------------------------------------------------------------
    def _mlp3d_folded(self):
        """Linear(768,504) with the eval-mode BatchNorm1d(504) folded in (SGFN_MMG/model.py:106-111)."""
        lin, bn = self.mlp_3d[0], self.mlp_3d[1]

        def build():
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            return (lin.weight * s[:, None]).contiguous(), ((lin.bias - bn.running_mean) * s + bn.bias).contiguous()
        return self._cache.get("mlp3d", (lin.weight, lin.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var), build)

    def _pair_weights(self):
        """triplet_projector_2d[0] split for ``cat[o[src], o[dst], e]`` (SGFN_MMG/model.py:260-265)."""
        lin = self.triplet_projector_2d[0]
        return self._cache.get("pair", (lin.weight,),
                               lambda: torch.cat([lin.weight[:, :512], lin.weight[:, 512:1024]], 0).contiguous())

    # ---- forward -----------------------------------------------------------------------------------
    def forward(self, obj_points, obj_2d_feats, edge_indices, descriptor=None, batch_ids=None, istrain=False):
        from . import train_path as T
        if not obj_points.is_cuda:
            raise RuntimeError(f"Mmgnet.forward: expected CUDA tensors (vlsat_b200 has no CPU path), got obj_points on {obj_points.device}")
        if T.differentiable(self):
            return T.mmgnet_forward(self, obj_points, obj_2d_feats, edge_indices, descriptor, batch_ids, istrain,
                                    use_spatial=bool(_need(self.mconfig, "USE_SPATIAL")))
        n = obj_points.shape[0]
        edge_indices = edge_indices.contiguous()
        obj_center = descriptor[:, :3].contiguous()

        def bookkeeping():
            # batch-only work without tensor-memory kernels (scene ranges, distance-bias table, CSR build, edge descriptor):
            # on the side stream next to the PointNet encoder, whose CTAs hold all of TMEM but leave registers / smem over
            self.mmg.prebuild(edge_indices, batch_ids, obj_center, n)
            return ops.edge_descriptor(descriptor.contiguous(), edge_indices)           # [E, 11]
        edge_feature, obj_feature = ops.fork_join(bookkeeping, lambda: self.obj_encoder(obj_points), obj_points.device)   # [N, 768]
        obj_feature_3d_mimic = obj_feature[..., :512].clone() if istrain else None

        w, b = self._mlp3d_folded()
        if _need(self.mconfig, "USE_SPATIAL"):
            node3d = torch.empty((n, w.shape[0] + 8), device=obj_points.device, dtype=torch.float32)
            ops.linear(obj_feature, w, b, act=ops.ACT_RELU, out=node3d[:, :w.shape[0]])
            ops.spatial_tail(descriptor.contiguous(), node3d, w.shape[0])
        else:
            node3d = ops.linear(obj_feature, w, b, act=ops.ACT_RELU)

        ef = edge_feature.unsqueeze(-1)
        # the two relationship encoders read the same input and are independent: two streams (ops.fork_join)
        rel_feature_2d, rel_feature_3d = ops.fork_join(lambda: self.rel_encoder_2d(ef), lambda: self.rel_encoder_3d(ef), ef.device)

        obj_2d = self.clip_adapter(obj_2d_feats.contiguous())
        obj_features_2d_mimic = obj_2d.clone() if istrain else None

        g3, g2, ge3, ge2 = self.mmg(node3d, obj_2d, rel_feature_3d, rel_feature_2d, edge_indices, batch_ids,
                                    obj_center, descriptor, istrain=istrain)

        gcn_edge_feature_2d_dis = None
        if istrain:
            # project per node, gather per edge: W[:, :512] o[src] + W[:, 512:1024] o[dst] + W[:, 1024:] e + b.
            # (The reference also evaluates this in eval mode and throws the result away, model.py:319-322.)
            p0, p3 = self.triplet_projector_2d[0], self.triplet_projector_2d[3]
            ab = ops.linear(g2, self._pair_weights(), None)                          # [N, 2048]
            hid = p0.weight.shape[0]
            h = ops.linear(ge2, p0.weight.detach()[:, 1024:], p0.bias.detach(), act=ops.ACT_RELU,
                           gather=(ab[:, :hid], edge_indices[0].contiguous(), ab[:, hid:], edge_indices[1].contiguous()))
            gcn_edge_feature_2d_dis = ops.linear(h, p3.weight.detach(), p3.bias.detach())

        ge3p, ge2p = getattr(self.mmg, "last_edge_pairs", (None, None))
        rel_cls_2d, rel_cls_3d = ops.fork_join(lambda: self.rel_predictor_2d(ge2, x_split=ge2p),
                                               lambda: self.rel_predictor_3d(ge3, x_split=ge3p), ge3.device)
        self.mmg.last_edge_pairs = (None, None)

        scale = self.obj_logit_scale.detach().reshape(1)
        obj_logits_2d, obj_logits_3d = ops.fork_join(
            lambda: ops.linear(ops.row_l2norm(g2), self.obj_predictor_2d.weight.detach(), self.obj_predictor_2d.bias.detach(), scale_ptr=scale),
            lambda: ops.linear(ops.row_l2norm(g3), self.obj_predictor_3d.weight.detach(), self.obj_predictor_3d.bias.detach(), scale_ptr=scale),
            g3.device)
        if not torch.cuda.is_current_stream_capturing():
            from .attention import validate_inputs
            validate_inputs(obj_points.device)           # unsorted batch_ids raise here, after the last launch (one host sync)
        if istrain:
            return (obj_logits_3d, obj_logits_2d, rel_cls_3d, rel_cls_2d, obj_feature_3d_mimic, obj_features_2d_mimic,
                    gcn_edge_feature_2d_dis, self.obj_logit_scale.detach().exp())
        return obj_logits_3d, obj_logits_2d, rel_cls_3d, rel_cls_2d


def adopt_parameters(dst: nn.Module, src: nn.Module) -> None:
    """Make ``dst`` use the very Parameter/buffer objects of ``src`` (same names by construction), so an
    optimiser or checkpoint loader holding ``src``'s tensors keeps driving ``dst``."""
    src_p, src_b = dict(src.named_parameters()), dict(src.named_buffers())
    for name, _ in list(dst.named_parameters()):
        if name not in src_p:
            raise KeyError(f"parameter {name} missing in the source module")
        mod, leaf = _owner(dst, name)
        mod._parameters[leaf] = src_p[name]
    for name, _ in list(dst.named_buffers()):
        if name in src_b:
            mod, leaf = _owner(dst, name)
            mod._buffers[leaf] = src_b[name]


def _owner(root: nn.Module, dotted: str):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        m = getattr(m, p)
    return m, parts[-1]


def accelerate_reference_model(ref_model: nn.Module) -> nn.Module:
    """Drop-in for ``main.py``: given an instance of the reference's ``Mmgnet`` (already on the GPU), run
    its ``forward`` on the vlsat_b200 kernels while sharing its parameters (optimiser, checkpointing and
    the loss/metric code in ``process_train`` / ``process_val`` stay the reference's own)."""
    fast = Mmgnet(ref_model.config, ref_model.num_class, ref_model.num_rel)
    fast = fast.to(next(ref_model.parameters()).device)
    adopt_parameters(fast, ref_model)
    fast.train(ref_model.training)
    ref_model._vlsat_fast = [fast]                    # list: keep it out of ref_model._modules (ckpt iteration)
    ref_model.forward = lambda *a, **k: ref_model._vlsat_fast[0].train(ref_model.training)(*a, **k)
    return ref_model
