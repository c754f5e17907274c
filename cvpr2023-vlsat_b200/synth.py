"""Seeded synthetic 3DSSG-shaped scenes, collated the way the reference hands them to ``Mmgnet.forward``.

Layout contract (reference file:line):
  * per-object descriptor = centroid3, std3 (unbiased), dims3, volume, max-length computed on the raw
    points (``src/utils/op_utils.py:47-64``), then the points are made zero-mean per object
    (``src/dataset/dataset_3dssg.py:293``);
  * edges are all ordered pairs ``i != j`` in ``itertools.product`` order
    (``src/dataset/dataset_3dssg.py:263-266``), optionally thinned to ``edges_per_scene`` keeping order;
  * scenes are concatenated, edge indices offset by the running node count and ``batch_ids`` records
    the scene of every node (``src/dataset/DataLoader.py:153-176``);
  * the trainer permutes points to channels-first ``[sum_N, 3, P]`` (``src/model/model.py:69-74``) and
    the model receives ``edge_indices`` transposed to ``[2, sum_E]`` (``src/model/SGFN_MMG/model.py:340``).

Host-side only (numpy/torch CPU); nothing here runs on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Union

import torch


@dataclass
class SceneBatch:
    obj_points: torch.Tensor      # [sum_N, 3, P] fp32, zero-mean per object
    obj_2d_feats: torch.Tensor    # [sum_N, 512] fp32
    edge_indices: torch.Tensor    # [2, sum_E] int64 (row 0 = subject/source, row 1 = object/destination)
    descriptor: torch.Tensor      # [sum_N, 11] fp32
    batch_ids: torch.Tensor       # [sum_N, 1] int64, non-decreasing scene ids
    num_scenes: int

    def to(self, device, non_blocking: bool = False) -> "SceneBatch":
        return SceneBatch(*(t.to(device, non_blocking=non_blocking) for t in self.tensors()),
                          num_scenes=self.num_scenes)

    def pin(self) -> "SceneBatch":
        return SceneBatch(*(t.pin_memory() for t in self.tensors()), num_scenes=self.num_scenes)

    def tensors(self):
        return (self.obj_points, self.obj_2d_feats, self.edge_indices, self.descriptor, self.batch_ids)

    def forward_args(self):
        """Positional arguments of ``Mmgnet.forward`` after ``istrain``."""
        return self.tensors()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())


def descriptor_of(points: torch.Tensor) -> torch.Tensor:
    """``points`` [N, P, 3] -> [N, 11]; same quantities as ``gen_descriptor`` (op_utils.py:47-64)."""
    centroid = points.mean(1)
    std = points.std(1)                      # unbiased, like torch's default in the reference
    dims = points.max(1)[0] - points.min(1)[0]
    volume = (dims[:, 0] * dims[:, 1] * dims[:, 2]).unsqueeze(1)
    length = dims.max(1)[0].unsqueeze(1)
    return torch.cat([centroid, std, dims, volume, length], dim=1)


def _scene_edges(n: int, edges_per_scene: Optional[int], shuffle: bool, gen: torch.Generator) -> torch.Tensor:
    idx = torch.arange(n)
    src = idx.repeat_interleave(n)
    dst = idx.repeat(n)
    keep = src != dst                       # itertools.product order with the diagonal dropped
    e = torch.stack([src[keep], dst[keep]], dim=0)
    total = e.shape[1]
    if edges_per_scene is not None and edges_per_scene < total:
        sel = torch.randperm(total, generator=gen)[:edges_per_scene].sort()[0]
        e = e[:, sel]
    if shuffle and e.shape[1] > 1:
        e = e[:, torch.randperm(e.shape[1], generator=gen)]
    return e


def make_batch(num_scenes: int,
               objects_per_scene: Union[int, Sequence[int]],
               points_per_object: int,
               edges_per_scene: Optional[int] = None,
               seed: int = 0,
               shuffle_edges: bool = False,
               feat_dim: int = 512) -> SceneBatch:
    """Build one collated batch. ``objects_per_scene`` may be a list (ragged scenes)."""
    gen = torch.Generator().manual_seed(seed)
    if isinstance(objects_per_scene, int):
        counts = [objects_per_scene] * num_scenes
    else:
        counts = list(objects_per_scene)
        assert len(counts) == num_scenes
    pts_l, feat_l, edge_l, desc_l, bid_l = [], [], [], [], []
    offset = 0
    for s, n in enumerate(counts):
        scale = torch.rand(n, 1, 3, generator=gen) * 0.9 + 0.1
        centre = torch.rand(n, 1, 3, generator=gen) * 6.0 - 3.0
        pts = torch.randn(n, points_per_object, 3, generator=gen) * scale + centre
        desc = descriptor_of(pts)
        pts = pts - pts.mean(1, keepdim=True)
        pts_l.append(pts)
        desc_l.append(desc)
        feat_l.append(torch.randn(n, feat_dim, generator=gen))
        edge_l.append(_scene_edges(n, edges_per_scene, shuffle_edges, gen) + offset)
        bid_l.append(torch.full((n, 1), s, dtype=torch.int64))
        offset += n
    obj_points = torch.cat(pts_l, 0).permute(0, 2, 1).contiguous().float()
    return SceneBatch(obj_points=obj_points,
                      obj_2d_feats=torch.cat(feat_l, 0).float().contiguous(),
                      edge_indices=torch.cat(edge_l, 1).long().contiguous(),
                      descriptor=torch.cat(desc_l, 0).float().contiguous(),
                      batch_ids=torch.cat(bid_l, 0).contiguous(),
                      num_scenes=num_scenes)


# The five BASELINE.json configurations, by name.
CONFIGS = {
    "cfg1": dict(num_scenes=1, objects_per_scene=10, points_per_object=128, edges_per_scene=30),
    "cfg2": dict(num_scenes=16, objects_per_scene=40, points_per_object=256, edges_per_scene=600),
    "cfg3": dict(num_scenes=64, objects_per_scene=40, points_per_object=512, edges_per_scene=None),
    "cfg4_per_gpu": dict(num_scenes=32, objects_per_scene=40, points_per_object=256, edges_per_scene=600),
    # BASELINE config #5: 3RScan-shaped sub-scenes (2..9 objects, 3DSSG statistics), fully connected, 128 points per object,
    # eval mode; the reference evaluates them one scene per forward (src/model/model.py:185), here 256 scenes share a batch.
    # The object counts are a fixed seeded draw so that every batch of a run has the same shape signature (one CUDA graph).
    "cfg5": dict(num_scenes=256, objects_per_scene=torch.randint(2, 10, (256,), generator=torch.Generator().manual_seed(7919)).tolist(),
                 points_per_object=128, edges_per_scene=None),
}


def make_config_batch(name: str, seed: int = 0, num_scenes: Optional[int] = None, **over) -> SceneBatch:
    kw = dict(CONFIGS[name])
    if num_scenes is not None:
        kw["num_scenes"] = num_scenes
    kw.update(over)
    return make_batch(seed=seed, **kw)


def make_real_shaped_batch(num_scenes: int, seed: int = 0, points_per_object: int = 128) -> SceneBatch:
    """Config #5: 3DSSG-like sub-scenes with 2..9 objects, fully connected."""
    gen = torch.Generator().manual_seed(seed + 7919)
    counts = torch.randint(2, 10, (num_scenes,), generator=gen).tolist()
    return make_batch(num_scenes, counts, points_per_object, None, seed=seed)


def shard_scenes(batch: SceneBatch, rank: int, world: int) -> SceneBatch:
    """Scene-parallel split: rank r keeps scenes {r, r+W, ...}, re-based to local node ids."""
    bids = batch.batch_ids.view(-1)
    mine = torch.arange(rank, batch.num_scenes, world)
    node_keep = torch.isin(bids, mine)
    new_id = torch.full((bids.numel(),), -1, dtype=torch.int64)
    new_id[node_keep] = torch.arange(int(node_keep.sum()))
    src, dst = batch.edge_indices
    edge_keep = node_keep[src]
    e = torch.stack([new_id[src[edge_keep]], new_id[dst[edge_keep]]], 0)
    scene_map = torch.full((batch.num_scenes,), -1, dtype=torch.int64)
    scene_map[mine] = torch.arange(mine.numel())
    return SceneBatch(obj_points=batch.obj_points[node_keep].contiguous(),
                      obj_2d_feats=batch.obj_2d_feats[node_keep].contiguous(),
                      edge_indices=e.contiguous(),
                      descriptor=batch.descriptor[node_keep].contiguous(),
                      batch_ids=scene_map[bids[node_keep]].view(-1, 1).contiguous(),
                      num_scenes=int(mine.numel()))


# --------------------------------------------------------------------------------------------------
# Seeded weights: values depend only on (key name, shape, seed), so the build container (where the
# reference can be imported to make golden vectors) and the GPU box (where it cannot) regenerate the
# identical state_dict without shipping 110 MB of parameters.
# --------------------------------------------------------------------------------------------------
def _key_seed(key: str, seed: int) -> int:
    import zlib
    return (zlib.crc32(key.encode()) * 2654435761 + seed * 97531) % (2 ** 31 - 1)


def seeded_tensor(key: str, shape, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    shape = tuple(shape)
    gen = torch.Generator().manual_seed(_key_seed(key, seed))
    leaf = key.split(".")[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_var":
        return torch.rand(shape, generator=gen) + 0.5
    if leaf == "running_mean":
        return torch.randn(shape, generator=gen) * 0.1
    if len(shape) == 0:                                   # logit scales
        return torch.tensor(2.6592600, dtype=dtype)
    if len(shape) == 1:
        if leaf == "weight":                              # LayerNorm / BatchNorm gains
            return 1.0 + 0.1 * torch.randn(shape, generator=gen)
        return 0.05 * torch.randn(shape, generator=gen)   # biases (non-zero on purpose)
    if key.startswith("obj_predictor") and leaf == "weight":
        w = torch.randn(shape, generator=gen)
        return w / w.norm(dim=-1, keepdim=True)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = (3.0 / fan_in) ** 0.5 * 1.4                   # variance ~2/fan_in keeps ReLU stacks O(1)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def make_state_dict(schema, seed: int = 0):
    """``schema``: mapping key -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``)."""
    return {k: seeded_tensor(k, tuple(shape), seed) for k, shape in schema.items()}


def load_seeded(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    sd = module.state_dict()
    module.load_state_dict(make_state_dict({k: v.shape for k, v in sd.items()}, seed))
    return module
