"""Device-side object preparation of the input pipeline.

``SSGDatasetGraph.data_preparation`` (src/dataset/dataset_3dssg.py:244-367) runs, per object and on the loader's CPU,
``obj_pointset = points[choice]``, ``gen_descriptor`` (src/utils/op_utils.py:47-64), ``zero_mean`` (:189-195, :293), and
the trainer then permutes to channels-first (src/model/model.py:71). ``prepare_objects`` does all four in one kernel from
a scan cloud resident on the GPU. What stays on the host: loading the scan (trimesh / ply), choosing the objects and
edges of a sub-scene and drawing the ``np.random.choice`` indices - the caller passes them (offset into the cloud).
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

import torch

from . import _lib, ops


def prepare_objects(cloud: torch.Tensor, choice: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """cloud [M, C] float32 (xyz first, then rgb / normals when USE_RGB / USE_NORMAL), choice [N, P] int64 row indices
    into ``cloud`` -> (obj_points [N, C, P] centred xyz, channels-first, as ``Mmgnet.forward`` takes it;
    descriptor [N, 11] = centroid, std, extent, volume, longest side of the sampled points before centring)."""
    if not cloud.is_cuda or cloud.dtype != torch.float32 or cloud.dim() != 2 or cloud.shape[1] < 3 or cloud.shape[0] < 1:
        raise TypeError(f"prepare_objects: cloud must be a non-empty [M, C >= 3] float32 CUDA tensor, got {tuple(cloud.shape)} {cloud.dtype} on {cloud.device}")
    if not choice.is_cuda or choice.dtype != torch.int64 or choice.dim() != 2 or choice.shape[1] < 1:
        raise TypeError("prepare_objects: choice must be an [N, P >= 1] int64 CUDA tensor")
    if cloud.stride(1) != 1:
        cloud = cloud.contiguous()
    choice = choice.contiguous()
    n, p = choice.shape
    c = cloud.shape[1]
    obj_points = torch.empty((n, c, p), device=cloud.device, dtype=torch.float32)
    descriptor = torch.empty((n, 11), device=cloud.device, dtype=torch.float32)
    _lib.check(ops._call("vlsat_object_prep_fwd", cloud.data_ptr(), cloud.stride(0), cloud.shape[0], c, choice.data_ptr(), n, p,
                         obj_points.data_ptr(), descriptor.data_ptr(), ops._stream(),
                         work=(0.0, n * p * (8.0 + 8.0 * c))), "vlsat_object_prep_fwd")
    return obj_points, descriptor


def sample_object_indices(instances: np.ndarray, nodes: Sequence[int], num_points: int, rng=np.random) -> np.ndarray:
    """Host side of the loader's per-object loop (src/dataset/dataset_3dssg.py:279-290): for every node, in order,
    ``rows = np.where(instances == id)[0]`` and ``choice = np.random.choice(len(rows), num_points, replace=True)`` - one RNG
    call per node exactly like the reference, so its random stream is preserved - returned as GLOBAL row indices
    ``rows[choice]`` into the scan cloud, [N, num_points] int64."""
    out = np.empty((len(nodes), num_points), dtype=np.int64)
    for i, instance_id in enumerate(nodes):
        rows = np.where(instances == instance_id)[0]
        if len(rows) == 0:
            raise ValueError(f"instance {instance_id} has no points in this scan")
        out[i] = rows[rng.choice(len(rows), num_points, replace=True)]
    return out


def prepare_scene(points, instances: np.ndarray, nodes: Sequence[int], num_points: int, rng=np.random, device="cuda"):
    """Loader hook for ``SSGDatasetGraph.data_preparation`` (dataset_3dssg.py:279-293): the scan cloud ``points`` [M, C]
    (numpy or a CUDA tensor already resident, e.g. cached per scan) and its per-point instance ids -> ``(obj_points
    [N, C, P], descriptor [N, 11])`` on the device, in the layout ``Mmgnet.forward`` takes (src/model/model.py:71).
    Sampling stays on the host with the reference's RNG stream; gather, descriptor, centring and the channels-first
    permute are one kernel (csrc/object_prep.cu)."""
    choice = torch.from_numpy(sample_object_indices(instances, nodes, num_points, rng)).to(device, non_blocking=True)
    cloud = points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32))
    return prepare_objects(cloud.to(device, non_blocking=True), choice)
