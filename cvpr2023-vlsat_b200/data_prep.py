"""Device-side object preparation of the input pipeline.

``SSGDatasetGraph.data_preparation`` (src/dataset/dataset_3dssg.py:244-367) runs, per object and on the loader's CPU,
``obj_pointset = points[choice]``, ``gen_descriptor`` (src/utils/op_utils.py:47-64), ``zero_mean`` (:189-195, :293), and
the trainer then permutes to channels-first (src/model/model.py:71). ``prepare_objects`` does all four in one kernel from
a scan cloud resident on the GPU. What stays on the host: loading the scan (trimesh / ply), choosing the objects and
edges of a sub-scene and drawing the ``np.random.choice`` indices - the caller passes them (offset into the cloud).
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Tuple

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
