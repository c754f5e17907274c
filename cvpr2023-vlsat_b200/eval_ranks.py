"""Evaluation ranks of ``--mode eval`` on the GPU: counting kernels in place of the reference's
per-edge python loops around CPU sorts (src/utils/eva_utils_acc.py:27-79, 137-211; ``Mmgnet.process_val``,
SGFN_MMG/model.py:463-472). Same function names and result order as the reference; the ground truth is passed as the
tensors ``process_val`` holds (``gt_cls`` [N], ``gt_rel_cls`` [E, 26], ``edge_indices`` [E, 2]) instead of ``get_gt``'s
python list. Results stay on the device (flat int64, the reference returns numpy arrays). No CPU fallback.

First run on a B200 in round 2: tests/test_eval_ranks_gpu.py (bit-exact against the oracle, which is pinned on the
reference's own functions and fixtures) is part of ``-m gpu``. ``train_metrics`` is the device-side form of the metric
code that follows ``self.backward(loss)`` in ``process_train`` (SGFN_MMG/model.py:422-432).
"""
from __future__ import annotations

import torch

from . import _lib, ops


def _f32(t: torch.Tensor, what: str, dim: int = 2) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != dim:
        raise TypeError(f"{what}: expected a {dim}-D float32 CUDA tensor, got {tuple(t.shape)} {t.dtype} on {t.device}")
    return t.contiguous()


def _i64(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.int64:
        raise TypeError(f"{what}: expected an int64 CUDA tensor, got {t.dtype} on {t.device}")
    return t.contiguous()


def _flatten(ranks: torch.Tensor) -> torch.Tensor:
    return ranks[ranks > -2 ** 31].to(torch.int64)      # row-major: edges in order, each edge's entries ascending; INT32_MIN = empty slot


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "softmax_rows")
    y = torch.empty_like(x)
    _lib.check(ops._call("vlsat_softmax_rows", x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], y.data_ptr(), ops._stream()), "vlsat_softmax_rows")
    return y


def evaluate_topk_object(objs_pred: torch.Tensor, objs_target: torch.Tensor, topk: int) -> torch.Tensor:
    """eva_utils_acc.py:27-39 -> [N] int64."""
    p, t = _f32(objs_pred, "objs_pred"), _i64(objs_target, "objs_target").view(-1)
    out = torch.empty((p.shape[0],), device=p.device, dtype=torch.int32)
    _lib.check(ops._call("vlsat_topk_object_ranks", p.data_ptr(), p.stride(0), t.data_ptr(), p.shape[0], p.shape[1], int(topk),
                         out.data_ptr(), ops._stream()), "vlsat_topk_object_ranks")
    return out.to(torch.int64)


def evaluate_topk_predicate(rels_pred: torch.Tensor, gt_rel_cls: torch.Tensor, topk: int, confidence_threshold: float = 0.5) -> torch.Tensor:
    """eva_utils_acc.py:42-79 (multi_rel_outputs=True) -> flat int64, one entry per (edge, label) or per label-free edge."""
    p, y = _f32(rels_pred, "rels_pred"), _f32(gt_rel_cls, "gt_rel_cls")
    if p.shape != y.shape:
        raise ValueError("rels_pred and gt_rel_cls must have the same shape")
    out = torch.empty(p.shape, device=p.device, dtype=torch.int32)
    _lib.check(ops._call("vlsat_topk_predicate_ranks", p.data_ptr(), y.data_ptr(), p.shape[0], p.shape[1], int(topk),
                         float(confidence_threshold), out.data_ptr(), ops._stream()), "vlsat_topk_predicate_ranks")
    return _flatten(out)


def evaluate_triplet_topk(objs_pred: torch.Tensor, rels_pred: torch.Tensor, gt_cls: torch.Tensor, gt_rel_cls: torch.Tensor,
                          edge_indices: torch.Tensor, topk: int, confidence_threshold: float = 0.5, use_clip: bool = True,
                          obj_probs: torch.Tensor = None) -> torch.Tensor:
    """eva_utils_acc.py:137-211, the ``res`` array (ranks). ``edge_indices`` [E, 2] (subject, object). ``use_clip=True``:
    ``objs_pred`` are logits, softmaxed here (:143-145); ``use_clip=False``: log-probabilities are not built - pass
    ``obj_probs`` (already normalised scores) to skip the softmax, e.g. to reproduce a CPU softmax bit for bit."""
    if obj_probs is None:
        if not use_clip:
            raise NotImplementedError("use_clip=False (np.exp of log_softmax scores): pass obj_probs instead")
        obj_probs = softmax_rows(objs_pred)
    q, r, y = _f32(obj_probs, "obj_probs"), _f32(rels_pred, "rels_pred"), _f32(gt_rel_cls, "gt_rel_cls")
    g, ed = _i64(gt_cls, "gt_cls").view(-1), _i64(edge_indices, "edge_indices")
    if ed.dim() != 2 or ed.shape[1] != 2 or ed.shape[0] != r.shape[0] or y.shape != r.shape or g.numel() != q.shape[0]:
        raise ValueError("evaluate_triplet_topk: inconsistent shapes (edge_indices must be [E, 2])")
    out = torch.empty(r.shape, device=r.device, dtype=torch.int32)
    _lib.check(ops._call("vlsat_topk_triplet_ranks", q.data_ptr(), q.shape[0], q.shape[1], r.data_ptr(), r.shape[1], g.data_ptr(),
                         y.data_ptr(), ed.data_ptr(), r.shape[0], int(topk), float(confidence_threshold), out.data_ptr(),
                         ops._stream()), "vlsat_topk_triplet_ranks")
    return _flatten(out)


def _ranks_raw(kind: str, pred, target, topk: int, threshold: float = 0.5) -> torch.Tensor:
    """int32 rank array in the kernels' own layout (INT32_MIN = empty slot): no boolean indexing, hence no host sync."""
    if kind == "obj":
        p, t = _f32(pred, "objs_pred"), _i64(target, "objs_target").view(-1)
        out = torch.empty((p.shape[0],), device=p.device, dtype=torch.int32)
        _lib.check(ops._call("vlsat_topk_object_ranks", p.data_ptr(), p.stride(0), t.data_ptr(), p.shape[0], p.shape[1], int(topk),
                             out.data_ptr(), ops._stream()), "vlsat_topk_object_ranks")
        return out
    p, y = _f32(pred, "rels_pred"), _f32(target, "gt_rel_cls")
    out = torch.empty(p.shape, device=p.device, dtype=torch.int32)
    _lib.check(ops._call("vlsat_topk_predicate_ranks", p.data_ptr(), y.data_ptr(), p.shape[0], p.shape[1], int(topk), float(threshold),
                         out.data_ptr(), ops._stream()), "vlsat_topk_predicate_ranks")
    return out


def recall_at(ranks: torch.Tensor, thresholds) -> torch.Tensor:
    """[3] float32 on the device: 100 * (ranks <= t).sum() / len(ranks) for three thresholds (SGFN_MMG/model.py:425-426)."""
    t = [int(x) for x in thresholds]
    out = torch.empty((3,), device=ranks.device, dtype=torch.float32)
    _lib.check(ops._call("vlsat_recall_at", ranks.data_ptr(), ranks.numel(), t[0], t[1], t[2], out.data_ptr(), ops._stream()), "vlsat_recall_at")
    return out


def train_metrics(obj_logits_3d, obj_logits_2d, rel_cls_3d, rel_cls_2d, gt_cls, gt_rel_cls) -> dict:
    """The twelve recall figures ``process_train`` computes after ``self.backward(loss)`` (SGFN_MMG/model.py:422-432), as
    device scalars under the reference's log names - no ``.item()``, no python loop over edges, CUDA-graph capturable."""
    o3 = recall_at(_ranks_raw("obj", obj_logits_3d.detach(), gt_cls, 11), (1, 5, 10))
    o2 = recall_at(_ranks_raw("obj", obj_logits_2d.detach(), gt_cls, 11), (1, 5, 10))
    r3 = recall_at(_ranks_raw("rel", rel_cls_3d.detach(), gt_rel_cls, 6), (1, 3, 5))
    r2 = recall_at(_ranks_raw("rel", rel_cls_2d.detach(), gt_rel_cls, 6), (1, 3, 5))
    return {"train/Obj_R1": o3[0], "train/Obj_R5": o3[1], "train/Obj_R10": o3[2],
            "train/Pred_R1": r3[0], "train/Pred_R3": r3[1], "train/Pred_R5": r3[2],
            "train/Obj_R1_2d": o2[0], "train/Obj_R5_2d": o2[1], "train/Obj_R10_2d": o2[2],
            "train/Pred_R1_2d": r2[0], "train/Pred_R3_2d": r2[1], "train/Pred_R5_2d": r2[2]}
