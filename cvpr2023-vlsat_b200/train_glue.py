"""Training-step glue around the hot path: the losses of ``Mmgnet.process_train``
(src/model/SGFN_MMG/model.py:337-412), the optimiser of ``Mmgnet.__init__`` / ``Mmgnet.backward`` (:136-158, :483-488)
and a ``TrainStep`` that chains forward -> losses -> backward (one CUDA graph) -> gradient all-reduce -> AdamW.

Everything numeric runs in vlsat_b200 kernels (csrc/train_glue.cu) through the C ABI: there is no PyTorch fallback and
CPU tensors raise. What stays outside: the CLIP text encoder behind ``get_rel_emb`` - the caller
passes the ``[E, 512]`` text embedding - and the top-k metric code that follows ``self.backward(loss)``.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import AdamWTensor

LOSS_TERMS = ("loss", "obj_3d", "obj_2d", "rel_3d", "rel_2d", "mimic", "rel_mimic_2d")


def _call(name, *args):
    _lib.check(ops._call(name, *args), name)


def _f32_rows(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError(f"{what}: expected a 2-D float32 CUDA tensor, got {tuple(t.shape)} {t.dtype} on {t.device}")
    return t if t.stride(1) == 1 and t.stride(0) >= t.shape[1] else t.contiguous()


class LossConfig:
    """The ``MODEL`` keys ``process_train`` reads (mmgnet.json values as defaults)."""

    def __init__(self, lambda_o: float = 0.1, weight_edge: str = "DYNAMIC", none_ratio: Optional[float] = 1.0,
                 ignore_none_rel: bool = False, margin: float = 0.8):
        if weight_edge not in ("DYNAMIC", "NONE"):
            # 'BG' / 'OCCU' need per-element / external weights (SGFN_MMG/model.py:348-352, 367-368)
            raise NotImplementedError(f"WEIGHT_EDGE={weight_edge!r}: only 'DYNAMIC' (mmgnet.json) and 'NONE' are built")
        lambda_r = 1.0
        lam = max(lambda_r, lambda_o)                        # SGFN_MMG/model.py:395-399
        self.coef_obj, self.coef_rel, self.coef_mimic = lambda_o / lam, 3.0 * lambda_r / lam, 0.1      # :412
        self.weight_edge, self.none_ratio, self.ignore_none_rel, self.margin = weight_edge, none_ratio, ignore_none_rel, margin

    @classmethod
    def from_mconfig(cls, mconfig) -> "LossConfig":
        from .mmgnet import _get
        if not _get(mconfig, "multi_rel_outputs", True):
            raise NotImplementedError("multi_rel_outputs=false (nll_loss head) is not built")
        return cls(lambda_o=float(_get(mconfig, "lambda_o", 0.1)), weight_edge=_get(mconfig, "WEIGHT_EDGE", "DYNAMIC"),
                   none_ratio=_get(mconfig, "NONE_RATIO", None))


def rel_class_weights(gt_rel: torch.Tensor, ignore_none_rel: bool = False) -> torch.Tensor:
    """WEIGHT_EDGE == 'DYNAMIC' (SGFN_MMG/model.py:353-366): ``|1 / (log(count_c + 1) + 1)|`` per relationship class.
    NONE_RATIO only scales the 'none' slot, which ``weight[1:]`` drops again, so it does not enter."""
    gt = _f32_rows(gt_rel, "gt_rel").contiguous()
    w = torch.empty((gt.shape[1],), device=gt.device, dtype=torch.float32)
    _call("vlsat_rel_class_weights", gt.data_ptr(), gt.shape[0], gt.shape[1], 1e-2 if ignore_none_rel else 1.0, w.data_ptr(), ops._stream())
    return w


class _ReferenceLoss(Function):
    """(loss, terms[6]) = process_train's scalar; backward writes the gradients of the six differentiable inputs."""

    @staticmethod
    def forward(ctx, log3, log2, rel3, rel2, f3, f2, e2d, gt_cls, gt_rel, text, cfg: LossConfig):
        log3, log2, f3, f2, e2d = (_f32_rows(t, n) for t, n in ((log3, "obj_logits_3d"), (log2, "obj_logits_2d"),
                                                               (f3, "obj_feature_3d"), (f2, "obj_feature_2d"), (e2d, "edge_feature_2d")))
        rel3, rel2 = _f32_rows(rel3, "rel_cls_3d").contiguous(), _f32_rows(rel2, "rel_cls_2d").contiguous()
        gt_rel, text = _f32_rows(gt_rel, "gt_rel_cls").contiguous(), _f32_rows(text, "rel_text_feat")
        if gt_cls.dtype != torch.int64 or not gt_cls.is_cuda:
            raise TypeError("gt_cls must be an int64 CUDA tensor")
        gt_cls = gt_cls.contiguous().view(-1)
        n, nc = log3.shape
        e, rc = rel3.shape
        d = f3.shape[1]
        if n == 0 or e == 0:
            raise ValueError("process_train losses need at least one object and one edge (means over empty sets)")
        if tuple(log2.shape) != (n, nc) or tuple(rel2.shape) != (e, rc) or tuple(gt_rel.shape) != (e, rc) or gt_cls.numel() != n \
                or tuple(f2.shape) != (n, d) or tuple(e2d.shape) != tuple(text.shape) or e2d.shape[0] != e:
            raise ValueError("process_train losses: inconsistent shapes")
        dev, st = log3.device, ops._stream()
        rows = torch.empty((4 * n + 3 * e,), device=dev, dtype=torch.float32)
        r_ce3, r_ce2, lse3, lse2 = rows[:n], rows[n:2 * n], rows[2 * n:3 * n], rows[3 * n:4 * n]
        r_b3, r_b2, r_l1 = rows[4 * n:4 * n + e], rows[4 * n + e:4 * n + 2 * e], rows[4 * n + 2 * e:]
        r_cos = torch.empty((n,), device=dev, dtype=torch.float32)
        terms = torch.empty((7,), device=dev, dtype=torch.float32)
        weight = rel_class_weights(gt_rel, cfg.ignore_none_rel) if cfg.weight_edge == "DYNAMIC" else None
        wp = weight.data_ptr() if weight is not None else None
        tp = terms.data_ptr()
        _call("vlsat_cross_entropy_fwd", log3.data_ptr(), log3.stride(0), gt_cls.data_ptr(), n, nc, r_ce3.data_ptr(), lse3.data_ptr(), st)
        _call("vlsat_cross_entropy_fwd", log2.data_ptr(), log2.stride(0), gt_cls.data_ptr(), n, nc, r_ce2.data_ptr(), lse2.data_ptr(), st)
        _call("vlsat_bce_fwd", rel3.data_ptr(), gt_rel.data_ptr(), wp, e, rc, r_b3.data_ptr(), st)
        _call("vlsat_bce_fwd", rel2.data_ptr(), gt_rel.data_ptr(), wp, e, rc, r_b2.data_ptr(), st)
        _call("vlsat_cosine_margin_fwd", f3.data_ptr(), f3.stride(0), f2.data_ptr(), f2.stride(0), n, d, cfg.margin, r_cos.data_ptr(), st)
        _call("vlsat_l1_unit_fwd", e2d.data_ptr(), e2d.stride(0), text.data_ptr(), text.stride(0), e, e2d.shape[1], r_l1.data_ptr(), st)
        # loss = coef_obj (obj_2d + obj_3d) + coef_rel (rel_2d + rel_3d) + 0.1 (mimic + rel_mimic_2d)   (:412)
        for i, (buf, scale, coef) in enumerate(((r_ce3, 1.0 / n, cfg.coef_obj), (r_ce2, 1.0 / n, cfg.coef_obj),
                                                (r_b3, 1.0 / (e * rc), cfg.coef_rel), (r_b2, 1.0 / (e * rc), cfg.coef_rel),
                                                (r_cos, 1.0 / n, cfg.coef_mimic), (r_l1, 1.0 / (e * e2d.shape[1]), cfg.coef_mimic))):
            _call("vlsat_sum_rows", buf.data_ptr(), buf.numel(), scale, tp + 4 * (i + 1), tp, coef, int(i > 0), st)
        ctx.cfg = cfg
        ctx.save_for_backward(log3, log2, rel3, rel2, f3, f2, e2d, gt_cls, gt_rel, text, lse3, lse2, weight)
        loss, parts = terms[0], terms[1:]
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    def backward(ctx, gloss, _gparts):
        log3, log2, rel3, rel2, f3, f2, e2d, gt_cls, gt_rel, text, lse3, lse2, weight = ctx.saved_tensors
        cfg, st = ctx.cfg, ops._stream()
        need = ctx.needs_input_grad
        g = gloss.reshape(1).contiguous().float()
        n, nc = log3.shape
        e, rc = rel3.shape
        d = f3.shape[1]
        wp = weight.data_ptr() if weight is not None else None
        out: List[Optional[torch.Tensor]] = [None] * 11
        for i, (x, lse) in enumerate(((log3, lse3), (log2, lse2))):
            if need[i]:
                out[i] = torch.empty((n, nc), device=x.device, dtype=torch.float32)
                _call("vlsat_cross_entropy_bwd", x.data_ptr(), x.stride(0), gt_cls.data_ptr(), lse.data_ptr(), g.data_ptr(),
                      cfg.coef_obj / n, out[i].data_ptr(), nc, n, nc, st)
        for i, p in ((2, rel3), (3, rel2)):
            if need[i]:
                out[i] = torch.empty((e, rc), device=p.device, dtype=torch.float32)
                _call("vlsat_bce_bwd", p.data_ptr(), gt_rel.data_ptr(), wp, g.data_ptr(), cfg.coef_rel / (e * rc), out[i].data_ptr(), e, rc, st)
        if need[4] or need[5]:
            da = torch.empty((n, d), device=f3.device, dtype=torch.float32) if need[4] else None
            db = torch.empty((n, d), device=f3.device, dtype=torch.float32) if need[5] else None
            _call("vlsat_cosine_margin_bwd", f3.data_ptr(), f3.stride(0), f2.data_ptr(), f2.stride(0), g.data_ptr(), cfg.coef_mimic / n,
                  cfg.margin, da.data_ptr() if need[4] else None, d, db.data_ptr() if need[5] else None, d, n, d, st)
            out[4], out[5] = da, db
        if need[6]:
            de = e2d.shape[1]
            out[6] = torch.empty((e, de), device=e2d.device, dtype=torch.float32)
            _call("vlsat_l1_unit_bwd", e2d.data_ptr(), e2d.stride(0), text.data_ptr(), text.stride(0), g.data_ptr(),
                  cfg.coef_mimic / (e * de), out[6].data_ptr(), de, e, de, st)
        return tuple(out)


def reference_loss(outputs: Sequence[torch.Tensor], gt_cls: torch.Tensor, gt_rel_cls: torch.Tensor, rel_text_feat: torch.Tensor,
                   cfg: Optional[LossConfig] = None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """``process_train``'s loss (SGFN_MMG/model.py:343-412) of the train-mode forward outputs
    ``(obj_logits_3d, obj_logits_2d, rel_cls_3d, rel_cls_2d, obj_feature_3d, obj_feature_2d, edge_feature_2d, ...)``.
    ``rel_text_feat`` [E, 512] is what ``get_rel_emb`` returns. Returns (loss, {term name: 0-dim tensor})."""
    cfg = cfg or LossConfig()
    loss, parts = _ReferenceLoss.apply(*outputs[:7], gt_cls, gt_rel_cls, rel_text_feat, cfg)
    return loss, {name: parts[i] for i, name in enumerate(LOSS_TERMS[1:])}


# ---------------------------------------------------------------------------------------------------------- optimiser
class FusedAdamW:
    """``torch.optim.AdamW`` + ``CosineAnnealingLR(T_max, eta_min=0)`` as ONE multi-tensor kernel launch per step
    (``vlsat_adamw_step``): same ``param_groups`` layout (list of dicts with ``params``, ``lr``, ``weight_decay``,
    ``amsgrad``), same update rule, parameters without a gradient skipped like torch does. The step counter lives on the
    device, so ``step()`` can be captured into a CUDA graph; ``last_lr`` mirrors ``lr_scheduler.get_last_lr()``."""

    def __init__(self, param_groups: Iterable[dict], betas=(0.9, 0.999), eps: float = 1e-8, t_max: int = 0, chunk_elems: int = 16384):
        self.param_groups = []
        for g in param_groups:
            ps = [g["params"]] if isinstance(g["params"], torch.Tensor) else list(g["params"])
            self.param_groups.append(dict(params=ps, lr=float(g["lr"]), weight_decay=float(g.get("weight_decay", 1e-2) or 0.0),
                                          amsgrad=bool(g.get("amsgrad", False))))
        seen = set()
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) in seen:
                    raise ValueError("some parameters appear in more than one parameter group")
                seen.add(id(p))
        self.betas, self.eps, self.t_max, self.chunk_elems = (float(betas[0]), float(betas[1])), float(eps), int(t_max), int(chunk_elems)
        self.state: Dict[int, dict] = {}
        self.steps_done = 0
        self._step_dev: Optional[torch.Tensor] = None
        self._table = None            # (signature, table tensor, chunk tensors, n_chunks, params)

    @property
    def last_lr(self) -> List[float]:
        f = 0.5 * (1.0 + math.cos(math.pi * self.steps_done / self.t_max)) if self.t_max > 0 else 1.0
        return [g["lr"] * f for g in self.param_groups]

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    def _build(self):
        entries, params = [], []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None or not p.requires_grad:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("FusedAdamW: parameters must be contiguous float32 CUDA tensors (no CPU fallback)")
                if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                    p.grad = p.grad.float().contiguous()
                s = self.state.get(id(p))
                if s is None:
                    s = self.state[id(p)] = dict(m=torch.zeros_like(p, memory_format=torch.contiguous_format),
                                                 v=torch.zeros_like(p, memory_format=torch.contiguous_format),
                                                 vmax=torch.zeros_like(p, memory_format=torch.contiguous_format) if g["amsgrad"] else None)
                entries.append((p, s, g))
                params.append(p)
        sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        if self._table is not None and self._table[0] == sig:
            return self._table
        if not entries:
            self._table = (sig, None, None, None, 0, params)
            return self._table
        dev = params[0].device
        arr = (AdamWTensor * len(entries))()
        ct, ci = [], []
        for i, (p, s, g) in enumerate(entries):
            a = arr[i]
            a.p, a.g, a.m, a.v = p.data_ptr(), p.grad.data_ptr(), s["m"].data_ptr(), s["v"].data_ptr()
            a.vmax = s["vmax"].data_ptr() if s["vmax"] is not None else None
            a.n, a.lr, a.weight_decay = p.numel(), g["lr"], g["weight_decay"]
            nch = (p.numel() + self.chunk_elems - 1) // self.chunk_elems
            ct += [i] * nch
            ci += list(range(nch))
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        chunk_tensor = torch.tensor(ct, dtype=torch.int32).to(dev)
        chunk_index = torch.tensor(ci, dtype=torch.int32).to(dev)
        self._table = (sig, table, chunk_tensor, chunk_index, len(ct), params)
        return self._table

    def step(self) -> None:
        sig, table, chunk_tensor, chunk_index, n_chunks, params = self._build()
        if not params:
            return
        if self._step_dev is None:
            self._step_dev = torch.full((1,), self.steps_done, device=params[0].device, dtype=torch.int64)
        _call("vlsat_adamw_step", table.data_ptr(), chunk_tensor.data_ptr(), chunk_index.data_ptr(), n_chunks, self.chunk_elems,
              self.betas[0], self.betas[1], self.eps, self._step_dev.data_ptr(), self.t_max, ops._stream())
        self.steps_done += 1
        self.mark_updated()

    # ---- checkpointing: the reference saves optimizer.state_dict() and lr_scheduler.state_dict() (model_base.py:72-73)
    # and restores them on resume (:115-122); both use torch.optim.AdamW's / CosineAnnealingLR's layout here, so
    # checkpoints move between the two implementations in either direction.
    def state_dict(self) -> dict:
        state, groups, idx = {}, [], 0
        for g, lr_now in zip(self.param_groups, self.last_lr):
            ids = []
            for p in g["params"]:
                s = self.state.get(id(p))
                if s is not None:
                    state[idx] = {"step": torch.tensor(float(self.steps_done)), "exp_avg": s["m"], "exp_avg_sq": s["v"]}
                    if s["vmax"] is not None:
                        state[idx]["max_exp_avg_sq"] = s["vmax"]
                ids.append(idx)
                idx += 1
            groups.append(dict(lr=lr_now, initial_lr=g["lr"], betas=self.betas, eps=self.eps, weight_decay=g["weight_decay"],
                               amsgrad=g["amsgrad"], maximize=False, foreach=None, capturable=False, differentiable=False,
                               fused=None, decoupled_weight_decay=True, params=ids))
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters per group")
        steps = 0
        for sg, g in zip(groups, self.param_groups):
            g["lr"] = float(sg.get("initial_lr", sg["lr"]))
            g["weight_decay"], g["amsgrad"] = float(sg.get("weight_decay", 0.0) or 0.0), bool(sg.get("amsgrad", False))
            for idx, p in zip(sg["params"], g["params"]):
                st = sd["state"].get(idx)
                if st is None:
                    self.state.pop(id(p), None)
                    continue
                conv = lambda t: t.detach().to(device=p.device, dtype=torch.float32).reshape(p.shape).contiguous().clone()
                vmax = conv(st["max_exp_avg_sq"]) if "max_exp_avg_sq" in st else (torch.zeros_like(p, memory_format=torch.contiguous_format) if g["amsgrad"] else None)
                self.state[id(p)] = dict(m=conv(st["exp_avg"]), v=conv(st["exp_avg_sq"]), vmax=vmax)
                steps = max(steps, int(float(st["step"])))
        if groups:
            self.betas = (float(groups[0]["betas"][0]), float(groups[0]["betas"][1]))
            self.eps = float(groups[0]["eps"])
        self.steps_done, self._step_dev, self._table = steps, None, None

    def scheduler_state_dict(self) -> dict:
        """What ``CosineAnnealingLR.state_dict()`` holds for the same point of the schedule."""
        return {"T_max": self.t_max, "eta_min": 0.0, "base_lrs": [g["lr"] for g in self.param_groups], "last_epoch": self.steps_done,
                "_step_count": self.steps_done + 1, "_last_lr": self.last_lr}

    def load_scheduler_state_dict(self, sd: dict) -> None:
        self.t_max = int(sd["T_max"])
        if float(sd.get("eta_min", 0.0)) != 0.0:
            raise NotImplementedError("CosineAnnealingLR with eta_min != 0 is not built (the reference uses the default 0)")
        self.steps_done, self._step_dev = int(sd["last_epoch"]), None

    def mark_updated(self) -> None:
        """The kernel writes the parameters behind autograd's back: bump their version counters so that every derived
        weight cache (packed / folded / split copies keyed on ``Tensor._version``) refreshes."""
        ps = [p for g in self.param_groups for p in g["params"]]
        if ps:
            torch.autograd.graph.increment_version(ps)


def reference_param_groups(model: torch.nn.Module, lr: float, weight_decay=False, amsgrad=False) -> List[dict]:
    """The 13 parameter groups of ``Mmgnet.__init__`` (SGFN_MMG/model.py:136-156), in the reference's order."""
    mmg_obj, mmg_rel = [], []
    for name, p in model.mmg.named_parameters():
        (mmg_rel if "nn_edge" in name else mmg_obj).append(p)
    wd = float(weight_decay or 0.0)                 # mmgnet.json: "W_DECAY": false
    g = lambda ps, scale=1.0: dict(params=list(ps), lr=float(lr) * scale, weight_decay=wd, amsgrad=bool(amsgrad))
    return [g(model.obj_encoder.parameters()), g(model.rel_encoder_2d.parameters()), g(model.rel_encoder_3d.parameters()),
            g(mmg_obj, 1 / 4), g(mmg_rel, 1 / 2), g(model.obj_predictor_2d.parameters(), 1 / 10),
            g(model.rel_predictor_2d.parameters()), g(model.obj_predictor_3d.parameters(), 1 / 10),
            g(model.rel_predictor_3d.parameters()), g(model.mlp_3d.parameters()), g(model.triplet_projector_3d.parameters()),
            g(model.triplet_projector_2d.parameters()), g([model.obj_logit_scale])]


def build_optimizer(model: torch.nn.Module, lr: float = 1e-4, weight_decay=False, amsgrad=False, max_iteration: int = 0) -> FusedAdamW:
    """``optim.AdamW(13 groups)`` + ``CosineAnnealingLR(T_max=max_iteration)`` of the reference (:143-157)."""
    return FusedAdamW(reference_param_groups(model, lr, weight_decay, amsgrad), t_max=max_iteration)


# ------------------------------------------------------------------------------------- N4: cached text supervision
class RelTextCache:
    """``Mmgnet.get_rel_emb`` (SGFN_MMG/model.py:221-255) without the per-step python loop and text-encoder call: the
    features of every prompt the function can build are computed ONCE into a table [S, O, R + 1, D] (160 * 160 * 27 prompts
    for mmgnet.json = 1.4 GB fp32), and the target of a batch is one gather-mean-normalise kernel (``vlsat_rel_text_embed``).

    ``fill(encode)`` takes any ``encode(list_of_prompts) -> [n, D]`` (the reference's: ``lambda p: clip_model.encode_text(
    clip.tokenize(p).cuda())``; CLIP weights are not available offline, so this package ships none). ``from_table`` adopts
    a table computed elsewhere. ``__call__(gt_cls, gt_rel_cls, edge_indices [E, 2])`` -> [E, D] fp32, the ``rel_text_feat``
    argument of ``reference_loss`` / ``TrainStep.step``."""

    def __init__(self, obj_names, rel_names, dim: int = 512, device="cuda"):
        self.obj_names, self.rel_names, self.dim = list(obj_names), list(rel_names), dim
        self.table = torch.zeros((len(self.obj_names), len(self.obj_names), len(self.rel_names) + 1, dim), device=device, dtype=torch.float32)
        self.filled = False

    @classmethod
    def from_table(cls, table: torch.Tensor):
        s, o, r1, d = table.shape
        me = cls([None] * s, [None] * (r1 - 1), d, table.device)
        me.table, me.filled = table.to(torch.float32).contiguous(), True
        return me

    def prompts(self, s: int):
        """The (O * (R + 1)) prompts of subject class ``s`` in table order (:232-240)."""
        a = self.obj_names[s]
        out = []
        for b in self.obj_names:
            out += [f"a point cloud of a {a} {rel} a {b}" for rel in self.rel_names]
            out.append(f"the {a} and the {b} has no relation in the point cloud")
        return out

    def fill(self, encode, batch: int = 4096) -> "RelTextCache":
        o, r1 = len(self.obj_names), len(self.rel_names) + 1
        with torch.no_grad():
            for s in range(len(self.obj_names)):
                p = self.prompts(s)
                feats = torch.cat([encode(p[i:i + batch]).to(self.table.device, torch.float32) for i in range(0, len(p), batch)])
                self.table[s] = feats.view(o, r1, self.dim)
        self.filled = True
        return self

    def __call__(self, gt_cls: torch.Tensor, gt_rel_cls: torch.Tensor, edge_indices: torch.Tensor) -> torch.Tensor:
        if not self.filled:
            raise RuntimeError("RelTextCache: fill(encode) or from_table(...) first (no text encoder ships with vlsat_b200)")
        if edge_indices.dim() != 2 or edge_indices.shape[1] != 2:
            raise ValueError("edge_indices must be [E, 2] (subject, object) as process_train holds them")
        e = edge_indices.shape[0]
        rel = gt_rel_cls.to(torch.float32).contiguous()
        ed, cls_ = edge_indices.to(torch.int64).contiguous(), gt_cls.to(torch.int64).contiguous()
        out = torch.empty((e, self.dim), device=self.table.device, dtype=torch.float32)
        s, o, r1, d = self.table.shape
        _lib.check(ops._call("vlsat_rel_text_embed", self.table.data_ptr(), s, r1 - 1, d, cls_.data_ptr(), rel.data_ptr(), rel.stride(0) if e else r1 - 1,
                         ed.data_ptr(), e, out.data_ptr(), d, ops._stream()), "vlsat_rel_text_embed")
        return out


# --------------------------------------------------------------------------------------------------------- train step
class TrainStep:
    """One ``process_train`` iteration up to and including ``self.backward(loss)`` (SGFN_MMG/model.py:337-413, 483-488):

        forward(istrain=True) -> reference_loss -> loss.backward()      one CUDA graph replay (graph.GraphedTrainStep)
        gradient all-reduce (mean) over the data-parallel group          dist.GradientAllReducer, N > 1 only
        AdamW + cosine schedule                                          one multi-tensor kernel (FusedAdamW)

    ``step(obj_points, obj_2d_feats, edge_index [2, E], descriptor, batch_ids, gt_cls, gt_rel_cls, rel_text_feat)``
    returns ``(loss, terms)`` as device tensors (no host sync; they are overwritten by the next step)."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdamW, loss_cfg: Optional[LossConfig] = None, reducer=None,
                 graphed: bool = True, metrics: bool = False):
        from .graph import GraphedTrainStep
        self.model, self.optimizer, self.cfg, self.reducer = model, optimizer, loss_cfg or LossConfig(), reducer
        # metrics=True: the twelve recall figures process_train logs after backward() (SGFN_MMG/model.py:422-432), computed
        # by the rank kernels inside the same captured step and left in ``self.metrics`` as device scalars
        self.with_metrics, self.metrics = metrics, None
        self._targets: Dict[Tuple, Tuple[torch.Tensor, ...]] = {}
        self._current: Optional[Tuple[torch.Tensor, ...]] = None
        self.terms: Optional[Dict[str, torch.Tensor]] = None
        self._graphed = GraphedTrainStep(model, self._loss) if graphed else None

    def _loss(self, outs):
        loss, self.terms = reference_loss(outs, *self._current, cfg=self.cfg)
        if self.with_metrics:
            from .eval_ranks import train_metrics
            self.metrics = train_metrics(outs[0], outs[1], outs[2], outs[3], self._current[0], self._current[1])
        return loss

    def step(self, obj_points, obj_2d_feats, edge_index, descriptor, batch_ids, gt_cls, gt_rel_cls, rel_text_feat, scene_stats=None):
        loss = self.forward_backward(obj_points, obj_2d_feats, edge_index, descriptor, batch_ids, gt_cls, gt_rel_cls, rel_text_feat,
                                     scene_stats=scene_stats)
        if self.reducer is not None:
            self.reducer.allreduce()
        self.optimizer.step()
        if self._graphed is None:
            self.optimizer.zero_grad(set_to_none=True)       # graph replays overwrite their static gradient buffers
        return loss, self.terms

    def forward_backward(self, obj_points, obj_2d_feats, edge_index, descriptor, batch_ids, gt_cls, gt_rel_cls, rel_text_feat, scene_stats=None):
        """forward(istrain=True) -> the six loss terms -> loss.backward(): everything of ``step`` before the collective and the
        optimiser (one CUDA graph replay). Every parameter's ``.grad`` holds this batch's gradient afterwards."""
        targets = (gt_cls, gt_rel_cls, rel_text_feat)
        if self._graphed is not None:
            key = tuple((tuple(t.shape), t.dtype) for t in targets) + (tuple(edge_index.shape), tuple(obj_points.shape))
            static = self._targets.get(key)
            if static is None:
                static = self._targets[key] = tuple(t.clone() for t in targets)
            for dst, src in zip(static, targets):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self._current = static
            loss, _ = self._graphed(obj_points, obj_2d_feats, edge_index, descriptor, batch_ids, scene_stats=scene_stats)
        else:
            self._current = targets
            with ops.zero_arena(self, obj_points.device):
                outs = self.model(obj_points, obj_2d_feats, edge_index, descriptor, batch_ids, istrain=True)
                loss = self._loss(outs)
                loss.backward()
        return loss

    @property
    def kernels_per_step(self) -> int:
        g = self._graphed.kernels_per_replay if self._graphed is not None else 0
        return g + 2
