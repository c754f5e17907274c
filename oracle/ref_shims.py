"""TEST INFRASTRUCTURE (not product code): import the unmodified reference from /root/reference.

Only usable in the build container (``/root/reference`` does not exist on the GPU box). Used by
``oracle/make_golden.py`` to produce the fixtures under ``tests/golden/`` and by the ``not gpu`` tests
that pin ``oracle/vlsat_oracle.py`` against the live reference when it is present.

The reference imports five things that are absent here; each gets a minimal stand-in:
  1. ``torch_geometric.nn.conv.MessagePassing`` - third-party, not vendored, not pinned by the
     reference (README.md:27-31 installs "torch-geometric" with torch 1.12.1). Only its private
     gather/aggregate plumbing is used (``network_util.py:56-58,68-72``, ``op_utils.py:73-75``). The
     stand-in restates the PyG <= 2.3 behaviour: ``x_i = x[edge_index[i]]``, ``x_j = x[edge_index[j]]``
     with ``(i, j) = (1, 0)`` for ``source_to_target`` and ``(0, 1)`` for ``target_to_source``;
     ``aggregate`` scatters onto ``edge_index[i]`` with max / add / mean, empty rows = 0
     (torch_scatter semantics). PARITY UNPINNED at this boundary: no reference test covers it; the
     anchor is the demo in ``network_util.py:75-99`` (see tests/test_oracle.py).
  2. ``src.lib.pointnet.graph`` - missing from the reference repo itself (network_PointNet.py:11).
  3. ``tkinter`` (model_base.py:2 ``from tkinter import N``, unused).
  4. ``clip`` - needs network + weights; ``Mmgnet.get_label_weight`` is patched to return seeded
     unit-norm text features.
  5. ``torch.Tensor.cuda`` is made a no-op on CPU-only hosts (network_MMG.py:185-186 hard-code it).
"""
from __future__ import annotations

import inspect
import os
import sys
import tempfile
import types

import torch

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _resolve_root() -> str:
    """VLSAT_REFERENCE_ROOT if set, else /root/reference (build container), else the staged byte-for-byte copy of the path's
    files under baseline/_ref (GPU box; oracle/stage_reference.py). Resolved at call time: test modules import this file
    in any order."""
    env = os.environ.get("VLSAT_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/src/model"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _resolve_root()


def reference_available() -> bool:
    global REF_ROOT
    REF_ROOT = _resolve_root()
    return os.path.isdir(os.path.join(REF_ROOT, "src", "model"))


class _MessagePassing(torch.nn.Module):
    """Minimal restatement of the PyG <= 2.3 private API used by the reference."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self.__user_args__ = [a for a in inspect.signature(self.message).parameters]
        outer = self

        class _Inspector:
            def distribute(self, name, coll):
                fn = getattr(outer, name)
                params = inspect.signature(fn).parameters
                return {k: coll[k] for k in params if k in coll}

        self.inspector = _Inspector()

    def __check_input__(self, edge_index, size):
        assert edge_index.dtype == torch.long and edge_index.dim() == 2 and edge_index.size(0) == 2
        return [None, None]

    def __collect__(self, args, edge_index, size, kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        out = {}
        for arg in args:
            if arg[-2:] not in ("_i", "_j"):
                continue
            data = kwargs.get(arg[:-2])
            if data is None:
                continue
            out[arg] = data.index_select(self.node_dim, edge_index[i if arg.endswith("_i") else j])
        out["index"] = edge_index[i]
        out["edge_index_i"], out["edge_index_j"] = edge_index[i], edge_index[j]
        out["ptr"] = None
        out["dim_size"] = None
        return out

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        red = {"max": "amax", "add": "sum", "mean": "mean"}[self.aggr]
        out = inputs.new_zeros((dim_size,) + tuple(inputs.shape[1:]))
        idx = index.view(-1, *([1] * (inputs.dim() - 1))).expand_as(inputs)
        return out.scatter_reduce_(0, idx, inputs, red, include_self=False)


def install() -> None:
    """Idempotently register the stand-ins and put the reference on ``sys.path``."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "torch_geometric.nn.conv" in sys.modules and getattr(sys.modules["torch_geometric.nn.conv"], "_vlsat_shim", False):
        return
    for p in (os.path.join(REF_ROOT, "src"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("tkinter", N="n")
    lib = mod("src.lib"); lib.__path__ = []
    pn = mod("src.lib.pointnet"); pn.__path__ = []
    mod("src.lib.pointnet.graph", GraphTripleConvNet=object)
    clip_inner = mod("clip.clip", load=None, tokenize=None)
    c = mod("clip", load=None, tokenize=None, clip=clip_inner); c.__path__ = []
    tg = mod("torch_geometric"); tg.__path__ = []
    tgn = mod("torch_geometric.nn"); tgn.__path__ = []
    mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing, _vlsat_shim=True)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self


def text_features(num_obj: int = 160, num_rel: int = 26, dim: int = 512, seed: int = 1234):
    """Seeded stand-in for the CLIP text features of ``get_label_weight`` (SGFN_MMG/model.py:189-219)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(num_obj, dim, generator=g)
    r = torch.randn(num_rel, dim, generator=g)
    return o / o.norm(dim=-1, keepdim=True), r / r.norm(dim=-1, keepdim=True)


def build_reference_mmgnet(seed: int = 0, overrides: dict | None = None, num_obj: int = 160, num_rel: int = 26):
    """Construct the reference ``Mmgnet`` (eval mode, CPU) from config/mmgnet.json (Appendix C recipe)."""
    install()
    from src.utils.config import Config
    from src.model.SGFN_MMG.model import Mmgnet

    cfg = Config(os.path.join(REF_ROOT, "config", "mmgnet.json"))
    cfg.PATH = tempfile.mkdtemp(prefix="vlsat_ref_")
    cfg.exp = "oracle"
    cfg.MODE = "eval"
    cfg.max_iteration = 1000
    cfg.MODEL.adapter_path = os.path.join(REF_ROOT, "clip_adapter", "checkpoint", "origin_mean.pth")
    for k, v in (overrides or {}).items():
        cfg.MODEL[k] = v

    def _labels(self, obj_label_path, rel_label_path):
        self.obj_label_list = [f"obj{i}" for i in range(num_obj)]
        self.rel_label_list = [f"rel{i}" for i in range(num_rel)]
        return text_features(num_obj, num_rel)

    Mmgnet.get_label_weight = _labels
    torch.manual_seed(seed)
    net = Mmgnet(cfg, num_obj, num_rel).eval()
    return net, cfg
