"""Seeded inputs of the golden cases. Shared by ``oracle/make_golden.py`` (which feeds them to the
UNMODIFIED reference modules in the build container and stores the outputs next to this file) and by the
tests (which feed them to the oracle and to the CUDA path). Nothing here depends on the reference."""
from __future__ import annotations

import torch

import vlsat_b200  # noqa: F401  (registers the package alias)
from vlsat_b200 import synth

GOLDEN_DIR = __file__.rsplit("/", 1)[0]

# ---- full-model cases: name -> (model-config overrides, batch builder) -------------------------------
MMGNET_CASES = {
    "mmgnet_cfg1": (dict(), lambda: synth.make_config_batch("cfg1", seed=0)),
    "mmgnet_ragged": (dict(), lambda: synth.make_batch(3, [4, 7, 2], 64, None, seed=3, shuffle_edges=True)),
    "mmgnet_l3h4": (dict(N_LAYERS=3, NUM_HEADS=4), lambda: synth.make_batch(2, [5, 6], 32, None, seed=5)),
    "mmgnet_cfg2x2": (dict(), lambda: synth.make_config_batch("cfg2", seed=1, num_scenes=2)),
    "mmgnet_addaggr": (dict(GCN_AGGR="add", N_LAYERS=1), lambda: synth.make_batch(2, [3, 5], 40, 6, seed=9)),
}
MMGNET_WEIGHT_SEED = 0


def model_config(overrides: dict) -> dict:
    cfg = dict(vlsat_b200.DEFAULT_MODEL_CONFIG)
    cfg.update(overrides)
    return {"MODEL": cfg}


# ---- graph-attention layer cases ---------------------------------------------------------------------
def gat_graph(seed: int, n_nodes: int, n_edges: int, isolated=()):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n_nodes, (n_edges,), generator=g)
    dst = torch.randint(0, n_nodes, (n_edges,), generator=g)
    for iso in isolated:                      # make sure some nodes have no outgoing / incoming edge at all
        src[src == iso] = (iso + 1) % n_nodes
        dst[dst == iso] = (iso + 1) % n_nodes
    return torch.stack([src, dst], 0)


GAT_CASES = {
    # name: (ctor kwargs, n_nodes, n_edges, isolated nodes, seed)
    "gat_max_h4": (dict(num_heads=4, dim_node=64, dim_edge=32, dim_atten=32, aggr="max", DROP_OUT_ATTEN=0.5), 9, 40, (2, 5), 11),
    "gat_add_h4": (dict(num_heads=4, dim_node=64, dim_edge=32, dim_atten=32, aggr="add", DROP_OUT_ATTEN=0.5), 9, 40, (2,), 12),
    "gat_mean_h2": (dict(num_heads=2, dim_node=32, dim_edge=64, dim_atten=128, aggr="mean", DROP_OUT_ATTEN=0.5), 7, 33, (), 13),
    "gat_noedge": (dict(num_heads=4, dim_node=64, dim_edge=32, dim_atten=32, aggr="max", use_edge=False, DROP_OUT_ATTEN=0.5), 6, 20, (0,), 14),
    "gat_s2t": (dict(num_heads=4, dim_node=64, dim_edge=32, dim_atten=32, aggr="max", flow="source_to_target", DROP_OUT_ATTEN=0.5), 8, 30, (3,), 15),
    "gat_mmg_dims": (dict(num_heads=8, dim_node=512, dim_edge=512, dim_atten=256, aggr="max", DROP_OUT_ATTEN=0.5), 12, 50, (7,), 16),
    "gat_nodrop": (dict(num_heads=4, dim_node=64, dim_edge=32, dim_atten=32, aggr="max"), 5, 12, (), 17),
}


def gat_inputs(name: str):
    kw, n, e, iso, seed = GAT_CASES[name]
    g = torch.Generator().manual_seed(seed + 100)
    x = torch.randn(n, kw["dim_node"], generator=g)
    ef = torch.randn(e, kw["dim_edge"], generator=g)
    return x, ef, gat_graph(seed, n, e, iso)


# ---- SGFN twin ---------------------------------------------------------------------------------------
GNN_CASE = dict(dim_node=512, dim_edge=256, dim_atten=256, num_layers=2, num_heads=8, aggr="max", DROP_OUT_ATTEN=0.5)


def gnn_inputs():
    b = synth.make_batch(3, [5, 3, 6], 16, None, seed=21)
    g = torch.Generator().manual_seed(22)
    n, e = b.descriptor.shape[0], b.edge_indices.shape[1]
    return (torch.randn(n, 512, generator=g), torch.randn(e, 256, generator=g), b.edge_indices,
            b.descriptor[:, :3].contiguous(), b.batch_ids)


# ---- encoders / attention ----------------------------------------------------------------------------
POINTNET_CASES = {
    "pointnet_obj": (dict(point_size=3, out_size=768), 5, 100, 31),
    "pointnet_big": (dict(point_size=3, out_size=768), 3, 257, 32),
    "pointnet_rel": (dict(point_size=11, out_size=512), 37, 1, 33),
    "pointnet_rgbn": (dict(point_size=9, out_size=256), 4, 64, 34),
}


def pointnet_inputs(name: str):
    kw, n, p, seed = POINTNET_CASES[name]
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, kw["point_size"], p, generator=g)


MHA_CASES = {"mha_h8": (512, 8, 70, 130, 41), "mha_h4": (512, 4, 33, 65, 42), "mha_self": (512, 8, 64, 64, 43)}


def mha_inputs(name: str):
    d, h, nq, nk, seed = MHA_CASES[name]
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(nq, d, generator=g)
    kv = q if name == "mha_self" else torch.randn(nk, d, generator=g)
    return q, kv


def seeded_state(module: torch.nn.Module, seed: int):
    return synth.make_state_dict({k: v.shape for k, v in module.state_dict().items()}, seed)


# ---- gradient cases (oracle/make_golden_grads.py) ----------------------------------------------------
GRAD_MMGNET_CASES = ["mmgnet_cfg1", "mmgnet_ragged", "mmgnet_addaggr"]
GRAD_GAT_CASES = ["gat_max_h4", "gat_add_h4", "gat_mean_h2", "gat_noedge", "gat_s2t", "gat_mmg_dims"]
GRAD_FULL_LIMIT = 8192        # gradients up to this many elements are stored whole
GRAD_HEAD = 512               # larger ones: the first GRAD_HEAD entries + sum + L2 norm


def loss_weights(outs, seed: int):
    """Fixed pseudo-random cotangents: the scalar loss of a gradient case is sum_i <out_i, R_i>."""
    return [synth.seeded_tensor(f"cotangent.{i}", tuple(o.shape), seed) for i, o in enumerate(outs)]


_cotangents = {}


def scalar_loss(outs, seed: int):
    """Cotangents are cached per (seed, shapes, device): a second call does no host->device copy, so the loss can be
    recorded into a CUDA graph after one eager warm-up call."""
    key = (seed, tuple(tuple(o.shape) for o in outs), str(outs[0].device))
    ws = _cotangents.get(key)
    if ws is None:
        ws = _cotangents[key] = [w.to(o.device, o.dtype) for o, w in zip(outs, loss_weights(outs, seed))]
    return sum((o * w).sum() for o, w in zip(outs, ws))


def grad_summary(g: torch.Tensor) -> dict:
    g = g.detach().double().cpu().reshape(-1)
    if g.numel() <= GRAD_FULL_LIMIT:
        return dict(full=g.float())
    return dict(head=g[:GRAD_HEAD].float(), sum=float(g.sum()), norm=float(g.norm()), absmax=float(g.abs().max()))


# ---- training-step cases (oracle/make_golden_train.py): N1, SURVEY 8f -------------------------------------------
TRAIN_CASES = ["mmgnet_cfg1", "mmgnet_ragged"]
TRAIN_STEPS = 2


def train_targets(batch, seed: int = 11, num_obj: int = 160, num_rel: int = 26):
    """Seeded supervision of a batch: object classes [N] int64, multi-label relationship targets [E, num_rel] of 0/1
    floats (about one label per edge, some edges with none), and a unit-norm [E, 512] stand-in for the CLIP text
    embedding that ``get_rel_emb`` (SGFN_MMG/model.py:221-255) would return."""
    n, e = batch.obj_points.shape[0], batch.edge_indices.shape[1]
    g = torch.Generator().manual_seed(seed)
    gt_cls = torch.randint(0, num_obj, (n,), generator=g)
    gt_rel = (torch.rand(e, num_rel, generator=g) < 1.0 / num_rel).float()
    text = torch.randn(e, 512, generator=g)
    return gt_cls, gt_rel, text / text.norm(dim=-1, keepdim=True)


# ---- object preparation (N2): seeded scan clouds + sampled indices ------------------------------------
PREP_CASES = {           # name -> (cloud rows, channels, objects, points per object, seed)
    "prep_xyz": (5000, 3, 12, 128, 21),
    "prep_rgbn": (3000, 9, 5, 256, 22),
    "prep_ragged_p": (700, 6, 3, 77, 23),
    "prep_one_point_pool": (40, 3, 2, 64, 24),
}


def prep_inputs(name: str):
    """A scan-like cloud (objects are blobs metres away from the origin) and, per object, indices drawn with
    replacement from that object's own points (np.random.choice(len(obj_pointset), num_points, replace=True))."""
    m, c, n, p, seed = PREP_CASES[name]
    g = torch.Generator().manual_seed(seed)
    owner = torch.randint(0, n, (m,), generator=g)
    owner[:n] = torch.arange(n)                                   # every object owns at least one point
    centre = (torch.rand(n, 3, generator=g) - 0.5) * 12.0
    scale = torch.rand(n, 3, generator=g) * 0.9 + 0.1
    cloud = torch.randn(m, c, generator=g)
    cloud[:, :3] = cloud[:, :3] * scale[owner] + centre[owner]
    choice = torch.empty(n, p, dtype=torch.int64)
    for o in range(n):
        pool = torch.nonzero(owner == o).view(-1)
        if name == "prep_one_point_pool" and o == 0:
            pool = pool[:1]                                       # an instance with a single point: std 0, extent 0
        choice[o] = pool[torch.randint(0, pool.numel(), (p,), generator=g)]
    return cloud, choice


# ---- evaluation ranks (N3): seeded predictions and targets -------------------------------------------
EVAL_CASES = {           # name -> (objects, edges, seed)
    "eval_small": (12, 40, 31),
    "eval_scene": (40, 300, 32),
}


def eval_inputs(name: str, num_obj: int = 160, num_rel: int = 26):
    """Object logits [N, 160], relationship probabilities [E, 26] (with tied rows, a confident and an unconfident
    label-free edge), targets and [E, 2] (subject, object) edges as ``process_val`` holds them."""
    n, e, seed = EVAL_CASES[name]
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, num_obj, generator=g) * 3
    rel = torch.sigmoid(torch.randn(e, num_rel, generator=g) * 2)
    gt_cls = torch.randint(0, num_obj, (n,), generator=g)
    gt_rel = (torch.rand(e, num_rel, generator=g) < 0.06).float()
    edges = torch.randint(0, n, (e, 2), generator=g)
    rel[3], rel[4] = 0.7, 0.2                                     # ties everywhere
    gt_rel[3], gt_rel[4] = 0, 0                                   # ... on label-free edges: all above / all below the threshold
    gt_rel[5] = 0
    gt_rel[5, [2, 9, 17]] = 1                                     # three labels on one edge: the sorted-minus-counter rule
    logits[0, gt_cls[0]] = logits[0].max() + 1                   # rank 1
    logits[1, gt_cls[1]] = logits[1].min() - 1                   # beyond top-k
    return logits, rel, gt_cls, gt_rel, edges


# ------------------------------------------------------------------------------------------------------------------
# N4: CLIP text supervision. name -> (object classes, predicate classes, nodes, edges, seed)
TEXT_CASES = {"text_small": (7, 5, 12, 40, 3), "text_mmgnet_vocab": (160, 26, 40, 600, 4)}


def text_table(n_obj_cls: int, n_rel_cls: int, seed: int, dim: int = 512) -> torch.Tensor:
    """Seeded stand-in for the CLIP text features of every prompt get_rel_emb can build: [S, O, R + 1, dim], slot R = the
    "no relation" prompt (SGFN_MMG/model.py:232-240)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(n_obj_cls, n_obj_cls, n_rel_cls + 1, dim, generator=g)


def text_inputs(name: str):
    """(gt_cls [N] int64, gt_rel_cls [E, R] float 0/1 with label-free, single- and multi-label edges, edges [E, 2] int64)."""
    n_obj_cls, n_rel_cls, n_nodes, n_edges, seed = TEXT_CASES[name]
    g = torch.Generator().manual_seed(2000 + seed)
    gt_cls = torch.randint(0, n_obj_cls, (n_nodes,), generator=g)
    gt_rel = (torch.rand(n_edges, n_rel_cls, generator=g) < 1.5 / n_rel_cls).float()
    gt_rel[::5] = 0                                              # label-free edges
    edges = torch.randint(0, n_nodes, (n_edges, 2), generator=g)
    return gt_cls, gt_rel, edges
