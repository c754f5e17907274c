"""Backward / training-mode parity on the GPU, all through the C ABI.

  * every backward primitive against a float64 torch restatement of the same formula;
  * module gradients (full Mmgnet in eval and train mode, GAT layers incl. add / mean / no-edge / source_to_target,
    PointNet, attention) against gradient fixtures produced by the UNMODIFIED reference
    (oracle/make_golden_grads.py -> tests/golden/grads.pt) and against the oracle's autograd on larger inputs;
  * the differentiable forward against the same forward fixtures as the inference path;
  * dropout statistics / determinism, BatchNorm running-stat updates.

Gradient tolerance (conftest.py): rtol 1e-3 + 2e-4 x max|reference gradient of that tensor| (+ 1e-6 x the largest
gradient of the case for gradients that are zero in exact arithmetic)."""
import math

import pytest
import torch
import torch.nn.functional as F

import cases
import vlsat_b200 as V
from conftest import FEATURE_ATOL_SCALE, assert_close, assert_grad_summary_close, grad_floor
from oracle import vlsat_oracle as O
from vlsat_b200 import autograd as A
from vlsat_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def close64(actual, expected, what, rtol=1e-4, atol_scale=2e-5):
    assert_close(actual, expected.float(), what, rtol=rtol, atol=0.0, atol_scale=atol_scale)


# ---------------------------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("shape", [(1, 1), (5, 7), (33, 64), (1000, 130), (9600, 64)])
def test_transpose_zero_pads(shape):
    x = rnd(*shape, seed=1)
    t = ops.transpose(x)
    assert t.shape == (shape[1], (shape[0] + 3) // 4 * 4)
    assert torch.equal(t[:, :shape[0]], x.t())
    assert torch.count_nonzero(t[:, shape[0]:]) == 0
    wide = rnd(shape[0], shape[1] + 8, seed=2)
    assert torch.equal(ops.transpose(wide[:, 4:4 + shape[1]])[:, :shape[0]], wide[:, 4:4 + shape[1]].t())
    xb = rnd(3, shape[0] % 50 + 1, shape[1], seed=3)
    tb = ops.transpose(xb)
    assert torch.equal(tb[:, :, :xb.shape[1]], xb.transpose(1, 2))


@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_RELU, ops.ACT_SIGMOID])
def test_act_bwd(act):
    m, n = 777, 70
    dy, pre = rnd(m, n, seed=4), rnd(m, n, seed=5)
    y = pre if act == ops.ACT_NONE else torch.relu(pre) if act == ops.ACT_RELU else torch.sigmoid(pre)
    s = torch.tensor([0.3], device=DEV)
    dz, db = ops.act_bwd(dy, y, act, scale_ptr=s)
    d = dy.double() * math.exp(0.3)
    want = d if act == ops.ACT_NONE else d * (y > 0) if act == ops.ACT_RELU else d * (y.double() * (1 - y.double()))
    close64(dz, want, "dz")
    close64(db, want.sum(0), "dbias")


def test_scatter_and_gather_rows():
    x = rnd(500, 48, seed=6)
    idx = torch.randint(0, 37, (500,), generator=torch.Generator().manual_seed(7)).to(DEV)
    out = ops.scatter_add_rows(x, idx, torch.zeros(37, 48, device=DEV))
    want = torch.zeros(37, 48, device=DEV, dtype=torch.float64).index_add_(0, idx, x.double())
    close64(out, want, "scatter_add_rows")
    src = rnd(37, 48, seed=8)
    assert torch.equal(ops.gather_rows(src, idx), src[idx])
    # rows (e, h) -> rows (node, h)
    H = 4
    xe = rnd(500 * H, 16, seed=9)
    out = ops.scatter_add_rows(xe, idx, torch.zeros(37 * H, 16, device=DEV), rows_per_idx=H)
    rows = (idx.view(-1, 1) * H + torch.arange(H, device=DEV)).reshape(-1)
    want = torch.zeros(37 * H, 16, device=DEV, dtype=torch.float64).index_add_(0, rows, xe.double())
    close64(out, want, "scatter_add_rows heads")
    assert torch.equal(ops.gather_rows(out, idx, rows_per_idx=H), out[rows])


@pytest.mark.parametrize("m,d,relu,res", [(1, 32, False, False), (300, 512, False, True), (77, 32, False, False), (1000, 512, True, True)])
def test_layernorm_bwd(m, d, relu, res):
    x, r, dy = rnd(m, d, seed=10), (rnd(m, d, seed=11) if res else None), rnd(m, d, seed=12)
    gamma, beta = rnd(d, seed=13) * 0.5 + 1.0, rnd(d, seed=14) * 0.1
    dx, dg, db = ops.add_layernorm_bwd(dy, x, r, gamma, beta, 1e-5, relu)
    x64 = (x.double() + (r.double() if res else 0)).requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y = F.layer_norm(x64, (d,), g64, b64, 1e-5)
    if relu:
        y = torch.relu(y)
    y.backward(dy.double())
    close64(dx, x64.grad, "dx")
    close64(dg, g64.grad, "dgamma")
    close64(db, b64.grad, "dbeta")


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (1000, 128, 64), (76800, 32, 128), (5000, 64, 3), (333, 100, 11)])
def test_wgrad_small(m, n, k):
    dz, x = rnd(m, n, seed=15), rnd(m, k, seed=16)
    close64(ops.wgrad_small(dz, x), dz.double().t() @ x.double(), "wgrad_small")


@pytest.mark.parametrize("batch_stats,relu", [(True, True), (False, True), (True, False)])
def test_batchnorm_fwd_bwd(batch_stats, relu):
    m, n = 640, 504
    x, dy = rnd(m, n, seed=17) * 2 + 0.3, rnd(m, n, seed=18)
    bn = torch.nn.BatchNorm1d(n).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(n, seed=19) * 0.3 + 1); bn.bias.copy_(rnd(n, seed=20) * 0.2)
        bn.running_mean.copy_(rnd(n, seed=21) * 0.1); bn.running_var.copy_(rnd(n, seed=22).abs() + 0.5)
    ref = torch.nn.BatchNorm1d(n).to(DEV).double()
    ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    ref.train(batch_stats)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr = torch.relu(yr) if relu else yr
    yr.backward(dy.double())
    xg = x.clone().requires_grad_(True)
    y = A.batchnorm(xg, bn, batch_stats, relu_out=relu)
    y.backward(dy)
    close64(y, yr, "bn y")
    close64(xg.grad, xr.grad, "bn dx")
    close64(bn.weight.grad, ref.weight.grad, "bn dgamma")
    close64(bn.bias.grad, ref.bias.grad, "bn dbeta")
    close64(bn.running_mean, ref.running_mean, "running_mean")
    close64(bn.running_var, ref.running_var, "running_var")
    assert int(bn.num_batches_tracked) == int(ref.num_batches_tracked)


def test_row_l2norm_bwd_and_dot():
    x, dy = rnd(100, 512, seed=23), rnd(100, 512, seed=24)
    x64 = x.double().requires_grad_(True)
    (x64 / x64.norm(dim=-1, keepdim=True)).backward(dy.double())
    close64(ops.row_l2norm_bwd(dy, x), x64.grad, "l2norm dx")
    out = torch.zeros(1, device=DEV)
    ops.dot_accum(x, dy, out)
    close64(out, (x.double() * dy.double()).sum().reshape(1), "dot")


def test_dropout_mask_statistics_and_backward():
    A.DropoutState.manual_seed(123)
    x = torch.ones(4000, 256, device=DEV, requires_grad=True)
    y = A.dropout(x, 0.3, True)
    kept = (y != 0)
    assert abs(kept.float().mean().item() - 0.7) < 0.005
    assert torch.allclose(y[kept], torch.full_like(y[kept], 1 / 0.7))
    y.backward(torch.ones_like(y))
    assert torch.equal(x.grad != 0, kept)                       # same mask in the backward
    y2 = A.dropout(x, 0.3, True)
    assert not torch.equal(y2 != 0, kept)                       # the stream advances
    A.DropoutState.manual_seed(123)
    assert torch.equal(A.dropout(x, 0.3, True), y)              # and is reproducible
    assert A.dropout(x, 0.3, False) is x and A.dropout(x, 0.0, True) is x

def _pair_sum(pair):
    return pair[0].float() + pair[1].float()


def test_pairs_emitted_by_dropout_permutation_views_and_small_k_projection():
    """Producers that hand the next projection its bf16 (hi, lo) pair: the pair must be the split of exactly the fp32 result
    (hi + lo within 2^-16 relative), remembered on the returned tensor, and survive the (rows, head) view."""
    A.DropoutState.manual_seed(7)
    x = rnd(1000, 128, seed=160)
    y = A.dropout(x, 0.4, True, emit_pair=True)
    A.DropoutState.manual_seed(7)
    assert torch.equal(A.dropout(x, 0.4, True), y)                          # same mask and values as the plain pass
    pair = ops.act_pair(y)
    assert pair is y._vlsat_pair[1], "dropout must have remembered the pair"
    assert torch.allclose(_pair_sum(pair), y, rtol=2e-5, atol=0.0)
    # permutation (scatter and gather forms)
    perm = torch.randperm(1000, generator=torch.Generator().manual_seed(3)).to(DEV, torch.int32)
    for gather in (True, False):
        z = A.permute_rows(x, perm, gather, emit_pair=True)
        assert torch.equal(z, A.permute_rows(x, perm, gather))
        assert torch.allclose(_pair_sum(z._vlsat_pair[1]), z, rtol=2e-5, atol=0.0)
    # (edge, head) rows as a view of a head-major projection
    w = rnd(512, 128, seed=161) / 11.0
    k_hm = A.linear(x, w, None, emit_pair=True)                             # [1000, 8 * 64]
    rows = ops.view_rows(k_hm, 8000, 64)
    assert rows.data_ptr() == k_hm.data_ptr() and tuple(rows.shape) == (8000, 64)
    assert torch.allclose(_pair_sum(ops.act_pair(rows)), rows, rtol=2e-5, atol=0.0)
    assert ops.act_pair(rows)[0].data_ptr() == k_hm._vlsat_pair[1][0].data_ptr()
    plain = ops.view_rows(rnd(1000, 128, seed=165), 2000, 64)               # no pair remembered: a plain view
    assert getattr(plain, "_vlsat_pair", None) is None
    # K = 3 register-row kernel (PointNet's first layer in the backward recompute)
    pts, w1, b1 = rnd(5000, 3, seed=162), rnd(64, 3, seed=163), rnd(64, seed=164)
    h, hp = ops.linear(pts, w1, b1, act=ops.ACT_RELU, emit_split="bf16")
    want = torch.relu(pts.double() @ w1.double().t() + b1.double()).float()
    close64(h, want.double(), "small-K projection", rtol=1e-5, atol_scale=1e-6)
    assert torch.allclose(_pair_sum(hp), h, rtol=2e-5, atol=0.0)


@pytest.mark.parametrize("aggr", ["max", "add", "mean"])
def test_gat_softmax_aggr_fwd_bwd(aggr):
    n, e, H, do = 30, 200, 4, 32
    ei = cases.gat_graph(5, n, e, isolated=(3,)).to(DEV)
    g = V.GraphContext(ei, n)
    t = rnd(e * H, do, seed=25).requires_grad_(True)
    v = rnd(n, H * do, seed=26).requires_grad_(True)
    xx, prob = A.gat_softmax_aggr(t, v, g, H, aggr)
    dxx = rnd(n, H * do, seed=27)
    xx.backward(dxx)
    t64, v64 = t.detach().double().requires_grad_(True), v.detach().double().requires_grad_(True)
    p = torch.softmax(t64.view(e, H, do), -1)                              # [(e,h), c]
    msg_hm = p * v64.view(n, H, do)[g.dst]                                  # [e, h, c]
    msg = msg_hm.permute(0, 2, 1).reshape(e, do * H)                        # interleaved c*H + h
    want = O.aggregate(msg, g.edge_index, n, aggr)
    want.backward(dxx.double())
    close64(xx, want, "xx")
    close64(prob, p.reshape(e * H, do), "prob")
    close64(t.grad, t64.grad, "dt")
    close64(v.grad, v64.grad, "dv")


def _dense_node_attn(q, k, v, bias_dense, same, H):
    n, d = q.shape
    dk = d // H
    s = torch.einsum("ahd,bhd->hab", q.view(n, H, dk), k.view(n, H, dk)) / math.sqrt(dk) + bias_dense
    s = s.masked_fill(~same.unsqueeze(0), -math.inf)
    return torch.einsum("hab,bhd->ahd", torch.softmax(s, -1), v.view(n, H, dk)).reshape(n, d)


@pytest.mark.parametrize("sizes,H", [([1, 40, 3], 8), ([150], 8), ([7, 7], 4)])
def test_node_attn_bias_fwd_bwd(sizes, H):
    from vlsat_b200 import train_path as T
    n, d = sum(sizes), 512
    bid = torch.cat([torch.full((s,), i) for i, s in enumerate(sizes)]).view(-1, 1).to(DEV)
    centres = rnd(n, 3, seed=28)
    sctx = T.SceneContextTrain(bid, centres)
    assert sctx.n_pairs == sum(s * s for s in sizes) and sctx.max_scene == max(sizes)
    q, k, v = (rnd(n, d, seed=29 + i).requires_grad_(True) for i in range(3))
    bias = rnd(sctx.n_pairs, H, seed=33).requires_grad_(True)
    out = A.node_attn(q, k, v, bias, sctx, H)
    dout = rnd(n, d, seed=34)
    out.backward(dout)
    # dense float64 restatement
    same = (bid.view(-1, 1) == bid.view(1, -1))
    q64, k64, v64, b64 = (t.detach().double().requires_grad_(True) for t in (q, k, v, bias))
    dense = torch.zeros(H, n, n, device=DEV, dtype=torch.float64)
    a_idx, b_idx = torch.nonzero(same, as_tuple=True)                      # row-major = pair order
    dense[:, a_idx, b_idx] = b64.t()
    want = _dense_node_attn(q64, k64, v64, dense, same, H)
    want.backward(dout.double())
    close64(out, want, "node attention")
    for name, a, b in (("dq", q, q64), ("dk", k, k64), ("dv", v, v64), ("dbias", bias, b64)):
        close64(a.grad, b.grad, name)
    # pair features: [c_b - c_a, |c_b - c_a|]
    diff = centres[b_idx] - centres[a_idx]
    assert_close(sctx.pair_feats, torch.cat([diff, diff.norm(dim=-1, keepdim=True)], 1), "pair features", atol=1e-6)


@pytest.mark.parametrize("streaming", [True, False])
@pytest.mark.parametrize("nq,nk,H", [(100, 257, 8), (1, 1, 8), (300, 129, 8), (50, 70, 4), (64, 128, 8), (1000, 2100, 8)])
def test_flash_attention_backward(nq, nk, H, streaming, monkeypatch):
    """A9 gradients against float64 autograd of attention.py:41-78, through the streaming tcgen05 backward
    (csrc/flash_attn_bwd.cu; H * 64 = 512 only) and through the round-1 block path that the other engines keep."""
    monkeypatch.setattr(A, "FLASH_BWD_QUERY_BLOCK", 128)                   # exercise the accumulation over query blocks
    monkeypatch.setattr(A, "FLASH_BWD_STREAMING", streaming)
    d = 512
    q, k, v = (rnd(n_, d, seed=40 + i).requires_grad_(True) for i, n_ in enumerate((nq, nk, nk)))
    out = A.flash_attn(q, k, v, H)
    dout = rnd(nq, d, seed=44)
    out.backward(dout)
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    dk = d // H
    s = torch.einsum("ahd,bhd->hab", q64.view(nq, H, dk), k64.view(nk, H, dk)) / math.sqrt(dk)
    want = torch.einsum("hab,bhd->ahd", torch.softmax(s, -1), v64.view(nk, H, dk)).reshape(nq, d)
    want.backward(dout.double())
    close64(out, want, "attention out", rtol=1e-3, atol_scale=1e-4)
    # one key: dq and dk vanish in exact arithmetic, so give them an absolute floor on the scale of the inputs
    assert_close(q.grad, q64.grad.float(), "dq", rtol=1e-3, atol=1e-4, atol_scale=2e-4)
    assert_close(k.grad, k64.grad.float(), "dk", rtol=1e-3, atol=1e-4, atol_scale=2e-4)
    close64(v.grad, v64.grad, "dv", rtol=1e-3, atol_scale=2e-4)


def test_streaming_flash_backward_at_config2_size_and_forced_splits(monkeypatch):
    """nq = 4800, nk = 9600 (config #2's key count), 8 heads, against float64; then the same call with the streamed
    range split over 3 CTAs per tile (slab sums) must reproduce itself to rounding."""
    nq, nk, H, d = 4800, 9600, 8, 512
    q, k, v = (rnd(n_, d, seed=140 + i).requires_grad_(True) for i, n_ in enumerate((nq, nk, nk)))
    dout = rnd(nq, d, seed=144)
    A.flash_attn(q, k, v, H).backward(dout)
    q64, k64, v64 = (t.detach().double().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("ahd,bhd->hab", q64.view(nq, H, 64), k64.view(nk, H, 64)) / 8.0
    torch.einsum("hab,bhd->ahd", torch.softmax(s, -1), v64.view(nk, H, 64)).reshape(nq, d).backward(dout.double())
    del s
    for name, a, b in (("dq", q, q64), ("dk", k, k64), ("dv", v, v64)):
        close64(a.grad, b.grad, name, rtol=1e-3, atol_scale=2e-4)
    base = [t.grad.clone() for t in (q, k, v)]
    for splits in ("1", "3"):
        monkeypatch.setenv("VLSAT_FLASH_BWD_SPLITS", splits)
        for t in (q, k, v):
            t.grad = None
        A.flash_attn(q, k, v, H).backward(dout)
        for name, t, b in zip(("dq", "dk", "dv"), (q, k, v), base):
            assert_close(t.grad, b, f"{name} splits={splits}", rtol=1e-4, atol_scale=1e-5)


@pytest.mark.parametrize("m,n,k", [(9600, 512, 1024), (640, 768, 512), (300, 64, 200), (9600, 1536, 1024), (77, 8, 40), (1000, 160, 512)])
def test_gemm_nn_on_stored_pairs(m, n, k):
    """dX = dZ W with W read as the forward stores it (MN-major B operand, no transposed copy) against float64."""
    a, b = rnd(m, k, seed=150), rnd(k, n, seed=151) / math.sqrt(k)
    got = ops.gemm_nn(ops.bf16_split(a), ops.bf16_split(b), n)
    close64(got, a.double() @ b.double(), "gemm_nn", rtol=1e-3, atol_scale=1e-4)


@pytest.mark.parametrize("k,m,n", [(9600, 1024, 1536), (76800, 128, 128), (76800, 32, 128), (640, 768, 768), (9600, 512, 512), (100, 504, 768),
                                   (1000, 160, 512), (70, 8, 64)])
def test_gemm_tn_on_stored_pairs(k, m, n):
    """dW = dZ^T X with both operands read as stored (MN-major A and B, reduction split over CTAs) against float64."""
    a, b = rnd(k, m, seed=152), rnd(k, n, seed=153)
    got = ops.gemm_tn(ops.bf16_split(a), ops.bf16_split(b), m, n)
    close64(got, a.double().t() @ b.double(), "gemm_tn", rtol=1e-3, atol_scale=1e-4)
    assert torch.equal(got, ops.gemm_tn(ops.bf16_split(a), ops.bf16_split(b), m, n)), "split reduction must be deterministic"


@pytest.mark.parametrize("m,n,k,act", [(300, 504, 768, ops.ACT_NONE), (9600, 64, 11, ops.ACT_RELU), (30, 26, 256, ops.ACT_SIGMOID),
                                       (1000, 160, 512, ops.ACT_NONE), (0, 64, 64, ops.ACT_RELU)])
def test_linear_backward(m, n, k, act):
    x, w, b = rnd(m, k, seed=50).requires_grad_(True), (rnd(n, k, seed=51) / math.sqrt(k)).requires_grad_(True), rnd(n, seed=52).requires_grad_(True)
    y = A.linear(x, w, b, act)
    dy = rnd(m, n, seed=53)
    y.backward(dy)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    z = x64 @ w64.t() + b64
    z = z if act == ops.ACT_NONE else torch.relu(z) if act == ops.ACT_RELU else torch.sigmoid(z)
    z.backward(dy.double())
    for name, a, r in (("dx", x, x64), ("dw", w, w64), ("db", b, b64)):
        close64(a.grad, r.grad, f"linear {name}", rtol=1e-3, atol_scale=1e-4)


def test_linear_backward_gather_residual_scale():
    m, n, k, rows = 700, 128, 64, 50
    g_ = torch.Generator().manual_seed(54)
    ia, ib = torch.randint(0, rows, (m,), generator=g_).to(DEV), torch.randint(0, rows, (m,), generator=g_).to(DEV)
    x, w = rnd(m, k, seed=55).requires_grad_(True), (rnd(n, k, seed=56) / 8).requires_grad_(True)
    ga, gb = rnd(rows, n, seed=57).requires_grad_(True), rnd(rows, n, seed=58).requires_grad_(True)
    res = rnd(m, n, seed=59).requires_grad_(True)
    s = torch.tensor([0.7], device=DEV, requires_grad=True)
    dy = rnd(m, n, seed=60)
    y1 = A.linear(x, w, None, ops.ACT_RELU, gather=(ga, ia, gb, ib))
    y2 = A.linear(x, w, None, ops.ACT_NONE, residual=res)
    y3 = A.linear(x, w, None, ops.ACT_NONE, scale=s)
    y4 = A.linear(x, w, None, ops.ACT_RELU, gather=(ga, ia, None, None))
    (y1 + y2 + y3 + y4).backward(dy)
    t64 = [t.detach().double().requires_grad_(True) for t in (x, w, ga, gb, res, s)]
    x6, w6, ga6, gb6, r6, s6 = t64
    z = x6 @ w6.t()
    # The ReLU gates are the ones the CUDA forward took: a pre-activation within the forward error of zero (1e-5 relative
    # on the BF16x3 engine) may land on either side, and one flipped gate moves a whole row of dx by O(|dy| |w|).
    p1, p4 = z + ga6[ia] + gb6[ib], z + ga6[ia]
    gate1, gate4 = (y1.detach() > 0).double(), (y4.detach() > 0).double()
    for gate, pre in ((gate1, p1), (gate4, p4)):
        flipped = gate != (pre.detach() > 0).double()
        assert flipped.sum() <= 8 and (pre.detach().abs()[flipped] < 1e-3).all()      # only values that are ~0 may flip
    want = gate1 * p1 + (z + r6) + z * s6.exp() + gate4 * p4
    want.backward(dy.double())
    for name, a, r in zip(("dx", "dw", "dga", "dgb", "dres", "dscale"), (x, w, ga, gb, res, s), t64):
        close64(a.grad, r.grad, f"linear epilogue {name}", rtol=1e-3, atol_scale=1e-4)


# ------------------------------------------------------------------------------------------------- modules
@pytest.fixture(scope="module")
def grads(golden):
    return golden("grads")


def _no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


# Arg-max routing. The path has two max reductions (PointNet max-pool over points, max aggregation over edges) whose
# backward sends a gradient entry to the winning element only. When the two best candidates are closer than the forward
# error of an engine (exact FFMA: ~1e-7 relative; tcgen05 3xTF32: ~1e-5), the winner - hence the route of that single
# gradient entry - may differ from the reference's although every forward value is within tolerance. Therefore:
#   * "strict": exact-fp32 engine, element-wise rtol 1e-3 (+ floors) against the reference fixtures;
#   * "norm":   tensor-core engines, and larger cases against the float32 oracle: ||g - g_ref|| <= bound * ||g_ref|| per
#               tensor, bound = 2e-2 for 3xTF32 ("tc", forward error ~1e-6) and 6e-2 for the default BF16x3 ("auto",
#               forward error ~1e-5: ten times as many ReLU gates / arg-max routes sit inside the noise, and every flipped
#               one moves a whole row of some gradient; the same ambiguity the fp32 reference has against exact arithmetic).
NORM_BOUND = {"tc": 2e-2, "auto": 6e-2}


@pytest.fixture(params=["simt", "tc", "auto"])
def engine(request):
    old = ops._engine
    ops.set_gemm_engine(request.param)
    yield request.param
    ops._engine = old


def _check_param_grads(module, want, what, extra=None, mode="strict", rel=2e-2):
    # gradients that vanish in exact arithmetic (softmax key bias) carry the engine's noise: 1e-5 of the largest gradient
    floor = 10 * grad_floor(want)
    params = dict(module.named_parameters())
    n, failures = 0, []
    for k, w in want.items():
        got = extra[k] if extra and k in extra else params[k].grad
        assert got is not None, f"{what}: no gradient for {k}"
        if mode == "strict":
            try:
                assert_grad_summary_close(got, w, f"{what} d{k}", floor=floor)
            except AssertionError as e:
                failures.append(str(e).splitlines()[0][:260])
        else:
            g = got.detach().double().cpu().reshape(-1)
            assert torch.isfinite(g).all(), f"{what} d{k}: non-finite"
            ref = (w["full"] if "full" in w else w["head"]).double()
            err = (g[:ref.numel()] - ref).norm().item()
            bound = rel * ref.norm().item() + floor * ref.numel() ** 0.5
            if err > bound:
                failures.append(f"{what} d{k}: ||g - ref|| = {err:.3g} > {bound:.3g} (||ref|| = {ref.norm().item():.3g})")
            if "norm" in w and abs(float(g.norm()) - w["norm"]) > rel * w["norm"] + floor * g.numel() ** 0.5:
                failures.append(f"{what} d{k}: norm {float(g.norm()):.6g} vs {w['norm']:.6g}")
        n += 1
    assert not failures, f"{len(failures)}/{n} gradients differ:\n" + "\n".join(failures)
    return n


@pytest.mark.parametrize("name", cases.GRAD_MMGNET_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_mmgnet_gradients_match_reference(name, mode, grads, engine):
    over, make = cases.MMGNET_CASES[name]
    model = V.Mmgnet(cases.model_config(over), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    model = _no_dropout(model.to(DEV)).train(mode == "train")
    b = make().to(DEV)
    outs = model(*b.forward_args(), istrain=True)
    want = grads[f"{name}.{mode}"]
    for i in (0, 1):
        assert_close(outs[i], want["outs"][i], f"{name}.{mode} output {i}", atol_scale=FEATURE_ATOL_SCALE)
    for i in (2, 3):
        assert_close(outs[i], want["outs"][i], f"{name}.{mode} output {i}")
    cases.scalar_loss(outs[:7], seed=7).backward()
    assert _check_param_grads(model, want["grads"], f"{name}.{mode}", mode="strict" if engine == "simt" else "norm",
                              rel=NORM_BOUND.get(engine, 2e-2)) >= 100
    for p in model.clip_adapter.parameters():
        assert p.grad is None                                                   # frozen (SGFN_MMG/model.py:179-182)
    if mode == "train":
        assert_close(model.mlp_3d[1].running_mean, want["bn_running_mean"], "BN running_mean", atol_scale=1e-5)
        assert_close(model.mlp_3d[1].running_var, want["bn_running_var"], "BN running_var", atol_scale=1e-5)


@pytest.mark.parametrize("name", cases.GRAD_GAT_CASES)
def test_gat_layer_gradients_match_reference(name, grads):
    kw, n, e, iso, seed = cases.GAT_CASES[name]
    layer = V.GraphEdgeAttenNetwork(**kw)
    layer.load_state_dict(cases.seeded_state(layer, seed))
    layer = layer.to(DEV).eval()
    x, ef, ei = (t.to(DEV) for t in cases.gat_inputs(name))
    x.requires_grad_(True); ef.requires_grad_(True)
    xo, eo = layer(x, ef, ei)
    cases.scalar_loss([xo, eo], seed=8).backward()
    _check_param_grads(layer, grads[name]["grads"], name, extra={"input.x": x.grad, "input.edge": ef.grad})


@pytest.mark.parametrize("name", ["pointnet_obj", "pointnet_big", "pointnet_rgbn"])
def test_pointnet_gradients_match_reference(name, grads):
    kw, n, p, seed = cases.POINTNET_CASES[name]
    enc = V.PointNetfeat(global_feat=True, batch_norm=False, input_transform=False, feature_transform=False, **kw)
    enc.load_state_dict(cases.seeded_state(enc, seed))
    enc = enc.to(DEV).eval()
    o = enc(cases.pointnet_inputs(name).to(DEV))
    cases.scalar_loss([o], seed=9).backward()
    _check_param_grads(enc, grads[name]["grads"], name)


@pytest.mark.parametrize("name", list(cases.MHA_CASES))
def test_mha_gradients_match_reference(name, grads):
    d, h, nq, nk, seed = cases.MHA_CASES[name]
    att = V.MultiHeadAttention(d_model=d, d_k=d // h, d_v=d // h, h=h)
    att.load_state_dict(cases.seeded_state(att, seed))
    att = att.to(DEV).eval()
    q, kv = cases.mha_inputs(name)
    q = q.to(DEV).requires_grad_(True)
    kv = q if name == "mha_self" else kv.to(DEV).requires_grad_(True)
    o = att(q.unsqueeze(0), kv.unsqueeze(0), kv.unsqueeze(0)).squeeze(0)
    cases.scalar_loss([o], seed=10).backward()
    extra = {"input.q": q.grad}
    if kv is not q:
        extra["input.kv"] = kv.grad
    _check_param_grads(att, grads[name]["grads"], name, extra=extra)


@pytest.mark.parametrize("name", list(cases.MMGNET_CASES))
def test_differentiable_forward_matches_reference_golden(name, golden):
    """The autograd-capable path (grad mode on, eval) reproduces the forward fixtures of the inference path."""
    over, make = cases.MMGNET_CASES[name]
    model = V.Mmgnet(cases.model_config(over), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    model = model.to(DEV).eval()
    b = make().to(DEV)
    tr = model(*b.forward_args(), istrain=True)
    assert tr[0].grad_fn is not None
    g = golden(name)
    for i in (0, 1):
        assert_close(tr[i], g["train"][i], f"{name} output {i}", atol_scale=FEATURE_ATOL_SCALE)
    for i in (2, 3):
        assert_close(tr[i], g["train"][i], f"{name} output {i}")
    for i in (4, 5, 6):
        assert_close(tr[i], g["train"][i], f"{name} output {i}", atol_scale=FEATURE_ATOL_SCALE)


def test_gnn_layers_differentiable_matches_golden(golden):
    net = V.GraphEdgeAttenNetworkLayers(**cases.GNN_CASE)
    net.load_state_dict(cases.seeded_state(net, 23))
    net = net.to(DEV).eval()
    node, edge, probs = net(*[t.to(DEV) for t in cases.gnn_inputs()])
    g = golden("gnn_layers")
    assert_close(node, g["node"], "node", atol_scale=FEATURE_ATOL_SCALE)
    assert_close(edge, g["edge"], "edge", atol_scale=FEATURE_ATOL_SCALE)
    for i, (p, w) in enumerate(zip(probs, g["probs"])):
        assert_close(p, w, f"prob layer {i}")
    (node.sum() + edge.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


def test_mmgnet_gradients_match_oracle_autograd_on_cfg2_scenes():
    """Two full-shape cfg2 scenes (80 objects x 256 points, 1200 edges): every parameter gradient against the
    oracle's autograd on the host (float32 torch CPU)."""
    over, make = cases.MMGNET_CASES["mmgnet_cfg2x2"]
    model = V.Mmgnet(cases.model_config(over), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    model = model.to(DEV).eval()
    b = make()
    outs_ref = O.mmgnet_forward(sd, *b.forward_args(), istrain=True)
    cases.scalar_loss(outs_ref[:7], seed=7).backward()
    outs = model(*b.to(DEV).forward_args(), istrain=True)
    cases.scalar_loss(outs[:7], seed=7).backward()
    n, bad = 0, []
    scale = max(sd[k].grad.norm().item() for k, p in model.named_parameters() if sd[k].grad is not None)
    for k, p in model.named_parameters():
        r = sd[k].grad
        if r is None or not p.requires_grad:
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        err = (p.grad.cpu() - r).norm().item()
        if err > 2e-2 * r.norm().item() + 1e-6 * scale:            # whole tensors: see the arg-max routing note above
            bad.append(f"{k}: ||g - ref|| = {err:.3g}, ||ref|| = {r.norm().item():.3g}")
        n += 1
    assert n >= 100 and not bad, "\n".join(bad)


def test_training_mode_with_dropout_runs_and_is_seeded():
    model = V.Mmgnet(cases.model_config({}), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    model = model.to(DEV).train()
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]().to(DEV)

    def step(seed):
        A.DropoutState.manual_seed(seed)
        model.zero_grad(set_to_none=True)
        outs = model(*b.forward_args(), istrain=True)
        cases.scalar_loss(outs[:7], seed=7).backward()
        return [o.detach().clone() for o in outs[:4]], model.mmg.gcn_3ds[0].prop[0].weight.grad.clone()
    o1, g1 = step(1)
    o2, g2 = step(1)
    o3, _ = step(2)
    assert all(torch.equal(a, b_) for a, b_ in zip(o1, o2))
    assert torch.allclose(g1, g2, rtol=1e-3, atol=1e-4 * g1.abs().max().item())   # atomics reorder sums, masks are identical
    assert not torch.equal(o1[2], o3[2])
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_eager_training_does_not_grow_the_weight_split_cache():
    """Round-1 advisor finding: the split cache keyed the head-major weight copies of the differentiable path by their
    storage address, a fresh buffer per call, and pinned them for ever (about 20 MB per eager step). Forty eager steps
    with parameter updates must leave the cache size and the allocated memory flat."""
    model = V.Mmgnet(cases.model_config({}), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    model = model.to(DEV).train()
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]().to(DEV)
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-6)
    sizes, mem = [], []
    for it in range(40):
        opt.zero_grad(set_to_none=True)
        cases.scalar_loss(model(*b.forward_args(), istrain=True)[:7], seed=7).backward()
        opt.step()
        if it in (4, 39):
            torch.cuda.synchronize()
            sizes.append(len(ops._weight_splits)); mem.append(torch.cuda.memory_allocated())
    assert sizes[1] == sizes[0], f"weight-split cache grew from {sizes[0]} to {sizes[1]} entries"
    assert mem[1] <= mem[0] + (4 << 20), f"allocated memory grew from {mem[0]} to {mem[1]} bytes over 35 eager steps"


@pytest.mark.parametrize("gemm,bound", [("tc", 1e-3), ("auto", 2e-2)])
def test_mmg_gradients_tensor_core_engine_float64_oracle(gemm, bound):
    """MMG.forward (2 layers, 8 heads, mmgnet.json dims) on the tensor-core engines against the oracle in float64, stage
    inputs and every parameter. 3xTF32 projections (forward error ~1e-6): no ReLU gate / arg-max route moves on this case
    and the backward is accurate to 1e-3 per tensor. Default BF16x3 projections (forward error ~1e-5): a handful of
    pre-activations that are zero to 1e-5 land on the other side of their ReLU, each moving one row of a gradient by
    O(|dy| |w|) - the routing ambiguity described above - so the per-tensor bound is the 'norm' one."""
    from vlsat_b200 import train_path as T
    old_engine = ops._engine
    ops.set_gemm_engine(gemm)
    try:
        _mmg_gradients_vs_float64(bound)
    finally:
        ops._engine = old_engine


def _mmg_gradients_vs_float64(bound):
    model = V.Mmgnet(cases.model_config({}), 160, 26)
    model.load_state_dict(cases.seeded_state(model, 0))
    m = model.mmg
    sd = {"mmg." + k: v.clone().double().requires_grad_(True) for k, v in m.state_dict().items()}
    b = cases.MMGNET_CASES["mmgnet_cfg1"][1]()
    n, e = b.descriptor.shape[0], b.edge_indices.shape[1]
    g = torch.Generator().manual_seed(3)
    ins = [torch.randn(n, 512, generator=g), torch.randn(n, 512, generator=g),
           torch.randn(e, 512, generator=g).relu(), torch.randn(e, 512, generator=g).relu()]
    centres = b.descriptor[:, :3].contiguous()
    ref_in = [t.double().requires_grad_(True) for t in ins]
    ref_out = O.mmg_forward(sd, "mmg.", *ref_in, b.edge_indices, b.batch_ids, centres.double(), 2, 8)
    cases.scalar_loss(ref_out, seed=11).backward()
    m = m.to(DEV).eval()
    got_in = [t.to(DEV).requires_grad_(True) for t in ins]
    got_out = m(*got_in, b.edge_indices.to(DEV), b.batch_ids.to(DEV), centres.to(DEV))
    cases.scalar_loss(got_out, seed=11).backward()
    for i, (a, r) in enumerate(zip(got_out, ref_out)):
        assert_close(a, r.float(), f"MMG output {i}", atol_scale=FEATURE_ATOL_SCALE)
    pairs = [(f"input {i}", a.grad, r.grad) for i, (a, r) in enumerate(zip(got_in, ref_in))]
    pairs += [(k, p.grad, sd["mmg." + k].grad) for k, p in m.named_parameters()]
    scale = max(r.norm().item() for _, _, r in pairs if r is not None)
    for name, a, r in pairs:
        if r is None:
            continue
        assert a is not None, name
        err = (a.cpu().double() - r).norm().item()
        print(f"[mmg grads] {name}: rel err {err / (r.norm().item() + 1e-30):.3g}")
        assert err <= bound * r.norm().item() + 1e-6 * scale, f"{name}: ||g - ref|| = {err:.3g}, ||ref|| = {r.norm().item():.3g}"


def test_graphed_train_step_matches_eager_and_redraws_dropout():
    model = V.Mmgnet(cases.model_config({}), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    model = model.to(DEV).train()
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]().to(DEV)
    b2 = cases.MMGNET_CASES["mmgnet_ragged"][1]().to(DEV)
    b2.obj_2d_feats.mul_(0.5)                                  # same shapes / scene composition, different data
    loss_fn = lambda outs: cases.scalar_loss(outs[:7], seed=7)
    saved = {m: m.p for m in model.modules() if isinstance(m, torch.nn.Dropout)}
    _no_dropout(model)
    step = V.GraphedTrainStep(model, loss_fn)
    for batch in (b, b2, b):                                   # capture, then replays on new data
        model.zero_grad(set_to_none=True)
        loss, outs = step(*batch.forward_args())
        got = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        got_loss = loss.detach().clone()
        model.zero_grad(set_to_none=True)
        model.mlp_3d[1].momentum = 0.0                         # keep BatchNorm running stats fixed for the comparison run
        ref_loss = loss_fn(model(*batch.forward_args(), istrain=True))
        ref_loss.backward()
        model.mlp_3d[1].momentum = 0.1
        assert torch.allclose(got_loss, ref_loss.detach(), rtol=1e-4, atol=1e-4)
        assert len(got) >= 100
        for k, p in model.named_parameters():
            if p.grad is not None:
                # floor 1e-6: gradients that are zero in exact arithmetic (key biases: softmax is shift invariant) are
                # rounding noise whose value depends on the order of the atomic accumulations
                if k.endswith("fc_k.bias"):
                    wmax = dict(model.named_parameters())[k[:-4] + "weight"].grad.abs().max().item()
                    assert got[k].abs().max().item() <= 1e-4 * wmax + 1e-5 and p.grad.abs().max().item() <= 1e-4 * wmax + 1e-5, k
                    continue
                # atomic accumulation order differs between the two runs: compare per tensor, not per element
                err, ref = (got[k] - p.grad).norm().item(), p.grad.norm().item()
                assert err <= 1e-3 * ref + 1e-6 * p.grad.numel() ** 0.5, f"{k}: ||graphed - eager|| = {err:.3g}, ||eager|| = {ref:.3g}"
    assert len(step._graphs) == 1 and step.kernels_per_replay > 500
    # dropout on: every replay draws new masks
    for m, p in saved.items():
        m.p = p
    step2 = V.GraphedTrainStep(model, loss_fn)
    l1 = step2(*b.forward_args())[0].detach().clone()
    l2 = step2(*b.forward_args())[0].detach().clone()
    assert torch.isfinite(l1) and torch.isfinite(l2) and not torch.equal(l1, l2)


def test_graph_capture_leaves_batchnorm_statistics_as_one_step_would():
    """Round-1 advisor finding: the warm-up passes of GraphedTrainStep.capture updated the BatchNorm running statistics
    (three updates and num_batches_tracked += 3 on the first step of every new signature). After the first call the buffers
    must equal those of ONE eager training step from the same state."""
    def fresh():
        m = V.Mmgnet(cases.model_config({}), 160, 26)
        m.load_state_dict(cases.seeded_state(m, cases.MMGNET_WEIGHT_SEED))
        return m.to(DEV).train()
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]().to(DEV)
    loss_fn = lambda outs: cases.scalar_loss(outs[:7], seed=7)
    eager, graphed_model = fresh(), fresh()
    _no_dropout(eager); _no_dropout(graphed_model)
    loss_fn(eager(*b.forward_args(), istrain=True)).backward()
    step = V.GraphedTrainStep(graphed_model, loss_fn)
    step(*b.forward_args())
    be, bg = eager.mlp_3d[1], graphed_model.mlp_3d[1]
    assert int(bg.num_batches_tracked) == int(be.num_batches_tracked) == 1
    assert_close(bg.running_mean, be.running_mean, "running_mean after the first graphed step", rtol=1e-5, atol=1e-6)
    assert_close(bg.running_var, be.running_var, "running_var after the first graphed step", rtol=1e-5, atol=1e-6)
    step(*b.forward_args())
    loss_fn(eager(*b.forward_args(), istrain=True)).backward()
    assert int(bg.num_batches_tracked) == int(be.num_batches_tracked) == 2
    assert_close(bg.running_mean, be.running_mean, "running_mean after the second step", rtol=1e-5, atol=1e-6)
