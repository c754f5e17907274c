"""Tail split of the tensor-core projections (csrc/gemm_tc.cu): tiles beyond the last whole round of the 148-CTA persistent
grid are split along K, their partials summed in split order by linear_tail_fixup_kernel which also applies the fused
epilogue. Checked against float64, against the unsplit schedule (VLSAT_TAIL_SPLIT=0: only the fp32 summation order of
the tail tiles differs) and for run-to-run determinism, for every epilogue form the path uses (bias / ReLU / row gathers /
residual / emitted bf16 pair / pair-only output) and for the stored-operand dX GEMM."""
import math

import pytest
import torch

from conftest import assert_close
from vlsat_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# config #2's edge projections (300 and 150 tiles), a ragged M edge inside the tail, K = 1024, N = 1024 (600 tiles)
SHAPES = [(9600, 512, 512), (9600, 256, 512), (9533, 512, 1024), (9600, 1024, 512), (19100, 264, 768)]


def test_plan_covers_the_config2_shapes():
    lib = _lib.load()
    assert lib.vlsat_linear_tail_workspace_bytes(9600, 512, 512, 4) == 4 * 8 * 128 * 128 * 4      # 4 tiles x 8 K blocks
    assert lib.vlsat_linear_tail_workspace_bytes(9600, 256, 512, 4) == 2 * 8 * 128 * 128 * 4
    assert lib.vlsat_linear_tail_workspace_bytes(640, 512, 512, 4) == 0                            # under one round
    assert lib.vlsat_linear_tail_workspace_bytes(9600, 512, 256, 4) == 0                           # too few K blocks to pay
    assert lib.vlsat_gemm_pairs_workspace_bytes(2, 9600, 512, 512) == 4 * 8 * 128 * 128 * 4


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("form", ["plain", "bias_relu_pair", "gather", "residual", "pair_only"])
def test_linear_tail_split(m, n, k, form, monkeypatch):
    x, w, b = rnd(m, k, seed=1), rnd(n, k, seed=2) / math.sqrt(k), rnd(n, seed=3)
    rows = 97
    g_ = torch.Generator().manual_seed(4)
    ia, ib = torch.randint(0, rows, (m,), generator=g_).to(DEV), torch.randint(0, rows, (m,), generator=g_).to(DEV)
    ga, gb, res = rnd(rows, n, seed=5), rnd(rows, n, seed=6), rnd(m, n, seed=7)
    z = x.double() @ w.double().t()

    def run():
        if form == "plain":
            return ops.linear(x, w), None
        if form == "bias_relu_pair":
            return ops.linear(x, w, b, act=ops.ACT_RELU, emit_split="bf16")
        if form == "gather":
            return ops.linear(x, w, b, act=ops.ACT_RELU, gather=(ga, ia, gb, ib)), None
        if form == "residual":
            return ops.linear(x, w, b, residual=res, alpha=0.5, beta=2.0), None
        return ops.linear(x, w, b, act=ops.ACT_RELU, emit_split="bf16", want_y=False)

    want = {"plain": z, "bias_relu_pair": torch.relu(z + b.double()), "pair_only": torch.relu(z + b.double()),
            "gather": torch.relu(z + b.double() + ga.double()[ia] + gb.double()[ib]),
            "residual": 0.5 * (z + b.double()) + 2.0 * res.double()}[form]
    assert ops._tail_bytes("linear", m, n, k, ops.ENGINES["bf16x3"]) > 0, "shape no longer exercises the tail split"
    y, pair = run()
    y2, pair2 = run()
    monkeypatch.setenv("VLSAT_TAIL_SPLIT", "0")
    y0, pair0 = run()
    if y is not None:
        assert_close(y, want.float(), f"tail split {form} vs float64", rtol=1e-4, atol=0.0, atol_scale=2e-5)
        assert_close(y, y0, f"tail split {form} vs unsplit", rtol=1e-5, atol=0.0, atol_scale=2e-6)
        assert torch.equal(y, y2), "tail split must be deterministic"
    if pair is not None:
        got = pair[0].float() + pair[1].float()
        assert_close(got, want.float(), f"tail split {form} pair vs float64", rtol=1e-4, atol=0.0, atol_scale=2e-5)
        assert_close(got, pair0[0].float() + pair0[1].float(), f"tail split {form} pair vs unsplit", rtol=1e-5, atol=0.0, atol_scale=2e-6)
        assert torch.equal(pair[0], pair2[0]) and torch.equal(pair[1], pair2[1])


@pytest.mark.parametrize("m,n,k", [(9600, 512, 512), (9600, 256, 1024), (9533, 1024, 512)])
def test_gemm_nn_tail_split(m, n, k, monkeypatch):
    a, b = rnd(m, k, seed=8), rnd(k, n, seed=9) / math.sqrt(k)
    ap, bp = ops.bf16_split(a), ops.bf16_split(b)
    assert ops._tail_bytes("nn", m, n, k) > 0
    got = ops.gemm_nn(ap, bp, n)
    assert_close(got, (a.double() @ b.double()).float(), "gemm_nn tail split vs float64", rtol=1e-3, atol=0.0, atol_scale=1e-4)
    assert torch.equal(got, ops.gemm_nn(ap, bp, n))
    monkeypatch.setenv("VLSAT_TAIL_SPLIT", "0")
    assert_close(got, ops.gemm_nn(ap, bp, n), "gemm_nn tail split vs unsplit", rtol=1e-5, atol=0.0, atol_scale=2e-6)


def test_tail_split_in_bf16_mode(monkeypatch):
    """Single-pass mode: the plan asks for more saved K blocks (shorter K blocks), K = 1024 still splits."""
    m, n, k = 9600, 512, 1024
    x, w = rnd(m, k, seed=10), rnd(n, k, seed=11) / math.sqrt(k)
    ops.set_precision("bf16")
    try:
        y = ops.linear(x, w)
        monkeypatch.setenv("VLSAT_TAIL_SPLIT", "0")
        y0 = ops.linear(x, w)
    finally:
        ops.set_precision("fp32")
    assert_close(y, y0, "bf16-mode tail split vs unsplit", rtol=1e-5, atol=0.0, atol_scale=2e-6)
