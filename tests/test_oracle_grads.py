"""The oracle's backward (torch.autograd over oracle/vlsat_oracle.py) is pinned on gradient fixtures produced by the
UNMODIFIED reference (oracle/make_golden_grads.py -> tests/golden/grads.pt). CPU only."""
import pytest
import torch

import cases
from conftest import assert_close, assert_grad_summary_close, grad_floor
from oracle import vlsat_oracle as O


def _leaf_state(module_or_schema, seed):
    sd = cases.seeded_state(module_or_schema, seed)
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}


@pytest.fixture(scope="module")
def grads(golden):
    return golden("grads")


@pytest.mark.parametrize("name", cases.GRAD_MMGNET_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_oracle_mmgnet_gradients_match_reference(name, mode, grads):
    import vlsat_b200 as V
    over, make = cases.MMGNET_CASES[name]
    cfg = cases.model_config(over)
    net = V.Mmgnet(cfg, 160, 26)
    sd = _leaf_state(net, cases.MMGNET_WEIGHT_SEED)
    b = make()
    m = cfg["MODEL"]
    outs = O.mmgnet_forward(sd, *b.forward_args(), istrain=True, depth=m["N_LAYERS"], num_heads=m["NUM_HEADS"],
                            aggr=m["GCN_AGGR"], train_bn=(mode == "train"))
    want = grads[f"{name}.{mode}"]
    for i in range(4):
        assert_close(outs[i], want["outs"][i], f"{name}.{mode} output {i}", atol_scale=1e-5)
    cases.scalar_loss(outs[:7], seed=7).backward()
    checked = 0
    floor = grad_floor(want["grads"])
    for k, w in want["grads"].items():
        assert sd[k].grad is not None, f"oracle produced no gradient for {k}"
        assert_grad_summary_close(sd[k].grad, w, f"{name}.{mode} d{k}", floor=floor)
        checked += 1
    assert checked >= 100


@pytest.mark.parametrize("name", cases.GRAD_GAT_CASES)
def test_oracle_gat_gradients_match_reference(name, grads):
    import vlsat_b200 as V
    kw, n, e, iso, seed = cases.GAT_CASES[name]
    layer = V.GraphEdgeAttenNetwork(**kw)
    sd = _leaf_state(layer, seed)
    x, ef, ei = cases.gat_inputs(name)
    x.requires_grad_(True); ef.requires_grad_(True)
    xo, eo, _ = O.gat_layer(sd, "", x, ef, ei, kw["num_heads"], kw["aggr"], kw.get("flow", "target_to_source"), kw.get("use_edge", True))
    cases.scalar_loss([xo, eo], seed=8).backward()
    want = grads[name]["grads"]
    floor = grad_floor(want)
    for k, w in want.items():
        got = x.grad if k == "input.x" else ef.grad if k == "input.edge" else sd[k].grad
        assert got is not None, k
        assert_grad_summary_close(got, w, f"{name} d{k}", floor=floor)
