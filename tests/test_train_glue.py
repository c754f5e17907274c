"""N1, CPU side: the oracle's restatement of process_train's losses is pinned on the loss the UNMODIFIED
reference computed (tests/golden/train_step.pt, oracle/make_golden_train.py), and the host logic of the optimiser
mirror (parameter groups, cosine schedule, no CPU fallback)."""
import math
import os

import pytest
import torch

import cases
import vlsat_b200 as V
from oracle import vlsat_oracle as O
from vlsat_b200 import train_glue as G


@pytest.mark.parametrize("name", cases.TRAIN_CASES)
def test_oracle_losses_match_the_reference_process_train(name, golden):
    gold = golden("train_step")[name]
    b = cases.MMGNET_CASES[name][1]()
    gt_cls, gt_rel, text = cases.train_targets(b)
    loss, terms = O.train_losses(gold["outs"], gt_cls, gt_rel, text)
    assert abs(float(loss) - gold["losses"][0]) <= 2e-6 * abs(gold["losses"][0])
    assert len(gold["losses"]) == cases.TRAIN_STEPS and gold["losses"][1] < gold["losses"][0]
    assert all(torch.isfinite(v) for v in terms.values())


def test_oracle_dynamic_weights_statement_by_statement():
    g = torch.Generator().manual_seed(3)
    gt = (torch.rand(50, 26, generator=g) < 0.08).float()
    gt[:, 5] = 0                                             # a class absent from the batch: weight 1 / (log 1 + 1) = 1
    w = O.rel_class_weights(gt)
    assert w.shape == (26,) and abs(float(w[5]) - 1.0) < 1e-7
    want = 1.0 / (torch.log(gt.sum(0) + 1) + 1)
    assert torch.allclose(w, want)
    assert torch.allclose(O.rel_class_weights(gt, ignore_none_rel=True), want * 1e-2)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_oracle_losses_against_the_live_reference_on_fresh_targets():
    from oracle import ref_shims
    ref_shims.install()
    net, _ = ref_shims.build_reference_mmgnet(seed=0)
    b = cases.MMGNET_CASES["mmgnet_ragged"][1]()
    gt_cls, gt_rel, text = cases.train_targets(b, seed=99)
    net.get_rel_emb = lambda *a, **k: text
    seen = {}

    class Done(Exception):
        pass

    def backward(loss):
        seen["loss"] = loss.detach().clone()
        raise Done()
    net.backward = backward
    fwd = net.forward

    def forward(*a, **k):
        seen["outs"] = fwd(*a, **k)
        return seen["outs"]
    net.forward = forward
    net.iteration = 0
    args = b.forward_args()
    with pytest.raises(Done):
        net.process_train(args[0], args[1], gt_cls, args[3], gt_rel, args[2].t().contiguous(), args[4])
    loss, _ = O.train_losses([o.detach() for o in seen["outs"]], gt_cls, gt_rel, text)
    assert torch.allclose(loss, seen["loss"], rtol=1e-6, atol=0)


def _model():
    return V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)


def test_parameter_groups_follow_the_reference_optimizer(golden):
    m = _model()
    groups = G.reference_param_groups(m, 1e-4)
    assert len(groups) == 13                                                   # SGFN_MMG/model.py:143-156
    assert [round(g["lr"] / 1e-4, 6) for g in groups] == [1, 1, 1, 0.25, 0.5, 0.1, 1, 0.1, 1, 1, 1, 1, 1]
    ids = [id(p) for g in groups for p in g["params"]]
    assert len(ids) == len(set(ids))
    named = {id(p): k for k, p in m.named_parameters()}
    assert all("nn_edge" in named[id(p)] for p in groups[4]["params"]) and groups[4]["params"]
    assert not any("nn_edge" in named[id(p)] for p in groups[3]["params"])
    assert all(named[id(p)].startswith("clip_adapter.") for p in m.parameters() if id(p) not in set(ids))
    assert all(g["weight_decay"] == 0.0 and not g["amsgrad"] for g in groups)   # mmgnet.json: W_DECAY false, AMSGRAD false
    # every parameter the reference's optimiser moved is in a group; the ones it never moved are triplet_projector_3d
    moved = set(golden("train_step")["mmgnet_cfg1"]["delta"])
    in_groups = {named[i] for i in ids}
    assert moved <= in_groups and all(k.startswith("triplet_projector_3d.") for k in in_groups - moved)


def test_cosine_schedule_matches_torch_scheduler():
    p = torch.nn.Parameter(torch.zeros(3))
    ref_opt = torch.optim.AdamW([dict(params=[p], lr=1e-4)])
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(ref_opt, T_max=50, last_epoch=-1)
    opt = G.FusedAdamW([dict(params=[torch.nn.Parameter(torch.zeros(3))], lr=1e-4, weight_decay=0.0)], t_max=50)
    for _ in range(60):
        assert math.isclose(opt.last_lr[0], sched.get_last_lr()[0], rel_tol=1e-9, abs_tol=1e-12)
        ref_opt.step(); sched.step()
        opt.steps_done += 1


def test_optimizer_and_losses_have_no_cpu_fallback():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = G.FusedAdamW([dict(params=[p], lr=1e-3)])
    with pytest.raises(TypeError):
        opt.step()
    with pytest.raises(ValueError):
        G.FusedAdamW([dict(params=[p], lr=1e-3), dict(params=[p], lr=1e-3)])
    with pytest.raises(NotImplementedError):
        G.LossConfig(weight_edge="BG")
    outs = [torch.zeros(2, 160), torch.zeros(2, 160), torch.full((3, 26), 0.5), torch.full((3, 26), 0.5),
            torch.ones(2, 512), torch.ones(2, 512), torch.ones(3, 512)]
    with pytest.raises(TypeError):
        G.reference_loss(outs, torch.zeros(2, dtype=torch.int64), torch.zeros(3, 26), torch.ones(3, 512))
    cfg = G.LossConfig(lambda_o=0.1)
    assert (cfg.coef_obj, cfg.coef_rel, cfg.coef_mimic) == (0.1, 3.0, 0.1)
    cfg = G.LossConfig(lambda_o=2.0)                                            # normalised by max(lambda_r, lambda_o)
    assert (cfg.coef_obj, cfg.coef_rel) == (1.0, 1.5)


def test_optimizer_checkpoints_move_between_torch_adamw_and_fused_adamw():
    """model_base.py:72-73,115-122 save / restore optimizer.state_dict() and lr_scheduler.state_dict()."""
    g = torch.Generator().manual_seed(0)
    mk = lambda: [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(i))) for i, s in enumerate([(4, 3), (7,), ()])]
    ref_p, my_p = mk(), mk()
    layout = lambda ps: [dict(params=ps[:2], lr=1e-3, weight_decay=0.0, amsgrad=False), dict(params=ps[2:], lr=1e-4, weight_decay=0.0, amsgrad=False)]
    ref = torch.optim.AdamW(layout(ref_p))
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(ref, T_max=10, last_epoch=-1)
    for _ in range(3):
        for p in ref_p[:2]:                       # the third parameter never gets a gradient: no state, like triplet_projector_3d
            p.grad = torch.randn(p.shape, generator=g)
        ref.step(); sched.step()
    mine = G.FusedAdamW(layout(my_p), t_max=0)
    mine.load_state_dict(ref.state_dict())
    mine.load_scheduler_state_dict(sched.state_dict())
    assert mine.steps_done == 3 and mine.t_max == 10
    assert [g_["lr"] for g_ in mine.param_groups] == [1e-3, 1e-4]              # base rates, not the scheduled ones
    assert all(math.isclose(a, b, rel_tol=1e-9) for a, b in zip(mine.last_lr, sched.get_last_lr()))
    for p, q in zip(my_p[:2], ref_p[:2]):
        assert torch.equal(mine.state[id(p)]["m"], ref.state[q]["exp_avg"]) and torch.equal(mine.state[id(p)]["v"], ref.state[q]["exp_avg_sq"])
    assert id(my_p[2]) not in mine.state
    # and back: torch's optimiser accepts the state dict FusedAdamW writes
    back = torch.optim.AdamW(layout(mk()))
    back.load_state_dict(mine.state_dict())
    sd = back.state_dict()
    assert sorted(sd["state"]) == [0, 1] and float(sd["state"][0]["step"]) == 3.0
    assert torch.equal(sd["state"][1]["exp_avg_sq"], ref.state[ref_p[1]]["exp_avg_sq"])
    assert math.isclose(sd["param_groups"][0]["lr"], sched.get_last_lr()[0], rel_tol=1e-9) and sd["param_groups"][0]["initial_lr"] == 1e-3
    s2 = mine.scheduler_state_dict()
    assert s2["last_epoch"] == sched.state_dict()["last_epoch"] and s2["T_max"] == 10
