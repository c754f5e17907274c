"""N1 on the GPU, all through the C ABI:

  * every loss kernel (forward value and gradient) against the float64 autograd of the oracle's restatement, including
    saturated probabilities (the -100 log clamp), classes absent from the batch, strided feature rows;
  * ``reference_loss`` on the forward outputs the UNMODIFIED reference produced against the loss it computed
    (tests/golden/train_step.pt);
  * ``FusedAdamW`` against ``torch.optim.AdamW`` + ``CosineAnnealingLR`` - the very classes the reference instantiates
    (SGFN_MMG/model.py:143-157) - over several steps, with weight decay, amsgrad, odd sizes and misaligned tensors;
  * two full ``process_train`` iterations (forward, losses, backward, AdamW) against the reference's own two iterations:
    the loss of both steps and the parameter changes;
  * the CUDA-graph ``TrainStep`` against the eager one.
"""
import math

import pytest
import torch
import torch.nn.functional as F

import cases
import vlsat_b200 as V
from oracle import vlsat_oracle as O
from vlsat_b200 import ops
from vlsat_b200 import train_glue as G

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand_outs(n, e, seed, strided=True):
    g = torch.Generator().manual_seed(seed)
    log3, log2 = torch.randn(n, 160, generator=g) * 3, torch.randn(n, 160, generator=g) * 3
    rel3, rel2 = torch.sigmoid(torch.randn(e, 26, generator=g) * 2), torch.sigmoid(torch.randn(e, 26, generator=g) * 2)
    wide = torch.randn(n, 768, generator=g)
    f2 = torch.randn(n, 512, generator=g)
    f2[: n // 2] = wide[: n // 2, :512] * 1.5 + 0.05 * f2[: n // 2]        # cos > 0.8 for half of the rows: margin inactive
    e2d = torch.randn(e, 512, generator=g) * 4
    gt_cls = torch.randint(0, 160, (n,), generator=g)
    gt_rel = (torch.rand(e, 26, generator=g) < 0.06).float()
    gt_rel[:, 7] = 0
    text = F.normalize(torch.randn(e, 512, generator=g), dim=-1)
    return [log3, log2, rel3, rel2, wide, f2, e2d], gt_cls, gt_rel, text


@pytest.mark.parametrize("n,e,seed", [(10, 30, 0), (640, 9600, 1), (33, 1, 2), (1, 77, 3)])
def test_losses_and_gradients_against_float64_autograd(n, e, seed):
    outs, gt_cls, gt_rel, text = _rand_outs(n, e, seed)
    if seed == 0:
        outs[2][0, :4] = torch.tensor([0.0, 1.0, 0.0, 1.0])          # saturated sigmoid outputs: log clamp at -100
        gt_rel[0, :4] = torch.tensor([1.0, 0.0, 0.0, 1.0])
    ref_in = [o.double().requires_grad_(True) for o in outs]
    ref_feats = list(ref_in)
    ref_feats[4] = ref_in[4][:, :512]
    want, want_terms = O.train_losses(ref_feats, gt_cls, gt_rel.double(), text.double())
    want.backward()

    dev_in = [o.to(DEV).requires_grad_(True) for o in outs]
    feats = list(dev_in)
    feats[4] = dev_in[4][:, :512]                                      # row stride 768, as obj_feature[..., :512]
    loss, terms = G.reference_loss(feats, gt_cls.to(DEV), gt_rel.to(DEV), text.to(DEV))
    loss.backward()
    assert abs(loss.item() - want.item()) <= 2e-5 * abs(want.item()), (loss.item(), want.item())
    for k, v in want_terms.items():
        assert abs(terms[k].item() - v.item()) <= 2e-5 * abs(v.item()) + 1e-7, k
    for i, (a, r) in enumerate(zip(dev_in, ref_in)):
        got, ref = a.grad.double().cpu(), r.grad
        assert torch.isfinite(got).all()
        tol = 1e-4 * ref.abs() + 1e-5 * ref.abs().max()
        bad = (got - ref).abs() > tol
        # sign(0) / the margin boundary are measure-zero events; saturated BCE entries are compared like the rest
        assert bad.sum().item() == 0, f"input {i}: {int(bad.sum())} of {bad.numel()} gradient entries off, worst {((got - ref).abs() - tol).max().item():.3g}"
    assert dev_in[4].grad[:, 512:].abs().max().item() == 0.0


def test_loss_is_bitwise_reproducible():
    outs, gt_cls, gt_rel, text = _rand_outs(640, 9600, 5)
    dev = [o.to(DEV) for o in outs]
    dev[4] = dev[4][:, :512]
    a = G.reference_loss(dev, gt_cls.to(DEV), gt_rel.to(DEV), text.to(DEV))[0].clone()
    b = G.reference_loss(dev, gt_cls.to(DEV), gt_rel.to(DEV), text.to(DEV))[0].clone()
    assert torch.equal(a, b)


def test_dynamic_class_weights_kernel():
    g = torch.Generator().manual_seed(4)
    gt = (torch.rand(9600, 26, generator=g) < 0.05).float()
    gt[:, 3] = 0
    w = G.rel_class_weights(gt.to(DEV)).cpu()
    assert torch.allclose(w, O.rel_class_weights(gt), rtol=1e-6, atol=0)
    assert torch.allclose(G.rel_class_weights(gt.to(DEV), ignore_none_rel=True).cpu(), O.rel_class_weights(gt, ignore_none_rel=True), rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", cases.TRAIN_CASES)
def test_reference_loss_on_the_reference_forward_outputs(name, golden):
    gold = golden("train_step")[name]
    b = cases.MMGNET_CASES[name][1]()
    gt_cls, gt_rel, text = cases.train_targets(b)
    loss, _ = G.reference_loss([o.to(DEV) for o in gold["outs"]], gt_cls.to(DEV), gt_rel.to(DEV), text.to(DEV))
    assert abs(loss.item() - gold["losses"][0]) <= 1e-5 * abs(gold["losses"][0])


@pytest.mark.parametrize("wd,amsgrad", [(0.0, False), (0.05, False), (0.01, True)])
def test_fused_adamw_against_torch_adamw_with_cosine_schedule(wd, amsgrad):
    g = torch.Generator().manual_seed(8)
    flat = torch.randn(70001, generator=g)
    shapes = [(), (3,), (129, 127), (16384,), (16385,), (1024, 64)]
    lrs = [1e-3, 1e-3, 2.5e-4, 5e-4, 1e-4, 1e-3]
    mine = [torch.nn.Parameter(torch.randn(s, generator=g).to(DEV)) for s in shapes]
    mis = flat.to(DEV)[1:]                                     # contiguous but only 4-byte aligned: the scalar path
    mine.append(torch.nn.Parameter(mis))
    lrs.append(1e-3)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    unused_mine, unused_ref = torch.nn.Parameter(torch.ones(5, device=DEV)), torch.nn.Parameter(torch.ones(5, device=DEV))
    groups = lambda ps, extra: [dict(params=[p], lr=lr, weight_decay=wd, amsgrad=amsgrad) for p, lr in zip(ps, lrs)] + \
                               [dict(params=[extra], lr=1e-3, weight_decay=wd, amsgrad=amsgrad)]
    opt = G.FusedAdamW(groups(mine, unused_mine), t_max=4)
    ref_opt = torch.optim.AdamW(groups(ref, unused_ref), betas=(0.9, 0.999), eps=1e-8)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(ref_opt, T_max=4, last_epoch=-1)
    for step in range(6):                                      # past T_max: the cosine turns back up, as torch's does
        for p, r in zip(mine, ref):
            gr = torch.randn(p.shape, generator=g).to(DEV) * (10.0 ** (step - 3))
            p.grad, r.grad = gr.clone(), gr.clone()
        v0 = mine[2]._version
        assert all(math.isclose(a, b, rel_tol=1e-6, abs_tol=1e-12) for a, b in zip(opt.last_lr, sched.get_last_lr()))
        opt.step(); ref_opt.step(); sched.step()
        opt.zero_grad(); ref_opt.zero_grad()
        assert mine[2]._version > v0                           # derived-weight caches see the update
        for i, (p, r) in enumerate(zip(mine, ref)):
            assert torch.allclose(p.detach(), r.detach(), rtol=2e-5, atol=2e-7), (step, i, (p - r).abs().max().item())
    assert torch.equal(unused_mine.detach(), torch.ones(5, device=DEV))     # no gradient: skipped, like torch does
    assert opt.steps_done == 6 and int(opt._step_dev.item()) == 6


def _train_model(engine):
    ops.set_gemm_engine(engine)
    model = V.Mmgnet(cases.model_config({}), 160, 26)
    model.load_state_dict(cases.seeded_state(model, cases.MMGNET_WEIGHT_SEED))
    return model.to(DEV).eval()                 # eval(): the mode the reference fixture ran in (make_golden_train.py)


@pytest.mark.parametrize("name", cases.TRAIN_CASES)
def test_two_process_train_iterations_against_the_reference(name, golden):
    gold = golden("train_step")[name]
    try:
        model = _train_model("simt")           # exact-fp32 engine: Adam turns gradient noise near zero into +-lr steps
        before = {k: p.detach().clone() for k, p in model.named_parameters()}
        opt = G.build_optimizer(model, lr=gold["lr"], max_iteration=gold["t_max"])
        step = G.TrainStep(model, opt, graphed=False)
        b = cases.MMGNET_CASES[name][1]().to(DEV)
        targets = [t.to(DEV) for t in cases.train_targets(cases.MMGNET_CASES[name][1]())]
        losses, gmax = [], {}
        for it in range(cases.TRAIN_STEPS):
            if it == 0:                        # gradient magnitudes of step 1, to tell signal from rounding noise below
                outs = model(*b.forward_args(), istrain=True)
                G.reference_loss(outs, *targets)[0].backward()
                gmax = {k: p.grad.abs().max().item() for k, p in model.named_parameters() if p.grad is not None}
                model.zero_grad(set_to_none=True)
            loss, _ = step.step(*b.forward_args(), *targets)
            losses.append(loss.item())
    finally:
        ops.set_gemm_engine("auto")
    assert abs(losses[0] - gold["losses"][0]) <= 1e-4 * gold["losses"][0], (losses, gold["losses"])
    assert abs(losses[1] - gold["losses"][1]) <= 1e-3 * gold["losses"][1], (losses, gold["losses"])
    assert all(math.isclose(a, b, rel_tol=1e-6) for a, b in zip(opt.last_lr, gold["last_lr"]))
    after = dict(model.named_parameters())
    moved = {k for k in before if not torch.equal(before[k], after[k].detach())}
    scale = max(gmax.values())
    # the same parameters move (frozen adapter and the unused triplet_projector_3d do not), up to noise-only gradients
    assert all(gmax.get(k, 0.0) < 1e-6 * scale for k in moved ^ set(gold["delta"])), sorted(moved ^ set(gold["delta"]))
    assert len(moved) >= 180
    checked = 0
    for k, summ in gold["delta"].items():
        if gmax[k] < 1e-6 * scale:             # zero in exact arithmetic (key biases under a softmax): Adam amplifies rounding noise
            continue
        d = (after[k].detach() - before[k]).double().cpu().reshape(-1)
        ref = summ["full"].double() if "full" in summ else summ["head"].double()
        got = d[: ref.numel()]
        err, nrm = (got - ref).norm().item(), ref.norm().item()
        assert err <= 0.05 * nrm + 1e-9, f"{k}: ||delta - ref|| = {err:.3g} vs ||ref|| = {nrm:.3g}"
        if "norm" in summ:
            assert abs(d.norm().item() - summ["norm"]) <= 0.02 * summ["norm"], k
        checked += 1
    assert checked >= 150, checked


def test_graphed_train_step_matches_eager_train_step():
    name = "mmgnet_ragged"
    b = cases.MMGNET_CASES[name][1]().to(DEV)
    targets = [t.to(DEV) for t in cases.train_targets(cases.MMGNET_CASES[name][1]())]
    runs = []
    for graphed in (False, True):
        model = _train_model("auto")
        opt = G.build_optimizer(model, lr=1e-4, max_iteration=1000)
        step = G.TrainStep(model, opt, graphed=graphed)
        with torch.no_grad():
            model(*b.forward_args(), istrain=False)                # fills every derived-weight cache with the initial weights
        losses = []
        for _ in range(3):
            loss, terms = step.step(*b.forward_args(), *targets)
            losses.append(loss.item())
            assert set(terms) == set(G.LOSS_TERMS[1:]) and all(torch.isfinite(v) for v in terms.values())
        runs.append((losses, {k: p.detach().clone() for k, p in model.named_parameters()}))
        if graphed:
            assert step.kernels_per_step > 500 and opt.steps_done == 3
            # the eval-mode inference path after training sees the updated weights (version counters were bumped)
            with torch.no_grad():
                outs = model(*b.forward_args(), istrain=False)
            for m in model.modules():
                if hasattr(m, "_cache"):
                    m._cache.clear()
            ops._weight_splits.clear()
            with torch.no_grad():
                fresh = model(*b.forward_args(), istrain=False)
            for a, c in zip(outs, fresh):
                assert torch.equal(a, c)
    (l_e, p_e), (l_g, p_g) = runs
    assert l_e[0] > l_e[2] and l_g[0] > l_g[2]                     # three AdamW steps on one batch reduce its loss
    for a, c in zip(l_e, l_g):
        assert abs(a - c) <= 2e-3 * abs(a), (l_e, l_g)
    # same update direction tensor by tensor (atomic accumulation order differs between the two runs)
    init = cases.seeded_state(_train_model("auto"), cases.MMGNET_WEIGHT_SEED)
    agree = []
    for k in p_e:
        de, dg = (p_e[k].cpu() - init[k]).flatten().double(), (p_g[k].cpu() - init[k]).flatten().double()
        if de.norm() > 0 and de.numel() >= 64:
            agree.append(float(torch.dot(de, dg) / (de.norm() * dg.norm() + 1e-30)))
    agree.sort()
    assert agree[len(agree) // 10] > 0.9, agree[:20]


def test_zero_arena_serves_the_backward_accumulators(monkeypatch):
    """The small zero-initialised accumulation targets of the backward (bias gradients, LayerNorm dgamma / dbeta, scatter-add
    targets) come out of one per-step arena (ops.zero_arena): same gradients as with one torch.zeros each, the second step
    fits entirely, and nothing aliases across steps (eager and graphed)."""
    from vlsat_b200 import autograd as A
    from vlsat_b200 import synth
    b = synth.make_config_batch("cfg1", seed=5).to(DEV)
    gen = torch.Generator().manual_seed(11)
    n, e = b.obj_points.shape[0], b.edge_indices.shape[1]
    gt_cls = torch.randint(0, 160, (n,), generator=gen).to(DEV)
    gt_rel = (torch.rand(e, 26, generator=gen) < 0.1).float().to(DEV)
    text = torch.nn.functional.normalize(torch.randn(e, 512, generator=gen), dim=-1).to(DEV)

    def grads(arena: str, graphed: bool):
        monkeypatch.setenv("VLSAT_ZERO_ARENA", arena)
        A.DropoutState.manual_seed(99)
        model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
        synth.load_seeded(model, 0)
        model = model.to(DEV).train()
        ts = G.TrainStep(model, G.build_optimizer(model, lr=0.0, max_iteration=10), graphed=graphed)
        out = []
        for _ in range(2):
            loss = ts.forward_backward(*b.forward_args(), gt_cls, gt_rel, text)
            out.append((float(loss.detach()), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}))
            if not graphed:
                model.zero_grad(set_to_none=True)
        owner = ts._graphed if graphed else ts
        return out, getattr(owner, "_zero_arena_bytes", 0), getattr(owner, "_zero_arena_fallbacks", -1)

    for graphed in (False, True):
        (ref, rb, _), (got, gb, gf) = grads("0", graphed), grads("1", graphed)
        assert rb == 0 and gb > 0 and gf == 0, (rb, gb, gf)
        for (l0, g0), (l1, g1) in zip(ref, got):
            assert abs(l0 - l1) <= 1e-5 * abs(l0)
            assert g0.keys() == g1.keys()
            # gradients that vanish in exact arithmetic (everything upstream of the distance bias) are run-to-run noise of the
            # atomic accumulation order: a floor relative to the largest gradient of the step, as in conftest.grad_floor
            floor = 1e-6 * max(g.abs().max().item() for g in g0.values())
            bad = {k: ((g0[k] - g1[k]).abs().max().item(), g0[k].abs().max().item()) for k in g0
                   if (g0[k] - g1[k]).abs().max().item() > 2e-4 * g0[k].abs().max().item() + floor}
            assert not bad, bad
