"""N3: the evaluation-rank kernels against the oracle (pinned bit-exactly on the reference's own functions,
tests/test_oracle_eval.py) and the reference's fixtures. First run on a B200 in round 2 (4 passed, gpurun_out/r2_gpunext.log), hence promoted from
``gpu_next`` to ``gpu``. Ranks are integers: the bar is bit-exact (probabilities fed from the same CPU softmax)."""
import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import vlsat_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(cases.EVAL_CASES))
def test_rank_kernels_match_reference_fixture(name, golden):
    from vlsat_b200 import eval_ranks as R
    gold = golden("eval_ranks")[name]
    logits, rel, gt_cls, gt_rel, edges = (t.cuda() for t in cases.eval_inputs(name))
    assert torch.equal(R.evaluate_topk_object(logits, gt_cls, 11).cpu(), gold["obj"])
    assert torch.equal(R.evaluate_topk_predicate(rel, gt_rel, 6).cpu(), gold["rel"])
    probs = F.softmax(logits.cpu(), dim=-1).cuda()                    # the reference's CPU softmax, bit for bit
    assert torch.equal(R.evaluate_triplet_topk(logits, rel, gt_cls, gt_rel, edges, 101, obj_probs=probs).cpu(), gold["triplet"])
    # device softmax: last-bit differences can only move a rank where two scores are within an ulp
    dev = R.evaluate_triplet_topk(logits, rel, gt_cls, gt_rel, edges, 101).cpu()
    assert dev.shape == gold["triplet"].shape and (dev != gold["triplet"]).float().mean().item() < 0.02
    assert torch.allclose(R.softmax_rows(logits).cpu(), F.softmax(logits.cpu(), -1), rtol=1e-5, atol=1e-9)


def test_rank_kernels_config2_scene_against_oracle():
    from vlsat_b200 import eval_ranks as R
    g = torch.Generator().manual_seed(3)
    n, e = 40, 600
    logits, rel = torch.randn(n, 160, generator=g) * 4, torch.sigmoid(torch.randn(e, 26, generator=g) * 3)
    gt_cls, gt_rel = torch.randint(0, 160, (n,), generator=g), (torch.rand(e, 26, generator=g) < 0.05).float()
    edges = torch.randint(0, n, (e, 2), generator=g)
    probs = F.softmax(logits, dim=-1)
    got = R.evaluate_triplet_topk(logits.cuda(), rel.cuda(), gt_cls.cuda(), gt_rel.cuda(), edges.cuda(), 101, obj_probs=probs.cuda()).cpu()
    assert torch.equal(got, O.topk_triplet_ranks(logits, rel, gt_cls, gt_rel, edges, 101))
    assert torch.equal(R.evaluate_topk_predicate(rel.cuda(), gt_rel.cuda(), 6).cpu(), O.topk_predicate_ranks(rel, gt_rel, 6))
    assert torch.equal(R.evaluate_topk_object(logits.cuda(), gt_cls.cuda(), 11).cpu(), O.topk_object_ranks(logits, gt_cls, 11))
    empty = R.evaluate_triplet_topk(logits.cuda(), rel[:0].cuda(), gt_cls.cuda(), gt_rel[:0].cuda(), edges[:0].cuda(), 101, obj_probs=probs.cuda())
    assert empty.numel() == 0


def test_tied_labels_reach_zero_and_negative_adjusted_ranks():
    from vlsat_b200 import eval_ranks as R
    rel = torch.full((2, 26), 0.3)
    rel[0, [1, 2, 3]] = 1.0
    gt_rel = torch.zeros(2, 26)
    gt_rel[0, [1, 2, 3]] = 1
    assert R.evaluate_topk_predicate(rel.cuda(), gt_rel.cuda(), 6).cpu().tolist() == [1, 0, -1, 1]
    logits, gt_cls, edges = torch.randn(2, 160, generator=torch.Generator().manual_seed(1)), torch.tensor([5, 7]), torch.tensor([[0, 1], [1, 0]])
    probs = F.softmax(logits, -1)
    got = R.evaluate_triplet_topk(logits.cuda(), rel.cuda(), gt_cls.cuda(), gt_rel.cuda(), edges.cuda(), 101, obj_probs=probs.cuda()).cpu()
    assert torch.equal(got, O.topk_triplet_ranks(logits, rel, gt_cls, gt_rel, edges, 101))


def test_train_metrics_match_the_reference_metric_code():
    """SURVEY 8f N1 remainder: the recall figures process_train logs after backward() (SGFN_MMG/model.py:422-432) from the
    rank kernels, as device scalars, against the oracle's rank functions (pinned on evaluate_topk_object / _predicate)."""
    from vlsat_b200 import eval_ranks as R
    g = torch.Generator().manual_seed(9)
    n, e = 640, 9600
    o3, o2 = torch.randn(n, 160, generator=g) * 3, torch.randn(n, 160, generator=g) * 3
    r3, r2 = torch.sigmoid(torch.randn(e, 26, generator=g) * 2), torch.sigmoid(torch.randn(e, 26, generator=g) * 2)
    gt_cls, gt_rel = torch.randint(0, 160, (n,), generator=g), (torch.rand(e, 26, generator=g) < 0.04).float()
    got = R.train_metrics(o3.cuda(), o2.cuda(), r3.cuda(), r2.cuda(), gt_cls.cuda(), gt_rel.cuda())
    assert len(got) == 12 and all(v.is_cuda and v.dim() == 0 for v in got.values())

    def want(ranks, ks):
        return [100.0 * float((ranks <= k).sum()) / len(ranks) for k in ks]
    exp = dict(zip(("train/Obj_R1", "train/Obj_R5", "train/Obj_R10"), want(O.topk_object_ranks(o3, gt_cls, 11), (1, 5, 10))))
    exp.update(zip(("train/Obj_R1_2d", "train/Obj_R5_2d", "train/Obj_R10_2d"), want(O.topk_object_ranks(o2, gt_cls, 11), (1, 5, 10))))
    exp.update(zip(("train/Pred_R1", "train/Pred_R3", "train/Pred_R5"), want(O.topk_predicate_ranks(r3, gt_rel, 6), (1, 3, 5))))
    exp.update(zip(("train/Pred_R1_2d", "train/Pred_R3_2d", "train/Pred_R5_2d"), want(O.topk_predicate_ranks(r2, gt_rel, 6), (1, 3, 5))))
    for k, v in exp.items():
        assert abs(float(got[k]) - v) <= 1e-4 * max(1.0, abs(v)), f"{k}: {float(got[k])} vs {v}"
