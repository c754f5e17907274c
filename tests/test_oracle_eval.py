"""N3, parity gate prepared ahead of the kernel: the oracle's vectorised restatement of the reference's
top-k evaluation (src/utils/eva_utils_acc.py - python loops over edges around a sort of 160*160*26 scores each) is
pinned bit-exactly on ranks the reference's own functions produced (tests/golden/eval_ranks.pt)."""
import os

import pytest
import torch

import cases
from oracle import vlsat_oracle as O


@pytest.mark.parametrize("name", list(cases.EVAL_CASES))
def test_oracle_ranks_match_reference_evaluation(name, golden):
    gold = golden("eval_ranks")[name]
    logits, rel, gt_cls, gt_rel, edges = cases.eval_inputs(name)
    assert torch.equal(O.topk_object_ranks(logits, gt_cls, 11), gold["obj"])
    assert torch.equal(O.topk_predicate_ranks(rel, gt_rel, 6), gold["rel"])
    assert torch.equal(O.topk_triplet_ranks(logits, rel, gt_cls, gt_rel, edges, 101), gold["triplet"])
    assert gold["obj"][0] == 1 and gold["obj"][1] == 12                      # best score / beyond top-11
    assert gold["rel"].numel() >= edges.shape[0]                              # one entry per label or per label-free edge


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_oracle_ranks_against_the_live_reference_on_fresh_inputs():
    import numpy as np
    from oracle import ref_shims
    ref_shims.install()
    from src.utils.eva_utils_acc import evaluate_topk_object, evaluate_topk_predicate, evaluate_triplet_topk, get_gt
    g = torch.Generator().manual_seed(77)
    n, e = 9, 25
    logits, rel = torch.randn(n, 160, generator=g) * 2, torch.sigmoid(torch.randn(e, 26, generator=g) * 3)
    gt_cls, gt_rel = torch.randint(0, 160, (n,), generator=g), (torch.rand(e, 26, generator=g) < 0.1).float()
    edges = torch.randint(0, n, (e, 2), generator=g)
    a = evaluate_topk_object(logits, gt_cls, topk=11)
    gt_edges = get_gt(gt_cls, gt_rel, edges, True)
    assert np.array_equal(a, O.topk_object_ranks(logits, gt_cls, 11).numpy())
    assert np.array_equal(evaluate_topk_predicate(rel, gt_edges, True, topk=6), O.topk_predicate_ranks(rel, gt_rel, 6).numpy())
    t = evaluate_triplet_topk(logits, rel, gt_edges, edges, True, topk=101, use_clip=True, obj_topk=a)[0]
    assert np.array_equal(t, O.topk_triplet_ranks(logits, rel, gt_cls, gt_rel, edges, 101).numpy())


def test_eval_rank_mirror_has_no_cpu_fallback():
    from vlsat_b200 import eval_ranks as R
    logits, rel, gt_cls, gt_rel, edges = cases.eval_inputs("eval_small")
    for call in (lambda: R.evaluate_topk_object(logits, gt_cls, 11), lambda: R.evaluate_topk_predicate(rel, gt_rel, 6),
                 lambda: R.evaluate_triplet_topk(logits, rel, gt_cls, gt_rel, edges, 101), lambda: R.softmax_rows(logits)):
        with pytest.raises(TypeError):
            call()


def test_tied_labels_take_the_adjusted_rank_to_zero_and_below():
    """Three ground-truth labels with exactly equal (saturated) scores: the reference emits 1, 0, -1 (eva_utils_acc.py:73-78)."""
    rel = torch.full((2, 26), 0.3)
    rel[0, [1, 2, 3]] = 1.0
    gt_rel = torch.zeros(2, 26)
    gt_rel[0, [1, 2, 3]] = 1
    assert O.topk_predicate_ranks(rel, gt_rel, 6).tolist() == [1, 0, -1, 1]
