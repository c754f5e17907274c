"""TEST INFRASTRUCTURE: evaluation-rank fixtures from the reference's own functions -> tests/golden/eval_ranks.pt.

Run in the build container only (needs /root/reference):  python oracle/make_golden_eval.py
``evaluate_topk_object`` (topk 11), ``get_gt`` + ``evaluate_topk_predicate`` (topk 6) and ``evaluate_triplet_topk``
(topk 101, use_clip=True) of src/utils/eva_utils_acc.py, called the way ``Mmgnet.process_val`` calls them
(SGFN_MMG/model.py:463-472), on the seeded predictions of tests/golden/cases.py."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "eval_ranks.pt")


def main():
    ref_shims.install()
    from src.utils.eva_utils_acc import evaluate_topk_object, evaluate_topk_predicate, evaluate_triplet_topk, get_gt
    out = {}
    for name in cases.EVAL_CASES:
        logits, rel, gt_cls, gt_rel, edges = cases.eval_inputs(name)
        top_k_obj = evaluate_topk_object(logits, gt_cls, topk=11)
        gt_edges = get_gt(gt_cls, gt_rel, edges, True)
        top_k_rel = evaluate_topk_predicate(rel, gt_edges, True, topk=6)
        top_k_triplet = evaluate_triplet_topk(logits, rel, gt_edges, edges, True, topk=101, use_clip=True, obj_topk=top_k_obj)[0]
        out[name] = dict(obj=torch.from_numpy(top_k_obj).long(), rel=torch.from_numpy(top_k_rel).long(),
                         triplet=torch.from_numpy(top_k_triplet).long())
        print(name, {k: tuple(v.shape) for k, v in out[name].items()})
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
