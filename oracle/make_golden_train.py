"""TEST INFRASTRUCTURE: training-step fixtures from the UNMODIFIED reference -> tests/golden/train_step.pt.

Run in the build container only (needs /root/reference):  python oracle/make_golden_train.py
For each case the reference's own ``Mmgnet.process_train`` (SGFN_MMG/model.py:337-418) runs TRAIN_STEPS times on the
seeded batch / supervision of tests/golden/cases.py - its loss code, its ``backward`` (loss.backward, AdamW over its 13
parameter groups, zero_grad, CosineAnnealingLR) - with two stand-ins only:
  * ``get_rel_emb`` (CLIP text encoder, needs the CLIP weights) returns the seeded text tensor of cases.train_targets;
  * the model stays in eval() mode, as ref_shims.build_reference_mmgnet creates it (dropout masks cannot be matched
    across RNGs; BatchNorm uses its running statistics). process_train itself never switches the mode.
The metric code after ``self.backward(loss)`` (:415-456) is skipped by raising from the wrapped ``backward``.
Stored per case: the seven differentiable forward outputs of step 1, the loss of every step, and for every parameter a
summary (cases.grad_summary) of its total change after the last step.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "train_step.pt")


class _StepDone(Exception):
    pass


def main():
    ref_shims.install()
    out = {}
    for name in cases.TRAIN_CASES:
        over, make = cases.MMGNET_CASES[name]
        net, cfg = ref_shims.build_reference_mmgnet(seed=0, overrides=over)
        net.load_state_dict(cases.seeded_state(net, cases.MMGNET_WEIGHT_SEED))
        net.eval()
        if not hasattr(net, "iteration"):
            net.iteration = 0
        b = make()
        gt_cls, gt_rel, text = cases.train_targets(b)
        before = {k: p.detach().clone() for k, p in net.named_parameters()}
        rec = dict(losses=[], outs=None)

        net.get_rel_emb = lambda *a, **k: text
        fwd = net.forward

        def forward(*a, **k):
            o = fwd(*a, **k)
            if rec["outs"] is None:
                rec["outs"] = [t.detach().clone() for t in o[:7]]
            return o
        net.forward = forward
        real_backward = net.backward

        def backward(loss):
            rec["losses"].append(float(loss.detach()))
            real_backward(loss)
            raise _StepDone()
        net.backward = backward

        obj_points, obj_2d, edge_index, descriptor, batch_ids = b.forward_args()
        for _ in range(cases.TRAIN_STEPS):
            try:
                net.process_train(obj_points, obj_2d, gt_cls, descriptor, gt_rel, edge_index.t().contiguous(), batch_ids)
            except _StepDone:
                pass
        delta = {k: cases.grad_summary(p.detach() - before[k]) for k, p in net.named_parameters()
                 if not torch.equal(p.detach(), before[k])}
        out[name] = dict(outs=rec["outs"], losses=rec["losses"], delta=delta,
                         lr=float(cfg.LR), t_max=int(cfg.max_iteration),
                         last_lr=[g["lr"] for g in net.optimizer.param_groups])
        print(name, "losses", rec["losses"], "changed parameters", len(delta))
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
