"""TEST INFRASTRUCTURE: fixtures for row N4 (CLIP text supervision) from the UNMODIFIED reference -> tests/golden/rel_text.pt.

Run in the build container only (needs /root/reference):  python oracle/make_golden_text.py
The reference's own ``Mmgnet.get_rel_emb`` (SGFN_MMG/model.py:221-255) runs unmodified: its per-edge python loop builds the
prompts, "encodes" them, averages the features of a multi-label edge and normalises. Two stand-ins only, because the CLIP
weights are not available offline:
  * ``clip.tokenize(prompt)`` returns the row index of that prompt in a seeded table (one row per (subject class, object
    class, predicate | "no relation") - exactly the set of prompts get_rel_emb can ever build);
  * ``self.clip_model.encode_text(tokens)`` looks those rows up.
So everything the reference does AROUND the text encoder is pinned: prompt selection per edge, the multi-label mean, the
normalisation, the order of the output rows. ``Tensor.cuda`` is a no-op on this CPU-only host (ref_shims).
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref_shims  # noqa: E402
import cases  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rel_text.pt")


def main():
    ref_shims.install()
    import src.model.SGFN_MMG.model as M
    out = {}
    for name, (n_obj_cls, n_rel_cls, n_nodes, n_edges, seed) in cases.TEXT_CASES.items():
        obj_names = [f"object{i}" for i in range(n_obj_cls)]
        rel_names = [f"relation {i} of" for i in range(n_rel_cls)]
        table = cases.text_table(n_obj_cls, n_rel_cls, seed)                   # [S, O, R + 1, 512]
        index = {}
        for s, a in enumerate(obj_names):
            for o, b in enumerate(obj_names):
                for r, rel in enumerate(rel_names):
                    index[f"a point cloud of a {a} {rel} a {b}"] = (s * n_obj_cls + o) * (n_rel_cls + 1) + r
                index[f"the {a} and the {b} has no relation in the point cloud"] = (s * n_obj_cls + o) * (n_rel_cls + 1) + n_rel_cls
        flat = table.reshape(-1, table.shape[-1])
        fake_clip = types.SimpleNamespace(tokenize=lambda p: torch.tensor([[index[p]]]))
        old_clip = M.clip
        M.clip = fake_clip
        try:
            me = types.SimpleNamespace(obj_label_list=obj_names, rel_label_list=rel_names,
                                       clip_model=types.SimpleNamespace(encode_text=lambda tok: flat[tok.view(-1)]))
            gt_cls, gt_rel, edges = cases.text_inputs(name)
            feats = M.Mmgnet.get_rel_emb(me, gt_cls, gt_rel, edges)              # the reference's code, unmodified
        finally:
            M.clip = old_clip
        out[name] = feats.clone()
        print(name, tuple(feats.shape), float(feats.norm(dim=-1).mean()))
    torch.save(out, OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
