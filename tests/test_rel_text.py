"""N4: the text supervision target of ``process_train``. The oracle restatement is pinned on the reference's own
``get_rel_emb`` (SGFN_MMG/model.py:221-255) run unmodified around a stand-in text encoder (oracle/make_golden_text.py);
the kernel behind ``RelTextCache`` is checked against the same fixtures on the GPU."""
import pytest
import torch

import cases
from oracle import vlsat_oracle as O


@pytest.mark.parametrize("name", list(cases.TEXT_CASES))
def test_oracle_rel_text_embed_matches_the_reference_get_rel_emb(name, golden):
    n_obj_cls, n_rel_cls, _, _, seed = cases.TEXT_CASES[name]
    gt_cls, gt_rel, edges = cases.text_inputs(name)
    got = O.rel_text_embed(cases.text_table(n_obj_cls, n_rel_cls, seed), gt_cls, gt_rel, edges)
    want = golden("rel_text")[name]
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert (gt_rel.sum(1) == 0).any() and (gt_rel.sum(1) > 1).any()          # label-free and multi-label edges are covered


def test_prompt_table_order_matches_the_reference_prompts():
    """RelTextCache.prompts builds the strings of get_rel_emb (:232-240) in table order; fill() routes them through any
    encoder. Host logic only: a fake encoder returns the prompt's index."""
    import vlsat_b200  # noqa: F401
    from vlsat_b200.train_glue import RelTextCache
    objs, rels = ["chair", "table", "sofa"], ["standing on", "left of"]
    cache = RelTextCache(objs, rels, dim=256, device="cpu")
    p = cache.prompts(1)
    assert len(p) == 3 * 3 and p[0] == "a point cloud of a table standing on a chair" and p[1] == "a point cloud of a table left of a chair"
    assert p[2] == "the table and the chair has no relation in the point cloud" and p[-1] == "the table and the sofa has no relation in the point cloud"
    seen = []

    def encode(prompts):
        seen.extend(prompts)
        base = len(seen) - len(prompts)
        return torch.arange(base, base + len(prompts), dtype=torch.float32).view(-1, 1).expand(-1, 256)
    cache.fill(encode, batch=4)
    assert len(seen) == 27 and cache.filled
    assert float(cache.table[2, 1, 0, 0]) == seen.index("a point cloud of a sofa standing on a table")
    assert float(cache.table[0, 2, 2, 5]) == seen.index("the chair and the sofa has no relation in the point cloud")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.TEXT_CASES))
def test_rel_text_cache_kernel_matches_the_reference_fixture(name, golden):
    from vlsat_b200.train_glue import RelTextCache
    n_obj_cls, n_rel_cls, _, _, seed = cases.TEXT_CASES[name]
    gt_cls, gt_rel, edges = cases.text_inputs(name)
    cache = RelTextCache.from_table(cases.text_table(n_obj_cls, n_rel_cls, seed).cuda())
    got = cache(gt_cls.cuda(), gt_rel.cuda(), edges.cuda())
    want = golden("rel_text")[name]
    assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6)
    assert cache(gt_cls.cuda(), gt_rel[:0].cuda(), edges[:0].cuda()).shape == (0, 512)
