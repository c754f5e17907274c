"""SURVEY 8(d) "reference baselines beside it", item (i): the reference's PyTorch forward on the SAME B200.

The reference itself is not on the GPU box; its stand-in is the oracle (the plain-PyTorch restatement pinned on the
reference's outputs, tests/test_oracle.py) run eagerly on cuda:0 with PyTorch's stock settings - cuBLAS / cuDNN library
kernels, dense [8, E, E] attention scores exactly as attention.py:41-78 builds them. This is the denominator of
north_star's ">= 10x the reference's 1-GPU PyTorch forward". The oracle is used as the checker / yardstick only (tests/
may import it); nothing of it is on the product path.

The measured numbers are written to gpurun_out/ref_gpu_speedup.json (copied to profiles/ by hand) so that the claim
has a committed source. The assertion itself is loose (>= 3x): a timing test must not turn the parity suite red.
"""
import json
import os

import pytest
import torch

from conftest import ROOT


def _time_ms(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


@pytest.mark.gpu
def test_forward_speedup_over_reference_pytorch_on_the_same_gpu():
    import vlsat_b200 as V
    from vlsat_b200 import synth
    from vlsat_b200.graph import GraphedForward
    from oracle import vlsat_oracle as O

    dev = torch.device("cuda:0")
    model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
    synth.load_seeded(model, 0)
    model = model.to(dev).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    graphed = GraphedForward(model)
    rows = []
    for scenes in (16, 32, 64):
        batch = synth.make_config_batch("cfg2", seed=1, num_scenes=scenes).to(dev)
        args = batch.forward_args()
        with torch.no_grad():
            ours = _time_ms(lambda: graphed(*args), 3, 20)
            got = [t.clone() for t in graphed(*args)]
            try:
                ref = _time_ms(lambda: O.mmgnet_forward(sd, *args, istrain=False), 2, 5)
                want = O.mmgnet_forward(sd, *args, istrain=False)
            except torch.OutOfMemoryError:
                # A9 of the reference materialises [8, E, E] fp32 score tensors (attention.py:55-66): 47 GB each at 64 scenes
                rows.append({"scenes": scenes, "ours_ms": round(ours, 4), "reference_torch_gpu_ms": None,
                             "note": "reference restatement out of memory on this GPU"})
                torch.cuda.empty_cache()
                continue
        # same inputs, same weights: the yardstick computes the same function (sigmoid outputs: rtol 1e-3, atol 1e-5)
        for g, w in zip(got[2:], want[2:]):
            assert torch.allclose(g, w, rtol=1e-3, atol=1e-5)
        del want
        torch.cuda.empty_cache()
        rows.append({"scenes": scenes, "ours_ms": round(ours, 4), "reference_torch_gpu_ms": round(ref, 4),
                     "ours_scenes_per_s": round(scenes / ours * 1e3, 1), "reference_scenes_per_s": round(scenes / ref * 1e3, 1),
                     "speedup": round(ref / ours, 2)})
    report = {"what": "Mmgnet eval forward, cfg2 scene shape (40 obj x 256 pts, 600 edges/scene), fp32, one B200; ours = CUDA-graph "
                      "replay of the C-ABI launches, reference = oracle restatement in eager PyTorch (stock cuBLAS/cuDNN settings) on the same GPU; "
                      "median device time, inputs resident",
              "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": rows}
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "ref_gpu_speedup.json"), "w") as f:
            json.dump(report, f, indent=1)
    except OSError:
        pass
    print(json.dumps(report))
    measured = [r for r in rows if r.get("speedup")]
    assert measured, "the reference restatement ran at no batch size"
    assert min(r["speedup"] for r in measured) >= 3.0, rows
