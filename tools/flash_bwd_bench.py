"""Time the streaming A9 backward (csrc/flash_attn_bwd.cu) at a given size: CUDA events, after warm-up.
usage: python tools/flash_bwd_bench.py [nq] [nk] [--splits S]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vlsat_b200 as V
from vlsat_b200 import ops

args = [a for a in sys.argv[1:] if not a.startswith("--")]
nq = int(args[0]) if args else 9600
nk = int(args[1]) if len(args) > 1 else nq
H, d = 8, 512
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, dout = (torch.randn(n, d, device="cuda", generator=g) for n in (nq, nk, nk, nq))
vt = ops.transpose(v)
out, lse = ops.flash_attn_bf16(q, k, vt, nk, H, want_lse=True)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {"nq": nq, "nk": nk}
res["fwd_ms"] = timeit(lambda: ops.flash_attn_bf16(q, k, vt, nk, H, want_lse=True))
res["bwd_total_ms"] = timeit(lambda: ops.flash_attn_bf16_bwd(q, k, v, dout, out, lse, H))
# the tensor-core launches alone (operands prepared once)
qp, qt = ops.bf16_split_t(q); kp, kt = ops.bf16_split_t(k); vp, _ = ops.bf16_split_t(v, want_t=False); dop, dot = ops.bf16_split_t(dout)
res["prep_ms"] = timeit(lambda: (ops.bf16_split_t(q), ops.bf16_split_t(k), ops.bf16_split_t(v, want_t=False), ops.bf16_split_t(dout),
                                 ops.flash_attn_bwd_stats(dout, out, lse, H)))
flops = 14.0 * nq * nk * d
res["bwd_tflops_fp32_eq"] = flops / (res["bwd_total_ms"] - res["prep_ms"]) / 1e9
print(json.dumps(res))
