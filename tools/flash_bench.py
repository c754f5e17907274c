"""Device time of the BF16x3 streaming attention alone at the cfg2 shape (9600 x 9600 x 8 heads), graph-replayed."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 9600
dev = "cuda"; g = torch.Generator().manual_seed(0)
q, k = torch.randn(n, 512, generator=g).to(dev), torch.randn(n, 512, generator=g).to(dev)
vt = torch.randn(512, n, generator=g).to(dev)
qp, kp, vp = ops.bf16_split(q), ops.bf16_split(k), ops.bf16_split(vt)
fn = lambda: ops.flash_attn_bf16(qp, kp, vp, n, 8)
for _ in range(3): fn()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(5): fn()
gr.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 5 * 1e3
print(f"n={n}: {us:.1f} us per call, {4.0 * n * n * 512 / us * 1e-6:.1f} TFLOP/s fp32-eq")
