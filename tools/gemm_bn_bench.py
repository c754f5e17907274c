"""Back-to-back time of the projection kernel at the large shapes of the path (exactly two tile rounds: M = 9472; the path's own M = 9600).
Round 2 used it to compare the default 128x128 tiles with an experimental 128x256 / two-stage instantiation (dropped: DESIGN.md section 8 item 4).
Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops
dev = "cuda"; g = torch.Generator().manual_seed(0)
for (m, n, k) in [(9472, 512, 512), (9600, 512, 512), (9472, 1024, 512), (9600, 1024, 512), (9600, 512, 1024), (18944, 512, 512)]:
    x, w, b = torch.randn(m, k, generator=g).to(dev), (torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.randn(n, generator=g).to(dev)
    xs, ws = ops.split_pair(x), ops.split_pair(w)
    y = torch.empty(m, n, device=dev)
    f = lambda: ops.linear(x, w, b, act=1, x_split=xs, w_split=ws, out=y)
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): f()
    c.record(); torch.cuda.synchronize()
    us = a.elapsed_time(c) / 50 * 1e3
    ref = torch.relu(x.double() @ w.double().t() + b.double())
    err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"{m}x{n}x{k}: {us:7.2f} us  {2*m*n*k/us*1e-6:7.1f} TFLOP/s fp32-eq  max err/scale {err:.2e}", flush=True)
