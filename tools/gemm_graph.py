"""Device time per call of dense projections replayed from a CUDA graph (no host gaps): shapes of one cfg2 forward.
For each shape: x given as fp32 (the C entry splits it first: 2 kernels) and x pre-split (1 kernel). Profiling aid."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)

def graph_time(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

shapes = [(640, 512, 512), (640, 1536, 512), (640, 3328, 512), (640, 768, 768), (640, 512, 768), (640, 504, 768), (640, 160, 512),
          (9600, 512, 512), (9600, 1024, 512), (9600, 512, 1024), (9600, 256, 512), (9600, 26, 256), (9600, 128, 64), (9600, 512, 128)]
for m, n, k in shapes:
    x, w, b = torch.randn(m, k, generator=g).to(dev), (torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.randn(n, generator=g).to(dev)
    xs = ops.split_pair(x)
    t_raw = graph_time(lambda: ops.linear(x, w, b))
    t_pre = graph_time(lambda: ops.linear(x, w, b, x_split=xs))
    print(f"{m:5d}x{n:5d}x{k:5d}   fp32 x: {t_raw:6.1f} us   pre-split x: {t_pre:6.1f} us   {2.0 * m * n * k / t_pre * 1e-6:6.1f} TFLOP/s")
