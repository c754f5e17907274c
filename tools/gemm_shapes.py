"""Per-shape timing of the dense-projection engine on the shapes of one cfg2 forward (CUDA events, L2-warm, 50 reps) and
the globaltimer trace of CTA 0 (per tile: mma start, mma issued, accumulator ready, epilogue done). Profiling aid."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]; lib.vlsat_debug_set_trace.restype = None
dev = "cuda"
g = torch.Generator().manual_seed(0)
eng = sys.argv[1] if len(sys.argv) > 1 else "auto"
ops.set_gemm_engine(eng)

def timeit(fn, reps=50):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

def trace(fn):
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); lib.vlsat_debug_set_trace(tr.data_ptr()); fn(); torch.cuda.synchronize(); lib.vlsat_debug_set_trace(None)
    t = tr.cpu().tolist(); t0 = t[0]
    return [[t[i * 4 + j] - t0 for j in range(4)] for i in range(6) if t[i * 4]]

shapes = [("nn_edge.0 gather+emit", 9600, 1024, 512, "ge"), ("nn_edge.2 emit", 9600, 512, 1024, "e"), ("proj_edge emit-only", 9600, 512, 512, "eo"),
          ("xattn q emit-only", 9600, 512, 512, "eo"), ("xattn vT", 512, 9600, 512, "t"), ("fc_o residual", 9600, 512, 512, "r"),
          ("rel fc1 relu", 9600, 512, 512, ""), ("rel fc2", 9600, 256, 512, ""), ("rel fc3 sigmoid", 9600, 26, 256, ""),
          ("relenc l2", 9600, 128, 64, ""), ("relenc l3", 9600, 512, 128, ""), ("node proj", 640, 3328, 512, ""), ("node qkv", 640, 1536, 512, ""),
          ("node 512", 640, 512, 512, ""), ("prop0", 640, 768, 768, ""), ("prop2", 640, 512, 768, "")]
tot = 0.0
for name, m, n, k, kind in shapes:
    x, w, b = torch.randn(m, k, generator=g).to(dev), (torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.randn(n, generator=g).to(dev)
    xs = ops.split_pair(x)
    kw = dict(x_split=xs)
    if "g" in kind:
        ga = torch.randn(640, 2 * n, generator=g).to(dev); ia = torch.randint(0, 640, (m,), generator=g).to(dev); ib = torch.randint(0, 640, (m,), generator=g).to(dev)
        kw.update(gather=(ga[:, :n], ia, ga[:, n:], ib), act=1)
    if "e" in kind:
        kw.update(emit_split=True)
    if "o" in kind:
        kw.update(want_y=False)
    if "r" in kind:
        kw.update(residual=torch.randn(m, n, generator=g).to(dev))
    if "t" in kind:
        fn = lambda: ops.linear(x, w, None, x_is_weight=True, w_split=ops.split_pair(w) if False else None)
        ws = ops.split_pair(w)
        fn = lambda: ops.linear(x, w, None, x_is_weight=True, w_split=ws, emit_split="bf16" if eng in ("auto", "bf16x3") else "tf32", want_y=False)
    else:
        fn = lambda: ops.linear(x, w, b, **kw)
    us = timeit(fn)
    fl = 2.0 * m * n * k
    print(f"{name:24s} {m:5d}x{n:5d}x{k:5d}  {us:8.1f} us  {fl / us * 1e-6:7.1f} TFLOP/s fp32-eq")
    if m * n >= 9600 * 512:
        for row in trace(fn)[:4]:
            print("      tile trace (ns): mma_start %6d issued %6d acc_ready %6d epi_done %6d" % tuple(row))
