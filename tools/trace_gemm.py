"""globaltimer trace of CTA 0 of the persistent GEMM: per tile [mma_start, mma_issued_all, accum_ready, epilogue_done]."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]; lib.vlsat_debug_set_trace.restype = None
dev = "cuda"
g = torch.Generator().manual_seed(0)
def run(label, fn):
    for _ in range(2): fn()
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); lib.vlsat_debug_set_trace(tr.data_ptr()); fn(); torch.cuda.synchronize(); lib.vlsat_debug_set_trace(None)
    t = tr.cpu().tolist(); t0 = t[0]
    print("---", label)
    for i in range(6):
        if t[i*4]: print(i, [t[i*4+j]-t0 for j in range(4)])
m, n, k = 9600, 1024, 512
x, w, b = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev), torch.randn(n, generator=g).to(dev)
xs = ops.tf32_split(x)
run("9600x1024x512 plain", lambda: ops.linear(x, w, b, act=1, x_split=xs))
ga = torch.randn(640, 2 * n, generator=g).to(dev); ia = torch.randint(0, 640, (m,), generator=g).to(dev); ib = torch.randint(0, 640, (m,), generator=g).to(dev)
run("9600x1024x512 gather+emit", lambda: ops.linear(x, w, b, act=1, x_split=xs, gather=(ga[:, :n], ia, ga[:, n:], ib), emit_split=True, want_y=False))
res = torch.randn(m, 512, generator=g).to(dev); w2 = torch.randn(512, k, generator=g).to(dev); b2 = b[:512].contiguous()
run("9600x512x512 residual", lambda: ops.linear(x, w2, b2, x_split=xs, residual=res))
run("9600x512x512 plain", lambda: ops.linear(x, w2, b2, x_split=xs))
