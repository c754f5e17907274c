import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]; lib.vlsat_debug_set_trace.restype = None
dev = "cuda"; g = torch.Generator().manual_seed(0)
for m, n, k in [(640, 512, 512), (640, 3328, 512), (640, 768, 768), (9600, 512, 512)]:
    x, w, b = torch.randn(m, k, generator=g).to(dev), (torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.randn(n, generator=g).to(dev)
    xs = ops.split_pair(x)
    fn = lambda: ops.linear(x, w, b, x_split=xs)
    for _ in range(3): fn()
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); lib.vlsat_debug_set_trace(tr.data_ptr()); e0.record(); fn(); e1.record(); torch.cuda.synchronize(); lib.vlsat_debug_set_trace(None)
    t = tr.cpu().tolist()
    print(m, n, k, "event us", round(e0.elapsed_time(e1) * 1e3, 1), [[t[i * 4 + j] - t[0] for j in range(4)] for i in range(3) if t[i * 4]])
