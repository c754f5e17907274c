"""Is the GEMM kernel's fixed cost a launch / shared-memory-carveout effect? Back-to-back GEMMs with
pre-split operands (no small kernel in between) vs. GEMMs alternating with a small kernel."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib
from vlsat_b200._lib import Epilogue, LinearOpts
lib = _lib.load()
dev = "cuda"
g = torch.Generator().manual_seed(0)
for (m, n, k) in [(640, 512, 512), (9600, 512, 512)]:
    x, w = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev)
    xh, xl = ops.tf32_split(x); wh, wl = ops.tf32_split(w)
    y = torch.empty(m, n, device=dev)
    epi = Epilogue(); epi.alpha = 1.0
    opts = LinearOpts(); opts.engine = 2
    opts.x_hi, opts.x_lo, opts.w_hi, opts.w_lo = xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr()
    def gemm():
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.vlsat_linear_fwd(x.data_ptr(), k, w.data_ptr(), k, y.data_ptr(), n, m, n, k, C.byref(epi), C.byref(opts), st)
        assert rc == 0
    small = torch.zeros(1024, device=dev)
    for mode in ("back-to-back", "alternating with a tiny kernel"):
        for _ in range(3): gemm()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                gemm()
                if mode != "back-to-back": ops.relu(small)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"{m}x{n}x{k} {mode:32s}: {min(ts) * 1e3 / 20:7.1f} us per GEMM(+tiny)", flush=True)
