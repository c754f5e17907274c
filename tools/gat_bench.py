"""Device time of the tensor-core GAT edge kernel (vlsat_gat_edge_tc_fwd) alone, at a BASELINE config shape:
20 calls captured in one CUDA graph (no host gaps), L2-warm. Profiling aid; run under ncu for per-kernel durations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
N, E = {"cfg2": (640, 9600), "cfg3": (2560, 99840), "cfg4": (1280, 19200)}[cfg]
H, de, hid, do = 8, 64, 128, 32
dev = "cuda"
g = torch.Generator().manual_seed(0)
src = torch.sort(torch.randint(0, N, (E,), generator=g)).values.to(dev)
dst = torch.randint(0, N, (E,), generator=g).to(dev)
k = torch.randn(E, H * de, generator=g).to(dev)
kp = ops.bf16_split(k)
qc = torch.randn(N, H * hid, generator=g).to(dev)
v = torch.randn(N, H * do, generator=g).to(dev)
c1k = ops.bf16_split((torch.randn(hid, de, generator=g) / 8).to(dev))
c2 = ops.bf16_split((torch.randn(do, hid, generator=g) / 11).to(dev))
c2b = torch.randn(do, generator=g).to(dev)
out = torch.empty(N, H * do, device=dev)

def call():
    ops.gat_edge_tc(kp, qc, v, src, dst, c1k, c2, c2b, N, H, out, d_n=64)

for _ in range(3):
    call()
torch.cuda.synchronize()
reps = 20
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(reps):
        call()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gr.replay(); torch.cuda.synchronize()
a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / reps * 1e3
by = E * (H * de * 4.0 + 16.0) + N * (H * 64 + 2.0 * H * do) * 4.0
print(f"{cfg}: {us:.1f} us per call (edge kernel + finalize), algorithmic {by / 1e6:.2f} MB -> {by / us * 1e-3:.0f} GB/s")

# per-phase trace of CTA 0 / warpgroup 0 (ns since the first stamp): tile start, acc1 ready, hidden written, acc2 ready,
# messages in smem, segmented max done
import ctypes
from vlsat_b200 import _lib
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]; lib.vlsat_debug_set_trace.restype = None
tr = torch.zeros(512, dtype=torch.int64, device=dev)
torch.cuda.synchronize(); lib.vlsat_debug_set_trace(tr.data_ptr()); call(); torch.cuda.synchronize(); lib.vlsat_debug_set_trace(None)
tl = tr.cpu().tolist()
for j in range(8):
    row = tl[j * 12:(j + 1) * 12]
    if row[0]:
        print("tile", j, [x - row[0] if x else 0 for x in row])
