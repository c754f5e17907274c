"""CUDA-event timings of the hot kernels at config #2 shapes (no profiler)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth
dev = "cuda"
g = torch.Generator().manual_seed(0)
flush = torch.empty(64 * 1024 * 1024, device=dev)

def timeit(name, fn, flops=0, bytes_=0, n=10):
    """GPU-only time: the call is captured in a CUDA graph so no host overhead is inside the events."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort(); t = ts[len(ts) // 2] * 1e-3
    print(f"{name:44s} {t*1e6:9.1f} us  {flops/t/1e12:8.1f} TFLOP/s  {bytes_/t/1e9:8.1f} GB/s", flush=True)

for (m, n, k) in [(9600, 512, 512), (9600, 1024, 512), (9600, 512, 1024), (640, 512, 512), (640, 2816, 512), (9600, 256, 512)]:
    x, w, b = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev), torch.randn(n, generator=g).to(dev)
    timeit(f"linear {m}x{n}x{k} (split+gemm)", lambda: ops.linear(x, w, b, act=1), 2.0 * m * n * k)
q, k_ = torch.randn(9600, 512, generator=g).to(dev), torch.randn(9600, 512, generator=g).to(dev)
vt = torch.randn(512, 9600, generator=g).to(dev)
timeit("flash_attn_tc 9600x9600 h8 (3 splits + attn)", lambda: ops.flash_attn_tc(q, k_, vt, 9600, 8), 4.0 * 9600 * 9600 * 512)
for sp in ("1", "2", "3", "4", ""):
    os.environ["VLSAT_FLASH_SPLITS"] = sp
    if not sp: del os.environ["VLSAT_FLASH_SPLITS"]
    timeit(f"flash_attn_bf16 9600x9600 h8 kv-splits={sp or 'auto'} (3 operand splits + attn)", lambda: ops.flash_attn_bf16(q, k_, vt, 9600, 8), 4.0 * 9600 * 9600 * 512)
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
synth.load_seeded(model, 0)
model = model.to(dev).eval()
b = synth.make_config_batch("cfg2", seed=1).to(dev)
with torch.no_grad():
    timeit("pointnet 640x256", lambda: model.obj_encoder(b.obj_points), 640 * 256 * 213376.0)
    layer = model.mmg.gcn_3ds[0]
    x = torch.randn(640, 512, device=dev); e = torch.randn(9600, 512, device=dev)
    timeit("gat layer (all kernels)", lambda: layer(x, e, b.edge_indices))
    timeit("full forward", lambda: model(*b.forward_args()))
