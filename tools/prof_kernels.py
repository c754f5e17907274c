"""Run each hot kernel of the path a few times at BASELINE config #2 shapes (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, synth

which = sys.argv[1:] or ["linear", "flash", "gat", "pointnet"]
dev = "cuda"
g = torch.Generator().manual_seed(0)
n_iter = 3
if "linear" in which:
    for (m, n, k) in [(9600, 512, 512), (640, 512, 512), (9600, 1024, 512)]:
        x, w, b = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev), torch.randn(n, generator=g).to(dev)
        for _ in range(n_iter):
            ops.linear(x, w, b, act=1)
if "flash" in which:
    q, k = torch.randn(9600, 512, generator=g).to(dev), torch.randn(9600, 512, generator=g).to(dev)
    vt = torch.randn(512, 9600, generator=g).to(dev)
    for _ in range(n_iter):
        ops.flash_attn_tc(q, k, vt, 9600, 8)
if "gat" in which or "pointnet" in which:
    model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26)
    synth.load_seeded(model, 0)
    model = model.to(dev).eval()
    b = synth.make_config_batch("cfg2", seed=1).to(dev)
    with torch.no_grad():
        for _ in range(n_iter):
            if "pointnet" in which:
                model.obj_encoder(b.obj_points)
            if "gat" in which:
                layer = model.mmg.gcn_3ds[0]
                x = torch.randn(640, 512, device=dev); e = torch.randn(9600, 512, device=dev)
                layer(x, e, b.edge_indices)
if "flash_fwd" in which:
    q, k, v = (torch.randn(9600, 512, generator=g).to(dev) for _ in range(3))
    (qp, _), (kp, _), (_, vt) = ops.bf16_split_t(q), ops.bf16_split_t(k), ops.bf16_split_t(v)
    for _ in range(2):
        ops.flash_attn_bf16(qp, kp, vt, 9600, 8)
if "flash_fwd_bf16" in which:                       # single-pass bf16 mode: which pipe fills up
    ops.set_precision("bf16")
    q, k, v = (torch.randn(9600, 512, generator=g).to(dev) for _ in range(3))
    (qp, _), (kp, _), (_, vt) = ops.bf16_split_t(q), ops.bf16_split_t(k), ops.bf16_split_t(v)
    for _ in range(2):
        ops.flash_attn_bf16(qp, kp, vt, 9600, 8)
    ops.set_precision("fp32")
if "flash_bwd" in which:
    q, k, v, dout = (torch.randn(9600, 512, generator=g).to(dev) for _ in range(4))
    (qp, qt), (kp, kt), (vp, vt) = ops.bf16_split_t(q), ops.bf16_split_t(k), ops.bf16_split_t(v)
    out, lse = ops.flash_attn_bf16(qp, kp, vt, 9600, 8, want_lse=True)
    for _ in range(2):
        ops.flash_attn_bf16_bwd(q, k, v, dout, out, lse, 8, prep=(qp, qt, kp, kt, vp))
if "gemm_pairs" in which:
    dz, x, w = torch.randn(9600, 1024, generator=g).to(dev), torch.randn(9600, 512, generator=g).to(dev), torch.randn(1024, 512, generator=g).to(dev)
    dzp, xp, wp = ops.bf16_split(dz), ops.bf16_split(x), ops.bf16_split(w)
    for _ in range(2):
        ops.gemm_nn(dzp, wp, 512)            # dX = dZ W      [9600, 512]
        ops.gemm_tn(dzp, xp, 1024, 512)      # dW = dZ^T X    [1024, 512], reduction 9600 split over CTAs
    t, h = torch.randn(76800, 128, generator=g).to(dev), torch.randn(76800, 128, generator=g).to(dev)
    tp, hp = ops.bf16_split(t), ops.bf16_split(h)
    for _ in range(2):
        ops.gemm_tn(tp, hp, 128, 128)        # attention-MLP weight gradient: one output tile, reduction 76,800
torch.cuda.synchronize()
print("done")
