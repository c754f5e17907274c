import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.vlsat_debug_set_trace.restype = None
dev = "cuda"
g = torch.Generator().manual_seed(0)
flush = torch.empty(64 * 1024 * 1024, device=dev)
import itertools
for (m, n, k), eng in itertools.product([(640, 512, 512), (640, 512, 2048), (9600, 512, 512)], ['tc', 'tc1']):
    ops.set_gemm_engine(eng); print('ENGINE', eng)
    x, w, b = torch.randn(m, k, generator=g).to(dev), torch.randn(n, k, generator=g).to(dev), torch.randn(n, generator=g).to(dev)
    for cold in (False,):
        for _ in range(2): ops.linear(x, w, b, act=1)
        tr = torch.zeros(128 + 1800, dtype=torch.int64, device=dev)
        if cold: flush.zero_()
        torch.cuda.synchronize()
        lib.vlsat_debug_set_trace(tr.data_ptr())
        ops.linear(x, w, b, act=1)
        torch.cuda.synchronize()
        lib.vlsat_debug_set_trace(None)
        t = tr.cpu().tolist()
        t0 = t[0]
        print(f"--- {m}x{n}x{k} cold={cold}: setup {t[1]-t0}, accum_ready {t[2]-t0}, end {t[3]-t0}")
        nct = ((m + 127) // 128) * ((n + 127) // 128)
        ct = [(t[128+3*i], t[128+3*i+1], t[128+3*i+2]) for i in range(min(nct, 600))]
        g0 = min(c[0] for c in ct)
        starts = sorted(c[0]-g0 for c in ct); ends = sorted(c[1]-g0 for c in ct); durs = sorted(c[1]-c[0] for c in ct)
        print(f"CTAs {nct}: start ns min/med/max {starts[0]}/{starts[len(starts)//2]}/{starts[-1]}  end ns min/med/max {ends[0]}/{ends[len(ends)//2]}/{ends[-1]}  dur ns min/med/max {durs[0]}/{durs[len(durs)//2]}/{durs[-1]}  sms {len(set(c[2] for c in ct))}")
        nkb = 0
        print('   ns from CTA(0,0) start: setup_done', t[100]-t[128], 'first_full', t[101]-t[128], 'last_full', t[102]-t[128], 'accum_ready', t[103]-t[128], 'epi_done', t[104]-t[128], 'end', t[105]-t[128])
        print("kb: producer_issue  mma_full_ready  mma_issued  (clk from start)")
        for kb in range(nkb):
            print(kb, t[8+3*kb]-t0, t[8+3*kb+1]-t0, t[8+3*kb+2]-t0)
