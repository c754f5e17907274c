import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops, _lib, synth
lib = _lib.load()
lib.vlsat_debug_set_trace.argtypes = [ctypes.c_void_p]; lib.vlsat_debug_set_trace.restype = None
dev = "cuda"
model = V.Mmgnet({"MODEL": V.DEFAULT_MODEL_CONFIG}, 160, 26); synth.load_seeded(model, 0); model = model.to(dev).eval()
b = synth.make_config_batch("cfg2", seed=1).to(dev)
with torch.no_grad():
    for _ in range(2): model.obj_encoder(b.obj_points)
    tr = torch.zeros(2048, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(); lib.vlsat_debug_set_trace(tr.data_ptr())
    model.obj_encoder(b.obj_points); torch.cuda.synchronize(); lib.vlsat_debug_set_trace(None)
t = tr.cpu().tolist(); t0 = t[0]
print("tile: L1_start b1_arrived d2_full b2_arrived d3_full epi3_done   (ns from first stamp)")
for it in range(12):
    print(it, [t[it*8+j]-t0 for j in range(6)])
