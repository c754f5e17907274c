"""One large dense projection (9600 x 512 x 512, pre-split operands) a few times: target of a single-kernel ncu capture."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vlsat_b200 as V
from vlsat_b200 import ops
dev = "cuda"; g = torch.Generator().manual_seed(0)
m, n, k = 9600, 512, 512
x, w, b = torch.randn(m, k, generator=g).to(dev), (torch.randn(n, k, generator=g) / k ** 0.5).to(dev), torch.randn(n, generator=g).to(dev)
xs = ops.split_pair(x)
for _ in range(4):
    ops.linear(x, w, b, act=1, x_split=xs)
torch.cuda.synchronize()
print("done")
