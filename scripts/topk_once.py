"""One top-k sparsify call per shape (for an ncu launch list): 50 M single layer, 200 layers x 250 k."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, flashe_b200 as fb
ctx = fb.DeviceContext(bytes(range(32)), 32)
n = 50_000_000
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(n, dtype=torch.float32, device="cuda").normal_(0.0, 0.1, generator=g)
res = torch.zeros(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    ctx.topk_sparsify(x, [n], [n // 100], residual=res, residual_out=res)
    ctx.topk_sparsify(x, np.cumsum([250_000] * 200), [2500] * 200, residual=res, residual_out=res)
torch.cuda.synchronize()
