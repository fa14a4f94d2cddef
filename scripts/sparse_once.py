#!/usr/bin/env python
"""The server / client sparse operations of BASELINE config 4 (top-1 % of 50 M, 32 clients), a few iterations each: run
under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_sparse|k_stream"` for
the per-kernel times and traffic of the tiled kernels (scripts/gpu_ncu_sparse.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

total, n, bits, n_jobs = 50_000_000, 32, 32, 16
k = total // 100
ctx = fb.DeviceContext(bytes(range(32)), bits, "cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(3)
base = torch.sort(torch.randperm(total, device="cuda:0", generator=g)[:k]).values
idxs = [torch.sort((base + 977 * c) % total).values.contiguous() for c in range(n)]
ct = torch.randint(0, 2 ** 31, (k,), device="cuda:0", dtype=torch.int64).to(torch.uint32)
fused = ctx.empty_words(total)
for _ in range(3):
    ctx.sparse_sum([ct] * n, idxs, total, [32768] * n, out=fused)
    ctx.sparse_apply_masks_batch(0, list(range(n)), -1, n_jobs, idxs, fused, validate=False)
    ctx.sparse_overlap(idxs, total)
torch.cuda.synchronize()
print("ok")
