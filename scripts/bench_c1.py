#!/usr/bin/env python
"""BASELINE config 1 (1M elements x 3 clients, int_bits 20) round, a few iterations: run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel device times of a small round."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

L, n, bits = 1_000_000, 3, 20
ctx = fb.DeviceContext(bytes(range(32)), bits, "cuda:0")
span = fb.VectorSpan(L, os.cpu_count() or 16)
codec = fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=n)
x = torch.randn(n, L, device="cuda:0") * 0.1
cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
out = torch.empty(L, dtype=torch.float64, device="cuda:0")
for _ in range(8):
    ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, fb.NoiseSpec(seed=7), span, out=cts)
    ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
    ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
torch.cuda.synchronize()
print("ok")
