#!/usr/bin/env python
"""A launch made of edge items only (every reference chunk shorter than one item): what the general path of k_stream
costs.  Run under ncu (scripts/gpu_ncu_edge.sh) or by itself for the event-timed figure."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

bits = int(os.environ.get("BITS", "20"))
n_jobs, per = 2368, 300
L, n = n_jobs * per, 3
ctx = fb.DeviceContext(bytes(range(32)), bits, "cuda:0")
span = fb.VectorSpan(L, n_jobs)
codec = fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=n)
x = torch.randn(n, L, device="cuda:0") * 0.1
cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
out = torch.empty(L, dtype=torch.float64, device="cuda:0")
for _ in range(4):
    ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, fb.NoiseSpec(seed=7), span, out=cts)
    ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
    ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, fb.NoiseSpec(seed=7), span, out=cts)
b.record()
torch.cuda.synchronize()
print("encode of %d x %d elements in %d chunks (edge items only): %.1f us" % (n, L, n_jobs, a.elapsed_time(b) / 20 * 1e3))
