#!/usr/bin/env python
"""The online step of the precomputed-mask schedule by itself: flashe_encode_add_premasked_batch on n clients x L
elements (12 B per element), CUDA-event timed; A/B of library variants through FLASHE_B200_LIB."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

L, n = 100_000_000, int(os.environ.get("CLIENTS", "16"))
ctx = fb.DeviceContext(bytes(range(32)), 32, "cuda:0")
span = fb.VectorSpan(L, 16)
codec = fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=n)
x = torch.randn(n, L, device="cuda:0") * 0.1
masks = torch.randint(0, 2 ** 31, (n, L), device="cuda:0", dtype=torch.int32).view(torch.uint32)
ct = ctx.empty_words(L, rows=n)
out = {}
for res in (53, 32):
    noise = fb.NoiseSpec(seed=1, stream=0, resolution=res)
    f = lambda: ctx.encode_add_premasked_batch(x, codec, noise, masks, span, out=ct)  # noqa: E731
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        f()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out["res%d" % res] = {"ms": ms, "gbs": n * L * 12 / ms / 1e6, "frac_of_6548": n * L * 12 / ms / 1e6 / 6548.2}
print(json.dumps({"lib": os.environ.get("FLASHE_B200_LIB", "default").split("_")[-1], "clients": n, **out}))
