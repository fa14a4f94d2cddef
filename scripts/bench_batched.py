#!/usr/bin/env python
"""Fused lane-batched round (int_bits 120, 25M elements x 10 clients) timed alone: one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

L, n, bits, e = 25_000_000, 10, 120, 16
dev = torch.device("cuda", 0)
ctx = fb.DeviceContext(bytes(range(32)), bits, dev)
f = (n - 1).bit_length()
bs = bits // (e + f)
nw = -(-L // bs)
span = fb.VectorSpan(nw, os.cpu_count() or 16)
codec = fb.CodecSpec(alpha=[0.5938345], element_bits=e, n_clients=n, seg_end=[L], batch_lane_bits=e + f)
x = torch.randn(n, L, device=dev) * 0.1
cts, agg = ctx.empty_words(nw, rows=n), ctx.empty_words(nw)
out = torch.empty(L, dtype=torch.float64, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]


def rnd(rec=False):
    if rec: ev[0].record()
    ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, fb.NoiseSpec(seed=7, stream=0), span, out=cts)
    if rec: ev[1].record()
    ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
    if rec: ev[2].record()
    ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
    if rec: ev[3].record()


for _ in range(3):
    rnd()
torch.cuda.synchronize()
tot = [0.0, 0.0, 0.0]
steps = 8
for _ in range(steps):
    rnd(True)
    torch.cuda.synchronize()
    for i in range(3):
        tot[i] += ev[i].elapsed_time(ev[i + 1]) / steps
print(json.dumps({"round_ms": sum(tot), "encode_ms": tot[0], "aggregate_ms": tot[1], "decrypt_decode_ms": tot[2],
                  "encode_g_blocks_per_s": 2 * n * nw / (tot[0] * 1e-3) / 1e9}))
