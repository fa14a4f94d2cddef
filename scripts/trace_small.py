#!/usr/bin/env python
"""Per-warp timeline of the stream kernels of a small dense round (tuning build only):
    scripts/build_variant.sh trace -DFLASHE_TRACE=1
    FLASHE_B200_LIB=$PWD/flashe_b200/_lib/libflashe_b200_trace.so FLASHE_TRACE_PRINT=1 python scripts/trace_small.py
Every k_stream launch prints one summary on stderr (span, prologue, percentiles of the warps' end times, edge items)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402


def case(L, n, n_jobs, bits=20, reps=3):
    print("== L %d clients %d n_jobs %d int_bits %d" % (L, n, n_jobs, bits), file=sys.stderr, flush=True)
    ctx = fb.DeviceContext(bytes(range(32)), bits, "cuda:0")
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=n)
    x = torch.randn(n, L, device="cuda:0") * 0.1
    cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
    out = torch.empty(L, dtype=torch.float64, device="cuda:0")
    for _ in range(reps):
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, fb.NoiseSpec(seed=7), span, out=cts)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)
    torch.cuda.synchronize()
    ctx.close()


if __name__ == "__main__":
    case(1_000_000, 3, 16)
    case(995_328, 3, 16)
    case(2_500_000, 10, 16)
    case(2_494_464, 10, 16)
    case(25_000_000, 10, 16, reps=2)
    case(2_500_000, 10, 16, bits=32)
