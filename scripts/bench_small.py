#!/usr/bin/env python
"""Where a SMALL dense round spends its time: each of the three kernels timed by itself (CUDA events over `reps`
back-to-back calls) for a sweep of (elements, clients, n_jobs).  Layouts whose chunks are whole items (chunk length a
multiple of 64 blocks and chunk begins on multiples of 64 counters) have no edge items; comparing them with their
neighbours separates the cost of the edge items from the fixed cost of a launch.
    python scripts/bench_small.py [--bits 20] > gpurun_out/small.jsonl"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

KEY = bytes(range(32))


def timed(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def case(L, n, n_jobs, bits, dev, reps, tag=""):
    ctx = fb.DeviceContext(KEY, bits, dev)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=7, stream=0)
    x = torch.randn(n, L, device=dev) * 0.1
    cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
    out = torch.empty(L, dtype=torch.float64, device=dev)
    enc = lambda: ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, noise, span, out=cts)  # noqa: E731
    ag = lambda: ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)  # noqa: E731
    dec = lambda: ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)  # noqa: E731

    def rnd():
        enc(); ag(); dec()

    g_enc, g_dec, g_rnd = ctx.capture(enc), ctx.capture(dec), ctx.capture(rnd)
    m = 128 // bits
    blocks = -(-L // m)
    r = {"tag": tag, "elements": L, "clients": n, "n_jobs": n_jobs, "int_bits": bits,
         "encode_us": timed(g_enc, reps), "aggregate_us": timed(ag, reps), "decode_us": timed(g_dec, reps),
         "round_graph_us": timed(g_rnd, reps)}
    r["encode_g_blocks_s"] = 2 * n * blocks / r["encode_us"] / 1e3
    r["decode_g_blocks_s"] = 2 * blocks / r["decode_us"] / 1e3
    print(json.dumps(r), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=20)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    dev = "cuda:0"
    m = 128 // a.bits
    item = 64 * m
    if a.quick:
        case(1_000_000, 3, 16, a.bits, dev, a.reps, "C1")
        case(2_500_000, 10, 16, a.bits, dev, a.reps, "C2")
        return
    # C1 and C2 as they are, and with whole-item chunks (no edge items)
    case(1_000_000, 3, 16, a.bits, dev, a.reps, "C1")
    case(16 * item * (1_000_000 // (16 * item)), 3, 16, a.bits, dev, a.reps, "C1 whole-item chunks")
    case(1_000_000, 3, 1, a.bits, dev, a.reps, "C1 one chunk")
    case(1_000_000, 3, 64, a.bits, dev, a.reps, "C1 64 chunks")
    case(2_500_000, 10, 16, a.bits, dev, a.reps, "C2")
    case(16 * item * (2_500_000 // (16 * item)), 10, 16, a.bits, dev, a.reps, "C2 whole-item chunks")
    # fixed cost of a launch: tiny vectors
    for L in (item * 16, item * 16 * 148, 250_000, 500_000, 2_000_000, 4_000_000):
        case(L, 3, 16, a.bits, dev, a.reps, "size sweep")


if __name__ == "__main__":
    main()
