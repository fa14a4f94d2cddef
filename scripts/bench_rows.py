#!/usr/bin/env python
"""Throughput of the kernels either side of the hot path (SURVEY §8 f1-f3 and the HBM-bound helpers) on
one B200: algorithmic bytes / CUDA-event time against the measured HBM copy peak.  Prints one JSON
line per kernel; `python scripts/bench_rows.py > profiles/<round>_rows.jsonl`."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

KEY = bytes(range(32))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


ONCE = "--once" in sys.argv          # one warm-up + one timed launch per kernel: for `ncu --set full` captures


def timed(fn, steps=5, warmup=3):
    if ONCE:
        steps, warmup = 1, 1
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def report(name, ms, nbytes, units, unit_name, **extra):
    peak = hbm_peak()
    gbs = nbytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "ms": ms, "algorithmic_bytes": int(nbytes), "gbs": gbs, "hbm_peak_gbs": peak, "frac_of_hbm": gbs / peak,
            "rate": units / (ms * 1e-3), "rate_unit": unit_name}
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev); g.manual_seed(1)
    L = 100_000_000
    # ---- f1 wire packing (working sets larger than the 126 MB L2)
    for bits in (20, 32):
        ctx = fb.DeviceContext(KEY, bits, dev)
        w = (torch.randint(0, 2 ** 31 - 1, (L,), device=dev, generator=g, dtype=torch.int64) & ((1 << bits) - 1)).to(torch.int32).view(torch.uint32)
        out = ctx.wire_pack(w)
        back = ctx.wire_unpack(out, L)
        assert torch.equal(back.view(torch.int32), w.view(torch.int32))
        ms = timed(lambda: ctx.wire_pack(w, out=out))
        report("wire_pack b=%d" % bits, ms, L * 4 + out.numel(), L, "elements/s", elements=L)
        ms = timed(lambda: ctx.wire_unpack(out, L, out=back))
        report("wire_unpack b=%d" % bits, ms, L * 4 + out.numel(), L, "elements/s", elements=L)
        del w, out, back
    ctx = fb.DeviceContext(KEY, 32, dev)
    # ---- f2 top-1% of a 50 M-element layer (BASELINE config 4), with residual
    n = 50_000_000
    x = torch.empty(n, dtype=torch.float32, device=dev).normal_(0.0, 0.1, generator=g)
    res = torch.zeros(n, dtype=torch.float32, device=dev)
    k = n // 100
    ms = timed(lambda: ctx.topk_sparsify(x, [n], [k], residual=res, residual_out=res))
    # algorithmic: read x once + residual in/out + compact out (the implementation re-reads x 5 times)
    report("topk_sparsify 1% of 50M (+residual)", ms, n * 12 + k * 12, n, "elements/s", elements=n, k=k,
           implementation_bytes=n * (5 * 4 + 8) + k * 12)
    # many layers: 200 layers of 250 k
    sizes = [250_000] * 200
    ends = np.cumsum(sizes)
    ks = [2500] * 200
    ms = timed(lambda: ctx.topk_sparsify(x, ends, ks, residual=res, residual_out=res))
    report("topk_sparsify 1% of 200 layers x 250k", ms, n * 12 + sum(ks) * 12, n, "elements/s", elements=n, layers=200)
    del x, res
    # ---- f3 layer statistics fused with the mean shift
    w = torch.empty(L, dtype=torch.float64, device=dev).normal_(0.0, 0.3, generator=g)
    wo = torch.empty_like(w)
    ends = np.cumsum([L // 50] * 50)
    ms = timed(lambda: ctx.segment_stats(w, ends, [0.1] * 50, out=wo))
    report("segment_stats 50 layers (shift + mean + std)", ms, L * 8 * 3, L, "elements/s", elements=L)
    del w, wo
    # ---- server sums
    nC, Ls = 16, 50_000_000
    cts = torch.randint(-2 ** 31, 2 ** 31 - 1, (nC, Ls), device=dev, generator=g, dtype=torch.int32).view(torch.uint32)
    agg = ctx.empty_words(Ls)
    ms = timed(lambda: ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg))
    report("aggregate elementwise n=16", ms, (nC + 1) * Ls * 4, nC * Ls, "client-elements/s")
    ms = timed(lambda: ctx.aggregate(cts, fb.AGG_PACKED, out=agg))
    report("aggregate packed-carry n=16", ms, (nC + 1) * Ls * 4, nC * Ls, "client-elements/s")
    del cts, agg
    # 16-byte words (the shipped int_bits = 120 batch mode): 16 clients x 12.5 M words = the same bytes
    ctx120 = fb.DeviceContext(KEY, 120, dev)
    Lw = Ls // 4
    ctw = torch.randint(-2 ** 63, 2 ** 63 - 1, (nC, Lw, 2), device=dev, generator=g, dtype=torch.int64)
    ctw[:, :, 1] &= (1 << 56) - 1
    ctw = ctw.view(torch.uint64)
    aggw = ctx120.empty_words(Lw)
    ms = timed(lambda: ctx120.aggregate(ctw, fb.AGG_ELEMENTWISE, out=aggw))
    report("aggregate elementwise n=16, int_bits 120", ms, (nC + 1) * Lw * 16, nC * Lw, "client-words/s")
    ms = timed(lambda: ctx120.aggregate(ctw, fb.AGG_PACKED, out=aggw))
    report("aggregate packed-carry n=16, int_bits 120", ms, (nC + 1) * Lw * 16, nC * Lw, "client-words/s")
    del ctw, aggw
    # ---- online step after precompute, masks of one client
    x = torch.empty(L, dtype=torch.float32, device=dev).normal_(0.0, 0.1, generator=g)
    span = fb.VectorSpan(L, 16)
    mask = ctx.masks(0, [0, 1], [1, -1], span)
    ct = ctx.empty_words(L)
    codec = fb.CodecSpec(alpha=0.5938345, element_bits=16)
    ms = timed(lambda: ctx.encode_add_premasked(x, codec, fb.NoiseSpec(seed=1, stream=0), mask, span, out=ct))
    report("encode_add_premasked (device noise)", ms, L * 12, L, "elements/s")
    ms = timed(lambda: ctx.encode_add_premasked(x, codec, fb.NoiseSpec(seed=1, stream=0, resolution=32), mask, span, out=ct))
    report("encode_add_premasked (device noise, 32-bit resolution)", ms, L * 12, L, "elements/s")
    ms = timed(lambda: ctx.add_premasked(ct, mask, 1, out=ct))
    report("add_premasked", ms, L * 12, L, "elements/s")
    ms = timed(lambda: ctx.masks(0, [0, 1], [1, -1], span, out=mask))
    report("masks fill (2 streams, b=32)", ms, L * 4, 2 * L // 4, "AES blocks/s")
    # write-only and read-only streams: what this part's HBM gives a kernel whose traffic is all stores / all loads
    # (the copy peak mixes them 1:1); decode writes 2 bytes for every byte it reads
    big = torch.empty(L * 2, dtype=torch.float32, device=dev)
    ms = timed(lambda: big.fill_(1.0))
    report("torch fill_ (write-only, 800 MB)", ms, L * 8, L * 2, "elements/s")
    ms = timed(lambda: big.sum())
    report("torch sum (read-only, 800 MB)", ms, L * 8, L * 2, "elements/s")
    del big
    out = torch.empty(L, dtype=torch.float64, device=dev)
    ms = timed(lambda: ctx.decode(ct, fb.CodecSpec(alpha=0.5938345, element_bits=16, n_clients=64), span, out=out))
    report("decode", ms, L * 12, L, "elements/s")


if __name__ == "__main__":
    main()
