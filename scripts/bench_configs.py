#!/usr/bin/env python
"""BASELINE configs C1-C4 on one B200, device-timed (CUDA events), inputs resident in HBM.  C5 is
bench.py.  Parity of these configurations is tests/test_configs_gpu.py; this script only times them.
One JSON line per configuration: `python scripts/bench_configs.py > profiles/<round>_configs.jsonl`."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashe_b200 as fb  # noqa: E402

KEY = bytes(range(32))
ALPHA = 5.938345 * 0.1


def timed(fn, steps=5, warmup=3):
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


def dense_round(name, L, n, bits, n_jobs, dev):
    """C1 / C2: encode+encrypt (n clients) -> aggregate (B) -> decrypt+decode."""
    ctx = fb.DeviceContext(KEY, bits, dev)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=7, stream=0)
    x = torch.randn(n, L, device=dev) * 0.1
    cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
    out = torch.empty(L, dtype=torch.float64, device=dev)

    def rnd():
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, noise, span, out=cts)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)

    ms_calls = timed(rnd)                       # three library calls per round from Python (host-side marshalling included)
    replay = ctx.capture(rnd)                   # the same three kernels as one CUDA graph launch
    ms = timed(replay, steps=20, warmup=5)
    noise32 = fb.NoiseSpec(seed=7, stream=0, resolution=32)

    def rnd32():
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, codec, noise32, span, out=cts)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(0, [n], [0], agg, codec, span, out=out)

    ms32 = timed(ctx.capture(rnd32), steps=20, warmup=5)
    m = 128 // bits
    print(json.dumps({"config": name, "elements": L, "clients": n, "int_bits": bits, "n_jobs": n_jobs, "ms_per_round": ms,
                      "ms_per_round_separate_calls": ms_calls, "ms_per_round_noise_32bit_resolution": ms32,
                      "g_aes_blocks_per_s_noise_32bit_resolution": (2 * n + 2) * -(-L // (128 // bits)) / (ms32 * 1e-3) / 1e9, "schedule": "CUDA graph of 3 kernels (encode+encrypt batch, aggregate, decrypt+decode)",
                      "client_elements_per_s": n * L / (ms * 1e-3), "aes_blocks_per_round": (2 * n + 2) * -(-L // m),
                      "g_aes_blocks_per_s": (2 * n + 2) * -(-L // m) / (ms * 1e-3) / 1e9}), flush=True)


def batched_round(dev, L=25_000_000, n=10, n_jobs=16, bits=120, element_bits=16):
    """The shipped batch configuration (int_bits 120, batch: true): encode -> 6 lanes per 120-bit word
    (jzf_quantize.py:162-185) -> encrypt (one AES block per word and stream) -> aggregate -> decrypt ->
    unbatch -> decode."""
    ctx = fb.DeviceContext(KEY, bits, dev)
    ctx32 = fb.DeviceContext(KEY, 32, dev)
    factor = (n - 1).bit_length()
    bs = bits // (element_bits + factor)
    nw = -(-L // bs)
    span_e, span_w = fb.VectorSpan(L, n_jobs), fb.VectorSpan(nw, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=element_bits, n_clients=n)
    x = torch.randn(n, L, device=dev) * 0.1
    cts = ctx.empty_words(nw, rows=n)

    def rnd():
        for c in range(n):
            q = ctx32.encode(x[c], codec, fb.NoiseSpec(seed=7, stream=c), span_e)
            w = ctx.batch_pack(q, element_bits, factor)
            ctx.encrypt(0, c, fb.SCHEME_DOUBLE, w, span_w, out=cts[c])
        agg = ctx.aggregate(cts, fb.AGG_ELEMENTWISE)
        p = ctx.decrypt(0, [n], [0], agg, span_w)
        ctx32.decode(ctx.batch_unpack(p, element_bits, factor)[:L].contiguous(), codec, span_e)

    def enc_only():
        for c in range(n):
            ctx.encrypt(0, c, fb.SCHEME_DOUBLE, cts[c], span_w, out=cts[c])

    # the same round through the fused entries: encode -> pack -> mask in ONE launch for all clients,
    # unmask -> unbatch -> decode in one launch (3 launches per round instead of 3 n + 4)
    bcodec = fb.CodecSpec(alpha=[ALPHA], element_bits=element_bits, n_clients=n, seg_end=[L], batch_lane_bits=element_bits + factor)
    cts2 = ctx.empty_words(nw, rows=n)
    agg2 = ctx.empty_words(nw)
    out2 = torch.empty(L, dtype=torch.float64, device=dev)

    def rnd_fused():
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, bcodec, fb.NoiseSpec(seed=7, stream=0), span_w, out=cts2)
        ctx.aggregate(cts2, fb.AGG_ELEMENTWISE, out=agg2)
        ctx.decrypt_decode(0, [n], [0], agg2, bcodec, span_w, out=out2)

    def rnd_fused_packed():
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, bcodec, fb.NoiseSpec(seed=7, stream=0), span_w, out=cts2)
        ctx.aggregate(cts2, fb.AGG_PACKED, out=agg2)
        ctx.decrypt_decode(0, [n], [0], agg2, bcodec, span_w, out=out2)

    ms = timed(rnd, steps=3, warmup=2)
    enc_ms = timed(enc_only, steps=3, warmup=1)
    rnd(); rnd_fused()
    same = bool(torch.equal(cts.view(torch.int64), cts2.view(torch.int64)))
    fused_ms = timed(rnd_fused, steps=5, warmup=2)

    def rnd_fused32():
        ctx.encode_encrypt_batch(0, 0, fb.SCHEME_DOUBLE, x, bcodec, fb.NoiseSpec(seed=7, stream=0, resolution=32), span_w, out=cts2)
        ctx.aggregate(cts2, fb.AGG_ELEMENTWISE, out=agg2)
        ctx.decrypt_decode(0, [n], [0], agg2, bcodec, span_w, out=out2)

    fused32_ms = timed(rnd_fused32, steps=5, warmup=2)
    fused_packed_ms = timed(rnd_fused_packed, steps=5, warmup=2)
    print(json.dumps({"config": "batched: 25M elements as 120-bit words (6 lanes), 10 clients, full round", "elements": L, "words": nw, "clients": n,
                      "int_bits": bits, "n_jobs": n_jobs, "ms_per_round": fused_ms, "client_elements_per_s": n * L / (fused_ms * 1e-3),
                      "ms_per_round_packed_carry_sum": fused_packed_ms, "ms_per_round_noise_32bit_resolution": fused32_ms,
                      "ms_per_round_unfused_7_kernels": ms, "fused_ciphertexts_equal_unfused": same,
                      "g_aes_blocks_per_s": (2 * n + 2) * nw / (fused_ms * 1e-3) / 1e9,
                      "encrypt_only_ms_10_clients": enc_ms, "encrypt_g_aes_blocks_per_s": 2 * n * nw / (enc_ms * 1e-3) / 1e9}), flush=True)


def precompute_c3(dev, L=25_000_000, rounds=16, n=10, bits=20, n_jobs=16):
    """C3: one client fills the combined masks of 16 future rounds, then the online round of each
    (encode + add of the stored mask); the server decrypts under 20 % dropout (8 survivors, 3 runs)."""
    ctx = fb.DeviceContext(KEY, bits, dev)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    x = torch.randn(L, device=dev) * 0.1
    ring = ctx.empty_words(L, rows=rounds)
    ct = ctx.empty_words(L)

    def fill():
        for t in range(rounds):
            ctx.masks(t, [3, 4], [1, -1], span, out=ring[t])

    def online():
        for t in range(rounds):
            ctx.encode_add_premasked(x, codec, fb.NoiseSpec(seed=7, stream=t), ring[t], span, out=ct)

    fill_ms = timed(fill, steps=2, warmup=1)
    online_ms = timed(online, steps=2, warmup=1)
    agg = ct
    out = torch.empty(L, dtype=torch.float64, device=dev)
    # survivors {0,1,3,4,5,6,8,9}: runs [0,1] [3..6] [8,9] -> add {2, 7, 10}, minus {0, 3, 8}
    dec_ms = timed(lambda: ctx.decrypt_decode(0, [2, 7, 10], [0, 3, 8], agg, codec, span, out=out))
    blocks = 2 * rounds * -(-L // (128 // bits))
    print(json.dumps({"config": "C3 mask precomputation, 16 rounds x 25M, double masking, 20% dropout", "elements": L, "rounds": rounds,
                      "int_bits": bits, "n_jobs": n_jobs, "fill_ms": fill_ms, "fill_g_aes_blocks_per_s": blocks / (fill_ms * 1e-3) / 1e9,
                      "fill_round_elements_per_s": rounds * L / (fill_ms * 1e-3), "ring_bytes": int(ring.numel() * 4),
                      "online_ms_16_rounds": online_ms, "online_elements_per_s": rounds * L / (online_ms * 1e-3),
                      "online_gbs": rounds * L * 12 / (online_ms * 1e-3) / 1e9,
                      "decrypt_decode_3_runs_ms": dec_ms, "decrypt_decode_elements_per_s": L / (dec_ms * 1e-3)}), flush=True)


def sparse_c4(dev, total=50_000_000, n=32, bits=32, n_jobs=16, frac=0.01):
    """C4: per client top-1 % of a 50M-element gradient (+ residual), single-mask encrypt of the compact
    values, expand to dense + element-wise sum on the server, per-index unmasking, overlap counts."""
    ctx = fb.DeviceContext(KEY, bits, dev)
    k = int(total * frac)
    span = fb.VectorSpan(k, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    g = torch.Generator(device=dev); g.manual_seed(3)
    x = torch.randn(total, device=dev, generator=g) * 0.1
    res = torch.zeros(total, device=dev)
    res_out = torch.empty_like(res)
    vals, idx, _ = ctx.topk_sparsify(x, [total], [k], residual=res, residual_out=res_out)
    topk_ms = timed(lambda: ctx.topk_sparsify(x, [total], [k], residual=res, residual_out=res_out))
    ct = ctx.empty_words(k)
    enc_ms = timed(lambda: ctx.encode_encrypt(0, 5, fb.SCHEME_SINGLE, vals, codec, fb.NoiseSpec(seed=1, stream=5), span, out=ct))
    # server: n clients with distinct index sets (shifted copies of the measured one keep the sizes exact)
    idxs = [torch.sort((idx + 977 * c) % total).values.contiguous() for c in range(n)]
    acc = ctx.zeros_words(total)

    def server_sum():
        a = acc
        for c in range(n):
            dense = ctx.sparse_expand(ct, idxs[c], total, 32768)
            a = ctx.aggregate(torch.stack([a.view(torch.int32), dense.view(torch.int32)]).view(torch.uint32))
        return a

    sum_ms = timed(server_sum, steps=2, warmup=1)
    fused = ctx.empty_words(total)
    fused_ms = timed(lambda: ctx.sparse_sum([ct] * n, idxs, total, [32768] * n, out=fused), steps=3, warmup=1)
    assert torch.equal(fused.view(torch.int32), server_sum().view(torch.int32))
    p = acc.clone()

    def unmask():
        for c in range(n):
            ctx.sparse_apply_masks(0, [c], [-1], span, idxs[c], p, validate=False)   # (index lists validated once, outside the timed loop)

    unmask_ms = timed(unmask, steps=2, warmup=1)
    p2 = acc.clone()
    unmask_batch_ms = timed(lambda: ctx.sparse_apply_masks_batch(0, list(range(n)), -1, n_jobs, idxs, p2, validate=False), steps=3, warmup=1)
    ov_ms = timed(lambda: ctx.sparse_overlap(idxs, total), steps=2, warmup=1)
    print(json.dumps({"config": "C4 index-sparse top-1% of 50M, 32 clients, single masking", "total": total, "k": k, "clients": n, "int_bits": bits,
                      "client_topk_sparsify_ms": topk_ms, "client_encode_encrypt_compact_ms": enc_ms,
                      "server_expand_and_sum_32_clients_ms": sum_ms, "server_fused_sparse_sum_32_clients_ms": fused_ms, "server_unmask_32_clients_ms": unmask_ms, "unmask_32_clients_one_call_ms": unmask_batch_ms, "overlap_counts_ms": ov_ms,
                      "client_elements_per_s_dense_equivalent": n * total / ((n * (topk_ms + enc_ms) + fused_ms + unmask_batch_ms) * 1e-3)}), flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n_jobs = os.cpu_count() or 16
    if "--sparse-only" in sys.argv:
        sparse_c4(dev, n_jobs=n_jobs)
        return
    dense_round("C1 1M elements, 3 clients, int_bits 20", 1_000_000, 3, 20, n_jobs, dev)
    dense_round("C2 2.5M elements, 10 clients, int_bits 20", 2_500_000, 10, 20, n_jobs, dev)
    dense_round("C2 at int_bits 32", 2_500_000, 10, 32, n_jobs, dev)
    dense_round("25M elements, 10 clients, int_bits 20 (C3-sized dense round)", 25_000_000, 10, 20, n_jobs, dev)
    dense_round("25M elements, 10 clients, int_bits 24", 25_000_000, 10, 24, n_jobs, dev)
    dense_round("25M elements, 10 clients, int_bits 64", 25_000_000, 10, 64, n_jobs, dev)
    if "--dense-only" in sys.argv:
        return
    batched_round(dev, n_jobs=n_jobs)
    precompute_c3(dev, n_jobs=n_jobs)
    sparse_c4(dev, n_jobs=n_jobs)


if __name__ == "__main__":
    main()
