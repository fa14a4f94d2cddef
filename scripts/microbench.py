#!/usr/bin/env python
"""Micro-benchmarks behind the PRF design decision, on one B200 (one JSON line each):

    lds        conflict-free 4-byte shared-memory lookups per clock and SM (k_lds_peak) — the measured ceiling of
               any table-driven AES; the 2-way-conflict variant must come out at half of it
    ttable     the product's mask-only kernel (flashe_masks: AES-256 from the replicated T-tables, nothing but the
               PRF and a 16-byte store per block) — the measured PRF peak `bench.py` quotes as roofline_prf.peak
    bitslice   a bit-sliced AES-256 (k_aes_bitslice, ~146-gate S-box, 32 blocks per thread), verified against
               FIPS-197 C.3 and the CPU oracle, then timed; rate scaled to the 113-gate Boyar-Peralta S-box too

`python scripts/microbench.py [--only lds|ttable|bitslice]`; run under ncu with --only to profile one kernel."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEY = bytes(range(32))


def micro_lib():
    from flashe_b200 import build
    path = build.MICRO_LIB
    if not os.path.exists(path):
        build.build_micro()
    lib = C.CDLL(path)
    lib.fm_last_error.restype = C.c_char_p
    lib.fm_lds_peak.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.fm_aes_bitslice.argtypes = [C.c_int, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_int,
                                    C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    return lib


def check(lib, rc):
    if rc != 0:
        raise RuntimeError(lib.fm_last_error().decode())


def sm_clock_mhz():
    import subprocess
    try:
        out = subprocess.check_output(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"], text=True)
        cur, mx = [float(v) for v in out.strip().split(",")]
        return cur, mx
    except Exception:
        return None, 1965.0


def bench_lds(lib, steps=20000):
    res = {}
    for conflict in (1, 2):
        ms, sms = C.c_double(), C.c_int()
        check(lib, lib.fm_lds_peak(0, steps, conflict, 5, C.byref(ms), C.byref(sms)))
        lookups = sms.value * 512 * steps * 16
        res[conflict] = (ms.value, lookups, sms.value)
    _, mx = sm_clock_mhz()
    ms, lookups, sms = res[1]
    rate = lookups / (ms * 1e-3)
    per_clk_sm = rate / (sms * mx * 1e6)
    line = {"bench": "lds_peak", "kernel": "k_lds_peak<1> (512 threads x %d SMs, 16 independent conflict-free LDS.32 per step)" % sms,
            "ms": ms, "lookups": lookups, "lookups_per_s": rate, "sm_max_mhz": mx,
            "lookups_per_clk_per_sm_at_max_clock": per_clk_sm,
            "two_way_conflict_ms": res[2][0], "two_way_conflict_ratio": res[2][0] / ms,
            "ttable_aes256_ceiling_g_blocks_per_s": {"197_lookups_per_block": rate / 197 / 1e9, "224_lookups_per_block": rate / 224 / 1e9}}
    print(json.dumps(line), flush=True)
    return line


def bench_ttable():
    import torch
    import flashe_b200 as fb
    ctx = fb.DeviceContext(KEY, 32, "cuda:0")
    L = 200_000_000
    span = fb.VectorSpan(L, 16)
    out = ctx.empty_words(L)
    res = {}
    for streams, sign in (([0], [1]), ([0, 1], [1, -1])):
        for _ in range(3):
            ctx.masks(0, streams, sign, span, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ctx.masks(0, streams, sign, span, out=out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        res[len(streams)] = (ms, len(streams) * L // 4)
    line = {"bench": "ttable_mask_only", "kernel": "k_stream<1, 4, M_MASKS, 0, 1> (flashe_masks, int_bits 32, 200M elements)",
            "one_stream_ms": res[1][0], "one_stream_g_blocks_per_s": res[1][1] / (res[1][0] * 1e-3) / 1e9,
            "two_streams_ms": res[2][0], "two_streams_g_blocks_per_s": res[2][1] / (res[2][0] * 1e-3) / 1e9,
            "bytes_written_per_block": {"one_stream": 16, "two_streams": 8}}
    print(json.dumps(line), flush=True)
    return line


def bench_bitslice(lib):
    from oracle import oracle as O
    key = (C.c_uint8 * 32)(*KEY)
    # FIPS-197 C.3: key 00..1f, plaintext 00112233445566778899aabbccddeeff = iter || prf || counter
    it, prf, ctr = 0x00112233, 0x44556677, 0x8899aabbccddeeff
    ctr0 = ctr & ~31
    nblocks = 64
    planes = np.zeros((nblocks // 32, 128), dtype=np.uint32)
    check(lib, lib.fm_aes_bitslice(0, key, it, prf, ctr0, nblocks, 1, planes.ctypes.data_as(C.c_void_p), 0, None))
    ok = True
    for t in range(nblocks // 32):
        for k in range(32):
            got = bytes(sum((int(planes[t, 8 * i + b] >> k) & 1) << b for b in range(8)) for i in range(16))
            blk = it.to_bytes(4, "big") + prf.to_bytes(4, "big") + (ctr0 + 32 * t + k).to_bytes(8, "big")
            want = O.aes256_encrypt_block(KEY, blk)
            ok = ok and got == want
            if ctr0 + 32 * t + k == ctr:
                ok = ok and got.hex() == "8ea2b7ca516745bfeafc49904b496089"
    nblocks = 1 << 28
    ms = C.c_double()
    check(lib, lib.fm_aes_bitslice(0, key, 0, 1, 0, nblocks, 0, None, 3, C.byref(ms)))
    rate = nblocks / (ms.value * 1e-3)
    line = {"bench": "aes256_bitsliced", "kernel": "k_aes_bitslice (32 blocks per thread, 128 bit planes in registers, generated 146-gate S-box)",
            "verified": "FIPS-197 C.3 + 64 blocks vs the CPU oracle: %s" % ("ok" if ok else "MISMATCH"),
            "ms": ms.value, "blocks": nblocks, "g_blocks_per_s": rate / 1e9,
            "g_blocks_per_s_scaled_to_113_gate_sbox": rate / 1e9 * (146.0 * 16 + 450) / (113.0 * 16 + 450),
            "note": "output stays bit-sliced (one word stored per 32 blocks): a usable kernel still has to transpose the planes"}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    lib = micro_lib()
    if args.only in ("", "lds"):
        bench_lds(lib)
    if args.only in ("", "ttable"):
        bench_ttable()
    if args.only in ("", "bitslice"):
        bench_bitslice(lib)


if __name__ == "__main__":
    main()
