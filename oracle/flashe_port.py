"""Python port of the reference's CPU path — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Purpose: `bench.py`'s cpu_baseline leg and `bench.py --impl reference` need the reference's own CPU
implementation timed on the GPU box's host cores.  The reference is Python and cannot travel
(/root/reference does not exist on the GPU box), so this module restates its algorithm with the SAME
cost structure — a fresh `multiprocessing.Pool(N_JOBS)` per call, one OpenSSL AES-256-ECB call per
16-byte block, per-element Python big-int loops, numpy object arrays — so that its timing stands in
for the reference's.  (oracle/flashe_oracle.c is the fast C restatement used as the checker.)

Each function names the reference lines it follows (federatedml/secureprotol/ unless stated).
Pinned against tests/golden/flashe_golden.npz (outputs of the real reference) by
tests/test_oracle_golden.py::test_python_port_matches_golden.

pycryptodome (reference pin 3.9.9) is not installed anywhere in this image; `cryptography` (OpenSSL)
provides the AES primitive with a comparable per-call cost.
"""
from functools import reduce
from multiprocessing import Pool, cpu_count

import numpy as np
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

N_JOBS = cpu_count()          # jzf_flashe.py:7


def chunk_ranges(length, n):
    """jzf_flashe.py:12-16."""
    d, r = divmod(length, n)
    out, start = [], 0
    for i in range(n):
        size = d + 1 if i < r else d
        out.append((start, start + size))
        start += size
    return out


def _aes_key(seed):
    """jzf_aes.py:21-28: low 32 bytes of the seed."""
    return (int.from_bytes(seed, 'big') & (256 ** 32 - 1)).to_bytes(32, 'big')


def _worker_streams(begin, end, seed, int_bits, prefixes):
    """One chunk, any number of keystreams — the loop of _static_prepare_encrypt /
    _static_prepare_decrypt (jzf_flashe.py:48-82, 115-152): per block one AES call per stream on
    prefix || (i+begin) as 8 bytes BE, then int_bits-wide slices from the least significant end.
    Returns one list per prefix."""
    enc = Cipher(algorithms.AES(_aes_key(seed)), modes.ECB()).encryptor()
    n = end - begin
    per_block = 128 // int_bits
    blocks = (n - 1) // per_block + 1
    lo_mask = (1 << int_bits) - 1
    outs = [[] for _ in prefixes]
    for i in range(blocks):
        first = i * per_block
        last = min(first + per_block, n)
        ctr = (i + begin).to_bytes(8, 'big')
        for k, prefix in enumerate(prefixes):
            s = int.from_bytes(enc.update(prefix + ctr), 'big')
            dst = outs[k]
            for _ in range(first, last):
                dst.append(s & lo_mask)
                s >>= int_bits
    return outs


def keystreams(seed, int_bits, it, prf_indices, length, n_jobs=None):
    """F(it, c) for every c in prf_indices, as object arrays (Pool fan-out of jzf_flashe.py:459-476)."""
    n_jobs = N_JOBS if n_jobs is None else n_jobs
    prefixes = [int(it).to_bytes(4, 'big') + int(c).to_bytes(4, 'big') for c in prf_indices]
    jobs = [(b, e, seed, int_bits, prefixes) for b, e in chunk_ranges(length, n_jobs)]
    pool = Pool(n_jobs)
    parts = pool.starmap(_worker_streams, jobs)
    pool.close()
    pool.join()
    streams = []
    for k in range(len(prefixes)):
        flat = []
        for part in parts:
            flat += part[k]
        streams.append(np.array(flat, dtype=object))
    return streams


def encrypt(seed, int_bits, it, idx, scheme, q, n_jobs=None):
    """FlasheCipher.encrypt (jzf_flashe.py:431-504)."""
    mask = (1 << int_bits) - 1
    if scheme == "double":
        add, minus = keystreams(seed, int_bits, it, [idx, idx + 1], len(q), n_jobs)
        ret = q + add - minus
    else:
        add, = keystreams(seed, int_bits, it, [idx], len(q), n_jobs)
        ret = q + add
    ret &= mask
    return ret


def collapse_runs(survivors):
    """set_idx_list(mode='decrypt') (jzf_flashe.py:356-367)."""
    add, minus = [], []
    for idx in sorted(survivors):
        if add and idx == add[-1]:
            add[-1] = idx + 1
        else:
            add.append(idx + 1)
            minus.append(idx)
    return add, minus


def decrypt(seed, int_bits, it, survivors, scheme, agg, n_jobs=None):
    """FlasheCipher.decrypt (jzf_flashe.py:506-594)."""
    mask = (1 << int_bits) - 1
    if scheme == "double":
        add_idx, minus_idx = collapse_runs(survivors)
        streams = keystreams(seed, int_bits, it, add_idx + minus_idx, len(agg), n_jobs)
        add = reduce(lambda a, b: a + b, streams[:len(add_idx)])
        minus = reduce(lambda a, b: a + b, streams[len(add_idx):])
        ret = agg + (add & mask) - (minus & mask)
    else:
        streams = keystreams(seed, int_bits, it, list(survivors), len(agg), n_jobs)
        ret = agg - (reduce(lambda a, b: a + b, streams) & mask)
    ret &= mask
    return ret


def quantize(x, alpha, element_bits=16, u=None):
    """_static_quantize_padding_asymmetric (jzf_quantize.py:55-67); alpha must be a Python float.
    `u` replaces np.random.random(size) when given (parity checks)."""
    v = np.clip(x, -alpha, alpha) + alpha
    v = v * ((1 << element_bits) - 1) / (2 * alpha)
    noise = np.random.random(v.shape) if u is None else u
    return np.floor(v + noise).astype(int).astype(object)


def unquantize(v, alpha, element_bits, n_clients):
    """_static_unquantize_padding_asymmetric (jzf_quantize.py:102-107)."""
    alpha = alpha * n_clients
    return v * (2 * alpha) / (((1 << element_bits) - 1) * n_clients) - alpha


def aggregate_elementwise(cts, int_bits):
    """framework/homo/procedure/jzf_aggregator.py:421-430."""
    mod = 1 << int_bits
    return reduce(lambda a, b: (a + b) % mod, cts)


def pack(ct, int_bits):
    """Value of JZFTransferableWeights.compress (framework/jzf_weights.py:155-195): first element
    most significant."""
    s = 0
    for v in ct:
        s = (s << int_bits) + int(v)
    return s


def aggregate_packed(cts, int_bits):
    """jzf_aggregator.py:406-419 on the packed integers, then the split of
    framework/jzf_weights.py:98-137."""
    length = len(cts[0])
    mod = 1 << (int_bits * length)
    total = reduce(lambda a, b: (a + b) % mod, [pack(ct, int_bits) for ct in cts])
    lo_mask = (1 << int_bits) - 1
    out = []
    for _ in range(length):
        out.append(total & lo_mask)
        total >>= int_bits
    out.reverse()
    return np.array(out, dtype=object)


def run_round(seed, int_bits, it, xs, alpha, element_bits=16, n_jobs=None, us=None, timings=None):
    """One full round driven like the reference's notebook (encrypt_test/final_big_table.ipynb:222-233,
    408-411) and client/arbiter code (jzf_aggregator.py:722-739, 421-430, 883-899):
    encode + encrypt per client, element-wise server sum, decrypt, decode."""
    import time
    n = len(xs)
    t = {"encode": 0.0, "encrypt": 0.0, "aggregate": 0.0, "decrypt": 0.0, "decode": 0.0}
    cts = []
    for c, x in enumerate(xs):
        t0 = time.perf_counter()
        q = quantize(x, alpha, element_bits, None if us is None else us[c])
        t1 = time.perf_counter()
        cts.append(encrypt(seed, int_bits, it, c, "double", q, n_jobs))
        t2 = time.perf_counter()
        t["encode"] += t1 - t0
        t["encrypt"] += t2 - t1
    t0 = time.perf_counter()
    agg = aggregate_elementwise(cts, int_bits)
    t1 = time.perf_counter()
    dec = decrypt(seed, int_bits, it, list(range(n)), "double", agg, n_jobs)
    t2 = time.perf_counter()
    out = unquantize(dec, alpha, element_bits, n)
    t3 = time.perf_counter()
    t["aggregate"] += t1 - t0
    t["decrypt"] += t2 - t1
    t["decode"] += t3 - t2
    if timings is not None:
        timings.update(t)
    return cts, agg, dec, out
