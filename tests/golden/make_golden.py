#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference code.

Run (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs: the reference's own `federatedml/secureprotol/jzf_flashe.py`, `jzf_aes_prp.py`,
`jzf_aes.py` and `jzf_quantize.py` imported from /root/reference, with
  * `oracle/ref_shim` standing in for pycryptodome (AES-256-ECB through OpenSSL), and
  * `multiprocessing.Pool` replaced by an in-process pool (same starmap semantics, no fork), so that
    `N_JOBS` can be pinned (the reference takes it from `cpu_count()`, jzf_flashe.py:7) and the
    fixtures do not depend on the machine that generated them.
The server-side sums cannot be imported (jzf_aggregator.py:16,19 import modules that do not exist),
so the two aggregate semantics are evaluated here with Python big ints exactly as written at
jzf_aggregator.py:404-430 on the packing of jzf_weights.py:45-84 (first element most significant).

Output: tests/golden/flashe_golden.npz (+ a JSON manifest stored inside under key "manifest").
Wide integers (int_bits > 64) are stored as (lo, hi) uint64 pairs.
"""
import json
import os
import sys
from functools import reduce

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(0, "/root/reference")

import federatedml.secureprotol.jzf_flashe as ref_flashe  # noqa: E402
import federatedml.secureprotol.jzf_quantize as ref_quant  # noqa: E402
from federatedml.secureprotol.jzf_aciq import ACIQ  # noqa: E402


class _InlinePool(object):
    """Same interface the reference uses (Pool(n).starmap / close / join), run in-process."""

    def __init__(self, n):
        self.n = n

    def starmap(self, fn, inputs):
        return [fn(*a) for a in inputs]

    def close(self):
        pass

    def join(self):
        pass


ref_flashe.Pool = _InlinePool

KEY = bytes(range(32))
OUT = {}
MANIFEST = {"key_hex": KEY.hex(), "cases": []}


def put(name, arr):
    assert name not in OUT, name
    OUT[name] = arr


def put_ints(name, values, bits):
    """Store a sequence of Python ints (< 2^bits)."""
    vals = [int(v) for v in values]
    if bits <= 64:
        put(name, np.array(vals, dtype=np.uint64))
    else:
        m64 = (1 << 64) - 1
        put(name + "__lo", np.array([v & m64 for v in vals], dtype=np.uint64))
        put(name + "__hi", np.array([v >> 64 for v in vals], dtype=np.uint64))


def make_cipher(int_bits, idx, it, n_jobs, mask="double", seed=KEY):
    ref_flashe.N_JOBS = n_jobs
    c = ref_flashe.FlasheCipher(int_bits, mask=mask)
    c.generate_prp_seed(seed)
    c.idx = idx
    c.set_iter_index(it)
    return c


def obj(vals):
    return np.array([int(v) for v in vals], dtype=object)


# ------------------------------------------------------------------ A. raw mask streams F(t, c)
def case_masks():
    cfgs = [
        # (int_bits, n_jobs, L, iter, prf_idx)
        (20, 8, 1000, 0, 0), (20, 8, 1003, 7, 5), (20, 1, 257, 1, 1), (20, 16, 3, 0, 2),
        (20, 16, 19, 2, 9), (32, 8, 1000, 0, 0), (32, 8, 4099, 3, 64), (32, 1024, 5000, 1, 3),
        (24, 8, 1001, 0, 1), (22, 8, 777, 5, 63), (16, 8, 500, 0, 0), (8, 4, 300, 0, 1),
        (40, 8, 333, 2, 2), (64, 8, 500, 1, 1), (63, 3, 100, 0, 0), (120, 8, 301, 2, 1),
        (128, 8, 100, 0, 3), (100, 5, 77, 4, 4), (32, 8, 1, 0, 0), (20, 8, 8, 0, 0),
        (20, 8, 6 * 8 + 1, 0, 0), (32, 7, 2 ** 16 + 5, 1000000, 2 ** 31 - 1),
    ]
    for i, (b, nj, L, it, pidx) in enumerate(cfgs):
        c = make_cipher(b, pidx, it, nj, mask="single")
        ct = c.encrypt(obj([0] * L))
        name = "masks_%02d" % i
        put_ints(name, ct, b)
        MANIFEST["cases"].append(dict(kind="masks", name=name, int_bits=b, n_jobs=nj, L=L,
                                      iter=it, prf_idx=pidx))


# ------------------------------------------------------------------ helpers for the server sums
def pack(vals, b):
    """jzf_weights.py:36-42 (`_to_bytes_old`), same value as `_to_bytes`/compress: first element
    most significant."""
    s = 0
    for v in vals:
        s = (s << b) + int(v)
    return s


def unpack(s, L, b):
    m = (1 << b) - 1
    out = []
    for _ in range(L):
        out.append(s & m)
        s >>= b
    out.reverse()
    return out


def aggregate_packed(cts, b):
    """jzf_aggregator.py:406-419 — is_compressed branch."""
    L = len(cts[0])
    mod = 1 << (b * L)
    total = reduce(lambda x, y: (x + y) % mod, [pack(ct, b) for ct in cts])
    return unpack(total, L, b)


def aggregate_elementwise(cts, b):
    """jzf_aggregator.py:421-430 — decompressed branch, object arrays."""
    mod = 1 << b
    return reduce(lambda x, y: (x + y) % mod, [obj(ct) for ct in cts])


def quantize(x_f32, alpha, seed, ebits=16):
    np.random.seed(seed)
    u = np.random.random(x_f32.shape)
    np.random.seed(seed)
    q = ref_quant._static_quantize_padding_asymmetric(x_f32, float(alpha), ebits)
    return q, u


# ------------------------------------------------------------------ B. full round trips
def case_roundtrip():
    cfgs = [
        # (name, int_bits, n_jobs, L, n_clients, iter, scheme)
        ("rt_b20_n3", 20, 8, 3001, 3, 0, "double"),
        ("rt_b20_n10", 20, 8, 1200, 10, 4, "double"),
        ("rt_b32_n5", 32, 16, 2050, 5, 1, "double"),
        ("rt_b32_single", 32, 8, 1500, 4, 2, "single"),
        ("rt_b20_single", 20, 3, 700, 3, 0, "single"),
        ("rt_b24_n6", 24, 8, 999, 6, 9, "double"),
        ("rt_b64_n3", 64, 8, 400, 3, 1, "double"),
    ]
    for name, b, nj, L, n, it, scheme in cfgs:
        sigma = 0.1
        alpha = float(ACIQ(16).get_alpha_gaus_direct(sigma))
        xs, us, qs, cts = [], [], [], []
        for c in range(n):
            x = (np.random.RandomState(1000 + c).standard_normal(L) * sigma).astype(np.float32)
            if c == 0:  # exercise the clip and the exact end points
                x[:6] = np.array([alpha, -alpha, 10.0, -10.0, 0.0, np.float32(alpha)], dtype=np.float32)
            q, u = quantize(x, alpha, 2000 + c)
            cipher = make_cipher(b, c, it, nj, mask=scheme)
            ct = cipher.encrypt(q.copy())
            xs.append(x); us.append(u); qs.append(q); cts.append(ct)
        agg_b = aggregate_elementwise(cts, b)
        agg_a = obj(aggregate_packed(cts, b))
        dec = {}
        for tag, agg in (("B", agg_b), ("A", agg_a)):
            cipher = make_cipher(b, 0, it, nj, mask=scheme)
            cipher.set_num_clients(n)
            cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
            dec[tag] = cipher.decrypt(agg.copy())
        decoded = ref_quant._static_unquantize_padding_asymmetric(dec["B"].copy(), float(alpha), 16, n)
        put(name + "_x", np.stack(xs)); put(name + "_u", np.stack(us))
        put_ints(name + "_q", np.concatenate(qs), 32)
        put_ints(name + "_ct", np.concatenate(cts), b)
        put_ints(name + "_aggB", agg_b, b); put_ints(name + "_aggA", agg_a, b)
        put_ints(name + "_decB", dec["B"], b); put_ints(name + "_decA", dec["A"], b)
        put(name + "_decoded", np.array([float(v) for v in decoded], dtype=np.float64))
        MANIFEST["cases"].append(dict(kind="roundtrip", name=name, int_bits=b, n_jobs=nj, L=L,
                                      n_clients=n, iter=it, scheme=scheme, alpha=alpha,
                                      element_bits=16))


# ------------------------------------------------------------------ C. dropout (no precompute)
def case_dropout():
    b, nj, L, n, it = 20, 8, 600, 4, 3
    qs, cts = [], []
    for c in range(n):
        q = obj(np.random.RandomState(500 + c).randint(0, 65536, size=L))
        cts.append(make_cipher(b, c, it, nj).encrypt(q.copy())); qs.append(q)
    put_ints("drop_q", np.concatenate(qs), 32)
    put_ints("drop_ct", np.concatenate(cts), b)
    sets = [[0, 1, 2, 3], [0, 1, 3], [1, 2, 3], [0, 2], [3], [2, 0, 3]]
    for k, surv in enumerate(sets):
        agg = aggregate_elementwise([cts[c] for c in surv], b)
        cipher = make_cipher(b, 0, it, nj)
        cipher.set_num_clients(n)
        cipher.set_idx_list(raw_idx_list=list(surv), mode="decrypt")
        add = [int.from_bytes(p[4:], "big") for p in cipher.index_prefix_for_add]
        minus = [int.from_bytes(p[4:], "big") for p in cipher.index_prefix_for_minus]
        dec = cipher.decrypt(agg.copy())
        expect = reduce(lambda x, y: x + y, [qs[c] for c in surv])
        assert all(int(a) == int(e) for a, e in zip(dec, expect)), surv
        put_ints("drop_%d_agg" % k, agg, b); put_ints("drop_%d_dec" % k, dec, b)
        MANIFEST["cases"].append(dict(kind="dropout", name="drop_%d" % k, int_bits=b, n_jobs=nj, L=L,
                                      n_clients=n, iter=it, survivors=list(surv), add=add, minus=minus))


# ------------------------------------------------------------------ D. precompute ≡ on the fly
def case_precompute():
    b, nj, L, n, it = 20, 8, 1234, 3, 5
    c = ref_flashe.FlasheCipher(b)
    ref_flashe.N_JOBS = nj
    c.generate_prp_seed(KEY); c.idx = 1; c.set_num_clients(n); c.set_num_params(L)
    c.set_iter_index(it - 1)
    c.prepare_encrypt()                      # masks for round `it`
    add = c.next_iter_encrypt_prepared["add"].copy()
    minus = c.next_iter_encrypt_prepared["minus"].copy()
    c.set_iter_index(it)
    q = obj(np.random.RandomState(77).randint(0, 65536, size=L))
    ct_pre = c.encrypt(q.copy())
    ct_fly = make_cipher(b, 1, it, nj).encrypt(q.copy())
    assert all(int(a) == int(e) for a, e in zip(ct_pre, ct_fly))
    # decrypt-side precompute with every client alive (the only case where the reference is right)
    c.prepare_decrypt()
    dadd = c.next_iter_decrypt_prepared["add"].copy()
    dminus = c.next_iter_decrypt_prepared["minus"].copy()
    put_ints("pre_q", q, 32); put_ints("pre_ct", ct_pre, b)
    put_ints("pre_enc_add", add, b); put_ints("pre_enc_minus", minus, b)
    put_ints("pre_dec_add", dadd, b); put_ints("pre_dec_minus", dminus, b)
    MANIFEST["cases"].append(dict(kind="precompute", name="pre", int_bits=b, n_jobs=nj, L=L,
                                  n_clients=n, iter=it, idx=1))


# ------------------------------------------------------------------ E. sparse, scheme single
def case_sparse():
    b, nj, n, it, total, k = 21, 8, 3, 2, 5000, 500
    masks, qs, cts, zeros = [], [], [], []
    for c in range(n):
        loc = np.sort(np.random.RandomState(4000 + c).choice(total, size=k + 7 * c, replace=False))
        masks.append([int(v) for v in loc])
        q = obj(np.random.RandomState(4100 + c).randint(0, 65536, size=len(loc)))
        zq = int(np.random.RandomState(4200 + c).randint(32767, 32769))  # that client's quantised 0.0
        cipher = make_cipher(b, c, it, nj, mask="single")
        cts.append(cipher.encrypt(q.copy())); qs.append(q); zeros.append(zq)
    # expand_to_dense, jzf_aggregator.py:150-165
    dense = []
    for c in range(n):
        e = np.zeros(total, dtype=object)
        e[masks[c]] = cts[c]
        zl = list(set(np.arange(total).tolist()) - set(masks[c]))
        e[zl] = zeros[c]
        dense.append(e)
    agg = aggregate_elementwise(dense, b)
    cipher = make_cipher(b, 0, it, nj, mask="single")
    cipher.masks = masks
    cipher.total = total
    cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
    dec = cipher.decrypt(agg.copy())
    for c in range(n):
        put("sparse_mask_%d" % c, np.array(masks[c], dtype=np.int64))
        put_ints("sparse_q_%d" % c, qs[c], 32)
        put_ints("sparse_ct_%d" % c, cts[c], b)
    put("sparse_zero", np.array(zeros, dtype=np.uint64))
    put_ints("sparse_agg", agg, b); put_ints("sparse_dec", dec, b)
    MANIFEST["cases"].append(dict(kind="sparse", name="sparse", int_bits=b, n_jobs=nj, n_clients=n,
                                  iter=it, total=total))


# ------------------------------------------------------------------ F. 120-bit lane batching
def case_batch():
    b, nj, L, n, it, ebits = 120, 8, 1001, 5, 1, 16
    factor = int(np.ceil(np.log2(n)))
    alpha = float(ACIQ(16).get_alpha_gaus_direct(0.05))
    xs, us, qs, ws, cts = [], [], [], [], []
    for c in range(n):
        x = (np.random.RandomState(6000 + c).standard_normal(L) * 0.05).astype(np.float32)
        q, u = quantize(x, alpha, 6100 + c)
        w = ref_quant._static_batching_padding_asymmetric(q, b, ebits, factor)
        ct = make_cipher(b, c, it, nj).encrypt(w.copy())
        xs.append(x); us.append(u); qs.append(q); ws.append(w); cts.append(ct)
    agg = aggregate_elementwise(cts, b)
    cipher = make_cipher(b, 0, it, nj)
    cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
    dec = cipher.decrypt(agg.copy())
    # the dense path really sums the PACKED uploads (jzf_aggregator.py:404-419): carries leak between words
    agg_a = obj(aggregate_packed(cts, b))
    cipher = make_cipher(b, 0, it, nj)
    cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
    dec_a = cipher.decrypt(agg_a.copy())
    put_ints("batch_aggA", agg_a, b); put_ints("batch_decA", dec_a, b)
    unb = ref_quant._static_unbatching_padding_asymmetric(dec, b, ebits, factor)[:L]
    decoded = ref_quant._static_unquantize_padding_asymmetric(unb, float(alpha), ebits, n)
    put("batch_x", np.stack(xs)); put("batch_u", np.stack(us))
    put_ints("batch_q", np.concatenate(qs), 32)
    put_ints("batch_w", np.concatenate(ws), b); put_ints("batch_ct", np.concatenate(cts), b)
    put_ints("batch_agg", agg, b); put_ints("batch_dec", dec, b)
    put_ints("batch_unb", unb, 32)
    put("batch_decoded", np.array([float(v) for v in decoded], dtype=np.float64))
    MANIFEST["cases"].append(dict(kind="batch", name="batch", int_bits=b, n_jobs=nj, L=L, n_clients=n,
                                  iter=it, alpha=alpha, element_bits=ebits, factor=factor,
                                  words=len(ws[0])))


# ------------------------------------------------------------------ G. encode edge cases
def case_quant_edges():
    alpha = 0.59383450
    f = np.float32
    x = np.array([0.0, -0.0, alpha, -alpha, 1e-30, -1e-30, 0.25, -0.25, 3.0, -3.0,
                  np.nextafter(f(alpha), f(0)), np.nextafter(f(-alpha), f(0)),
                  1.17549435e-38, 0.1, 0.2, 0.3, 0.5938345], dtype=np.float32)
    x = np.concatenate([x, (np.random.RandomState(9).standard_normal(3000) * 0.3).astype(np.float32)])
    q, u = quantize(x, alpha, 4242)
    q1, u1 = quantize(x, 1.0, 4243)          # the 'zzz' layer alpha (jzf_quantize.py:434-435)
    q2, u2 = quantize(x, 0.1, 4244, ebits=8)  # alpha==0 fallback value (jzf_quantize.py:411-412)
    put("qe_x", x)
    put("qe_u", np.stack([u, u1, u2]))
    put_ints("qe_q", np.concatenate([q, q1, q2]), 32)
    MANIFEST["cases"].append(dict(kind="quant_edges", name="qe", alphas=[alpha, 1.0, 0.1],
                                  element_bits=[16, 16, 8], L=len(x)))
    # decode KAT on arbitrary sums
    vals = obj(np.random.RandomState(10).randint(0, 65536 * 7, size=2000))
    dec = ref_quant._static_unquantize_padding_asymmetric(vals.copy(), float(alpha), 16, 7)
    put_ints("qd_v", vals, 32)
    put("qd_out", np.array([float(v) for v in dec], dtype=np.float64))
    MANIFEST["cases"].append(dict(kind="decode", name="qd", alpha=alpha, element_bits=16, n_clients=7))


# ------------------------------------------------------------------ H. set_idx_list run collapse
def case_runs():
    rs = np.random.RandomState(31)
    rows = []
    for n in (1, 2, 4, 10, 64):
        for _ in range(6):
            keep = sorted(int(v) for v in np.nonzero(rs.rand(n) < 0.7)[0])
            if not keep:
                keep = [int(rs.randint(n))]
            c = make_cipher(20, 0, 0, 8)
            perm = list(keep)
            rs.shuffle(perm)
            c.set_idx_list(raw_idx_list=[int(v) for v in perm], mode="decrypt")
            rows.append(dict(n=n, survivors=[int(v) for v in perm],
                             add=[int.from_bytes(p[4:], "big") for p in c.index_prefix_for_add],
                             minus=[int.from_bytes(p[4:], "big") for p in c.index_prefix_for_minus]))
    MANIFEST["cases"].append(dict(kind="runs", name="runs", rows=rows))


# ------------------------------------------------------------------ I. packed sum of wide words
def case_packed_wide():
    """jzf_aggregator.py:406-419 on words wider than 64 bits (the shipped int_bits = 120 batch mode uploads
    those): random words plus runs of all-ones digits so that carries ripple across many words."""
    cfgs = [(65, 3, 200), (100, 16, 257), (120, 16, 300), (120, 2, 64), (127, 7, 130), (128, 5, 128), (128, 16, 300)]
    for i, (b, n, L) in enumerate(cfgs):
        rs = np.random.RandomState(7000 + i)
        full = (1 << b) - 1
        cts = [[int.from_bytes(rs.bytes(16), "little") & full for _ in range(L)] for _ in range(n)]
        lo, hi = L // 3, L // 3 + L // 4
        for j in range(lo, hi):                 # digit sums of exactly 2^b - 1: a carry entering at hi ripples to lo
            for c in range(n):
                cts[c][j] = full if c == 0 else 0
        cts[n - 1][hi] = full                   # and one enters: digit hi sums past 2^b
        cts[0][hi] = 1 if n > 1 else full
        name = "pkw_%d" % i
        agg = aggregate_packed(cts, b)
        put_ints(name + "_ct", [v for ct in cts for v in ct], b)
        put_ints(name + "_aggA", agg, b)
        MANIFEST["cases"].append(dict(kind="packed_wide", name=name, int_bits=b, n_clients=n, L=L))


def main():
    case_masks(); case_roundtrip(); case_dropout(); case_precompute(); case_sparse()
    case_batch(); case_quant_edges(); case_runs(); case_packed_wide()
    MANIFEST["generator"] = dict(numpy=np.__version__, python=sys.version.split()[0],
                                 reference="/root/reference (SamuelGong/FLASHE)")
    OUT["manifest"] = np.frombuffer(json.dumps(MANIFEST).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "flashe_golden.npz")
    np.savez_compressed(path, **OUT)
    print("wrote", path, os.path.getsize(path), "bytes,", len(MANIFEST["cases"]), "cases")


if __name__ == "__main__":
    main()
