"""ctypes binding of oracle/flashe_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  See the header of flashe_oracle.c for the parity status ("pinned by FIPS-197 C.3 and
by tests/golden/flashe_golden.npz, which holds outputs of the reference's own Python code").

Word layout (same as the device library): int_bits <= 32 -> uint32, <= 64 -> uint64,
<= 128 -> uint64 pairs (lo, hi), shape [L, 2].
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libflashe_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement with gcc (no GPU, no reference needed)."""
    src = os.path.join(_HERE, "flashe_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, vp, u64, u32, i32 = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.c_void_p,
                                        C.c_uint64, C.c_uint32, C.c_int)
        L.fo_aes256_expand.argtypes = [u8p, u8p]
        L.fo_aes256_encrypt_block.argtypes = [u8p, u8p, u8p]
        L.fo_reduce_key.argtypes = [u8p, C.c_size_t, u8p]
        L.fo_chunk_bounds.argtypes = [u64, u32, u32, C.POINTER(u64), C.POINTER(u64)]
        L.fo_masks.argtypes = [u8p, i32, u32, u32, i32p, i32p, i32, u64, u64, u64, vp]
        L.fo_apply_masks.argtypes = [u8p, i32, u32, u32, i32p, i32p, i32, u64, u64, u64, vp, vp]
        L.fo_collapse_runs.argtypes = [i32p, i32, i32p, i32p]
        L.fo_aggregate_elementwise.argtypes = [i32, i32, u64, u64, vp, vp]
        L.fo_aggregate_packed.argtypes = [i32, i32, u64, u64, vp, vp, u32, C.POINTER(u32)]
        L.fo_quantize.argtypes = [vp, vp, u64, C.c_double, i32, vp]
        L.fo_unquantize.argtypes = [vp, u64, i32, C.c_double, i32, i32, vp]
        L.fo_batch.argtypes = [vp, u64, i32, i32, i32, vp]
        L.fo_unbatch.argtypes = [vp, u64, i32, i32, i32, vp]
        L.fo_expand_to_dense.argtypes = [i32, vp, vp, u64, u64, vp, vp]
        L.fo_sparse_stream_accumulate.argtypes = [u8p, i32, u32, u32, C.c_int32, C.c_int32, vp, u64, u64, vp]
        L.fo_dynamic_masking.argtypes = [C.POINTER(vp), C.POINTER(u64), i32, C.POINTER(u64), C.POINTER(u64)]
        L.fo_set_threads.argtypes = [i32]
        for f in ("fo_masks", "fo_apply_masks", "fo_collapse_runs", "fo_aggregate_elementwise",
                  "fo_aggregate_packed", "fo_quantize", "fo_unquantize", "fo_batch", "fo_unbatch",
                  "fo_expand_to_dense", "fo_sparse_stream_accumulate", "fo_dynamic_masking",
                  "fo_num_threads", "fo_word_bytes"):
            getattr(L, f).restype = C.c_int
        _lib = L
    return _lib


def set_threads(n):
    lib().fo_set_threads(int(n))


# ----------------------------------------------------------------------------- word helpers
def word_bytes(int_bits):
    return 4 if int_bits <= 32 else (8 if int_bits <= 64 else 16)


def empty_words(n, int_bits):
    wb = word_bytes(int_bits)
    if wb == 4:
        return np.empty(n, dtype=np.uint32)
    if wb == 8:
        return np.empty(n, dtype=np.uint64)
    return np.empty((n, 2), dtype=np.uint64)


def to_words(values, int_bits):
    """Sequence of Python ints (< 2^int_bits) -> word array."""
    wb = word_bytes(int_bits)
    if wb == 4:
        return np.array([int(v) for v in values], dtype=np.uint32)
    if wb == 8:
        return np.array([int(v) for v in values], dtype=np.uint64)
    m = (1 << 64) - 1
    out = np.empty((len(values), 2), dtype=np.uint64)
    out[:, 0] = [int(v) & m for v in values]
    out[:, 1] = [int(v) >> 64 for v in values]
    return out


def from_words(arr, int_bits):
    """Word array -> list of Python ints."""
    if word_bytes(int_bits) == 16:
        return [int(lo) | (int(hi) << 64) for lo, hi in arr]
    return [int(v) for v in arr]


def _key(key):
    key = bytes(key)
    out = (C.c_uint8 * 32)()
    lib().fo_reduce_key((C.c_uint8 * len(key)).from_buffer_copy(key), len(key), out)
    return out


def _i32(xs):
    return (C.c_int32 * max(1, len(xs)))(*[int(x) for x in xs])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise ValueError("oracle %s failed (rc=%d)" % (what, rc))


# ----------------------------------------------------------------------------- primitives
def aes256_encrypt_block(key, block):
    rk = (C.c_uint8 * 240)()
    out = (C.c_uint8 * 16)()
    lib().fo_aes256_expand(_key(key), rk)
    lib().fo_aes256_encrypt_block(rk, (C.c_uint8 * 16).from_buffer_copy(bytes(block)), out)
    return bytes(out)


def chunk_bounds(L, n_jobs, i):
    b, e = C.c_uint64(), C.c_uint64()
    lib().fo_chunk_bounds(L, n_jobs, i, C.byref(b), C.byref(e))
    return b.value, e.value


def masks(key, int_bits, n_jobs, it, prf_idx, sign, L, j0=0, cnt=None):
    """sum_k sign[k]*F(it, prf_idx[k])[j] mod 2^b for j in [j0, j0+cnt)."""
    cnt = L - j0 if cnt is None else cnt
    out = empty_words(cnt, int_bits)
    _check(lib().fo_masks(_key(key), int_bits, n_jobs, it & 0xFFFFFFFF, _i32(prf_idx), _i32(sign),
                          len(prf_idx), L, j0, cnt, _ptr(out)), "masks")
    return out


def apply_masks(key, int_bits, n_jobs, it, prf_idx, sign, words, L=None, j0=0):
    words = np.ascontiguousarray(words)
    cnt = words.shape[0]
    L = cnt if L is None else L
    out = np.empty_like(words)
    _check(lib().fo_apply_masks(_key(key), int_bits, n_jobs, it & 0xFFFFFFFF, _i32(prf_idx),
                                _i32(sign), len(prf_idx), L, j0, cnt, _ptr(words), _ptr(out)), "apply")
    return out


def encrypt(key, int_bits, n_jobs, it, idx, scheme, q_words, L=None, j0=0):
    """FlasheCipher.encrypt (jzf_flashe.py:490-504)."""
    if scheme == "double":
        return apply_masks(key, int_bits, n_jobs, it, [idx, idx + 1], [1, -1], q_words, L, j0)
    return apply_masks(key, int_bits, n_jobs, it, [idx], [1], q_words, L, j0)


def collapse_runs(survivors):
    n = len(survivors)
    add, minus = (C.c_int32 * max(1, n))(), (C.c_int32 * max(1, n))()
    r = lib().fo_collapse_runs(_i32(survivors), n, add, minus)
    return list(add[:r]), list(minus[:r])


def decrypt(key, int_bits, n_jobs, it, survivors, scheme, agg_words, L=None, j0=0):
    """set_idx_list(mode='decrypt') + FlasheCipher.decrypt (jzf_flashe.py:306-314, 354-386, 506-594)."""
    if scheme == "double":
        add, minus = collapse_runs(survivors)
        idx = add + minus
        sign = [1] * len(add) + [-1] * len(minus)
    else:
        idx = list(survivors)
        sign = [-1] * len(idx)
    return apply_masks(key, int_bits, n_jobs, it, idx, sign, agg_words, L, j0)


def aggregate(int_bits, cts, mode="elementwise", carry_in=0, return_carry=False):
    """cts: array [n, L] (+[,2] for wide words).  mode 'elementwise' (B) or 'packed' (A)."""
    cts = np.ascontiguousarray(cts)
    n, L = cts.shape[0], cts.shape[1]
    out = np.empty_like(cts[0])
    if mode == "elementwise":
        _check(lib().fo_aggregate_elementwise(int_bits, n, L, L, _ptr(cts), _ptr(out)), "aggregate")
        return out
    co = C.c_uint32()
    _check(lib().fo_aggregate_packed(int_bits, n, L, L, _ptr(cts), _ptr(out), carry_in, C.byref(co)),
           "aggregate_packed")
    return (out, co.value) if return_carry else out


def quantize(x, u, alpha, element_bits=16):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    q = np.empty(x.shape[0], dtype=np.uint32)
    _check(lib().fo_quantize(_ptr(x), _ptr(u), x.shape[0], float(alpha), element_bits, _ptr(q)), "quantize")
    return q


def unquantize(v, alpha, element_bits, n_clients):
    v = np.ascontiguousarray(v)
    assert v.dtype in (np.uint32, np.uint64)
    out = np.empty(v.shape[0], dtype=np.float64)
    _check(lib().fo_unquantize(_ptr(v), v.shape[0], v.dtype.itemsize, float(alpha), element_bits,
                               n_clients, _ptr(out)), "unquantize")
    return out


def batch(q, int_bits, element_bits, factor):
    q = np.ascontiguousarray(q, dtype=np.uint32)
    bs = int_bits // (element_bits + factor)
    nw = (q.shape[0] + bs - 1) // bs
    out = empty_words(nw, int_bits)
    _check(lib().fo_batch(_ptr(q), q.shape[0], int_bits, element_bits, factor, _ptr(out)), "batch")
    return out


def unbatch(words, int_bits, element_bits, factor):
    words = np.ascontiguousarray(words)
    bs = int_bits // (element_bits + factor)
    out = np.empty(words.shape[0] * bs, dtype=np.uint32)
    _check(lib().fo_unbatch(_ptr(words), words.shape[0], int_bits, element_bits, factor, _ptr(out)), "unbatch")
    return out


def expand_to_dense(int_bits, compact, index, total, zero):
    compact = np.ascontiguousarray(compact)
    index = np.ascontiguousarray(index, dtype=np.int64)
    dense = empty_words(total, int_bits)
    z = to_words([zero], int_bits)
    _check(lib().fo_expand_to_dense(int_bits, _ptr(compact), _ptr(index), index.shape[0], total,
                                    _ptr(z), _ptr(dense)), "expand")
    return dense


def sparse_single_decrypt(key, int_bits, n_jobs, it, index_lists, total, agg_words):
    """jzf_flashe.py:315-343 + 528-535: subtract every client's scattered compact stream."""
    acc = np.ascontiguousarray(agg_words).copy()
    for c, index in enumerate(index_lists):
        index = np.ascontiguousarray(index, dtype=np.int64)
        _check(lib().fo_sparse_stream_accumulate(_key(key), int_bits, n_jobs, it & 0xFFFFFFFF, c, -1,
                                                 _ptr(index), index.shape[0], total, _ptr(acc)), "sparse")
    return acc


def dynamic_masking(index_lists):
    n = len(index_lists)
    arrs = [np.ascontiguousarray(ix, dtype=np.int64) for ix in index_lists]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ks = (C.c_uint64 * n)(*[a.shape[0] for a in arrs])
    sc, dc = C.c_uint64(), C.c_uint64()
    r = lib().fo_dynamic_masking(ptrs, ks, n, C.byref(sc), C.byref(dc))
    return ("single" if r == 0 else "double"), sc.value, dc.value


# ---------------------------------------------------------------------------------------------------
# Device noise generator, restated (throughput mode; include/flashe_b200.h: flashe_noise).  The
# reference draws np.random.random (MT19937, never seeded); the device library instead documents a
# counter-based stream: Philox4x32-10 (Salmon et al., SC'11; Random123 known answers in
# tests/test_oracle_golden.py), key = the 64-bit seed, counter = (j >> 1, stream), and the two
# word pairs of the output mapped to [0, 1) by numpy's res53 construction.
# ---------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint64 arrays holding 32-bit values; returns four uint64 arrays."""
    m32 = np.uint64(0xFFFFFFFF)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & m32, p0 & m32, n0, n2
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def noise_uniform(seed, stream, begin, count, resolution=53):
    """u_j for j in [begin, begin+count): what flashe_rng_uniform / the fused encode kernels draw.
    resolution=32 (throughput mode): u_j = word (j & 3) of Philox(counter = (j >> 2, stream)) * 2^-32."""
    j = np.arange(begin, begin + count, dtype=np.uint64)
    if resolution == 32:
        c = j >> np.uint64(2)
        z = np.zeros_like(c)
        o = philox4x32_10(c & np.uint64(0xFFFFFFFF), c >> np.uint64(32), z + np.uint64(int(stream) & 0xFFFFFFFF),
                          z + np.uint64((int(stream) >> 32) & 0xFFFFFFFF), int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
        k = (j & np.uint64(3)).astype(np.int64)
        w = np.choose(k, [o[0], o[1], o[2], o[3]])
        return w.astype(np.float64) / 4294967296.0
    c = j >> np.uint64(1)
    z = np.zeros_like(c)
    o = philox4x32_10(c & np.uint64(0xFFFFFFFF), c >> np.uint64(32), z + np.uint64(int(stream) & 0xFFFFFFFF),
                      z + np.uint64((int(stream) >> 32) & 0xFFFFFFFF), int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    odd = (j & np.uint64(1)).astype(bool)
    a = np.where(odd, o[2], o[0])
    b = np.where(odd, o[3], o[1])
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


# ---------------------------------------------------------------------------------------------------
# SURVEY §8 "next" rows, restated with numpy (test infrastructure like everything above).
# ---------------------------------------------------------------------------------------------------
def wire_pack(words, bits):
    """_to_bytes (framework/jzf_weights.py:45-84): s = sum_j a[j] << ((L-1-j)*bits), returned as the
    big-endian byte string of ceil(L*bits/8) bytes.  words: uint32/uint64 [L] or uint64 [L, 2] (lo, hi).
    The reference's per-batch loop only bounds its intermediate integers; the bits of s are the
    concatenation of the fields, first element first, right-aligned in the byte string."""
    w = np.ascontiguousarray(words)
    L = w.shape[0]
    nbytes = (L * bits + 7) // 8
    if L == 0:
        return np.zeros(0, dtype=np.uint8)
    if w.ndim == 2:
        lo, hi = w[:, 0].astype(np.uint64), w[:, 1].astype(np.uint64)
    else:
        lo, hi = w.astype(np.uint64), np.zeros(L, dtype=np.uint64)
    pos = np.arange(bits - 1, -1, -1, dtype=np.uint64)             # most significant bit of a field first
    src = np.where(pos >= 64, hi[:, None], lo[:, None])
    field_bits = ((src >> (pos % np.uint64(64))[None, :]) & np.uint64(1)).astype(np.uint8)
    stream = np.concatenate([np.zeros(8 * nbytes - L * bits, dtype=np.uint8), field_bits.reshape(-1)])
    return np.packbits(stream)


def wire_unpack(data, count, bits):
    """_from_bytes + reverse (framework/jzf_weights.py:98-137, 224) on the byte string; returns uint64 [L]
    (bits <= 64) or uint64 [L, 2]."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    nbytes = (count * bits + 7) // 8
    assert data.shape[0] == nbytes
    stream = np.unpackbits(data)[8 * nbytes - count * bits:].reshape(count, bits).astype(np.uint64)
    pos = np.arange(bits - 1, -1, -1, dtype=np.uint64)
    lo = (stream[:, pos < 64] << pos[pos < 64][None, :]).sum(axis=1, dtype=np.uint64)
    if bits <= 64:
        return lo
    hi = (stream[:, pos >= 64] << (pos[pos >= 64] - np.uint64(64))[None, :]).sum(axis=1, dtype=np.uint64)
    return np.stack([lo, hi], axis=1)


def sparsify_k(sparsity, size):
    """idx = max(1, int(np.floor(self._sparsity * size))) — jzf_aggregator.py:598."""
    return max(1, int(np.floor(sparsity * np.int64(size))))


def sparsify(layers, remain, sparsity):
    """Client.sparsify (proc/jzf_aggregator.py:578-623) over a list of float32 layers in walking order.
    Returns (compact values per layer, new residual per layer, global locations int64, base).  Ties at the
    k-th largest |x| are resolved as a STABLE ascending argsort followed by [-k:] does (highest indices
    win); the reference's default introsort leaves that choice unspecified."""
    out_vals, out_rem, locations, base = [], [], [], 0
    for i, layer in enumerate(layers):
        flatten = np.array(layer, dtype=np.float32).flatten()
        size = flatten.size
        abs_flatten = np.abs(flatten)
        if remain is not None:
            flatten += remain[i]
        k = sparsify_k(sparsity, size)
        location = sorted(abs_flatten.argsort(kind="stable")[-k:][::-1])
        out_vals.append(flatten[location].copy())
        flatten[location] = 0.0
        out_rem.append(flatten)
        locations += [int(l) + base for l in location]
        base += size
    return out_vals, out_rem, np.array(locations, dtype=np.int64), base


def unnormalize_stats(w, seg_end, shift, order="pairwise"):
    """QuantizingClient.unnormalize (sp/jzf_quantize.py:549-564) on a flat float64 vector of layers:
    returns (w + shift per layer, [(mean, std)] per layer) with numpy's own np.mean / np.std — the very
    functions the reference calls.  order="pairwise": float64 ndarrays (the batched mode); "sequential":
    object arrays of Python floats (the un-batched mode), which numpy sums left to right."""
    w = np.array(w, dtype=np.float64)
    stats, b = [], 0
    for e, s in zip(seg_end, shift):
        w[b:e] += s
        layer = w[b:e] if order == "pairwise" else w[b:e].astype(object)
        if e > b:
            stats.append((float(np.mean(layer)), float(np.std(layer))))
        else:
            stats.append((float("nan"), float("nan")))
        b = e
    return w, np.array(stats, dtype=np.float64)
