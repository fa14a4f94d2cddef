"""GPU parity for the SURVEY §8 "next" rows through the C ABI: wire bit-packing (f1), layer-wise top-k
sparsification with residuals (f2), per-layer statistics around decode (f3) — against the fixtures made
by executing the reference's own source and against the numpy oracle on seeded inputs.  Bytes, indices
and float32 values bit-exact; the float64 statistics to 1e-12 relative (summation order differs)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KEY = bytes(range(32))


def _np(t):
    return t.cpu().numpy()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda")


@pytest.fixture(scope="module")
def fb():
    import flashe_b200
    return flashe_b200


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(HERE, "golden", "flashe_golden_f.npz"))
    return d, json.loads(bytes(d["manifest"]).decode())


def words_of(d, name, bits):
    lo, hi = d[name + "_lo"], d[name + "_hi"]
    if bits <= 32:
        return lo.astype(np.uint32)
    return lo if bits <= 64 else np.stack([lo, hi], axis=1)


# ------------------------------------------------------------------------------------------- f1
def test_wire_golden(fb, gold):
    d, man = gold
    ctx = fb.DeviceContext(KEY, 32)
    for c in man["cases"]:
        if not c["name"].startswith("wire_") or c.get("ref_raises"):
            continue
        bits, L = c["bits"], c["L"]
        w = words_of(d, c["name"], bits)
        packed = ctx.wire_pack(_dev(w), bits)
        assert np.array_equal(_np(packed), d[c["name"] + "_bytes"]), c
        back = _np(ctx.wire_unpack(packed, L, bits))
        assert np.array_equal(back, w), c


@pytest.mark.parametrize("bits,L,wb", [(20, 1, 4), (20, 2, 4), (20, 1000003, 4), (32, 999999, 4), (24, 65537, 4), (8, 4097, 4),
                                       (1, 77, 4), (31, 12345, 4), (26, 500000, 8), (27, 300001, 8), (33, 70001, 8),
                                       (64, 50001, 8), (7, 1000, 8), (120, 40001, 16), (128, 5003, 16), (65, 9999, 16),
                                       (100, 3, 16), (20, 999, 16),
                                       # tiled 4-byte kernels: several tiles, every tail length, narrow and odd widths
                                       (12, 100001, 4), (9, 33333, 4), (17, 262149, 4), (32, 4096, 4), (16, 16384, 4),
                                       (25, 1, 4), (3, 5000, 4), (20, 4096 * 3 + 1, 4), (29, 8191, 4), (8, 16, 4), (13, 7, 4)])
def test_wire_vs_oracle(fb, bits, L, wb):
    rs = np.random.RandomState(bits * 1000 + L % 1000)
    ctx = fb.DeviceContext(KEY, 32)
    lo = rs.randint(0, 2 ** 63, L, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rs.randint(0, 2, L).astype(np.uint64)
    hi = rs.randint(0, 2 ** 63, L, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rs.randint(0, 2, L).astype(np.uint64)
    if bits < 64:
        lo &= np.uint64((1 << bits) - 1)
    if bits <= 64:
        hi[:] = 0
    elif bits < 128:
        hi &= np.uint64((1 << (bits - 64)) - 1)
    w = lo.astype(np.uint32) if wb == 4 else (lo if wb == 8 else np.stack([lo, hi], axis=1))
    want = O.wire_pack(w, bits)
    got = ctx.wire_pack(_dev(w), bits)
    assert np.array_equal(_np(got), want)
    back = _np(ctx.wire_unpack(got, L, bits, word_bytes=wb))
    assert np.array_equal(back, w)


def test_wire_unaligned_views(fb):
    # source / destination views that start off a 16-byte boundary
    ctx = fb.DeviceContext(KEY, 32)
    for bits, L, off in ((20, 70001, 1), (24, 9000, 3), (32, 5001, 2)):
        w = np.random.RandomState(L).randint(0, 1 << bits, L + off, dtype=np.int64).astype(np.uint32)
        t = _dev(w)[off:]
        got = ctx.wire_pack(t, bits)
        assert np.array_equal(_np(got), O.wire_pack(w[off:], bits)), (bits, L, off)
        assert np.array_equal(_np(ctx.wire_unpack(got, L, bits)), w[off:])


def test_wire_ciphertext_roundtrip_matches_python_bigint(fb):
    # the wire integer of a real ciphertext: int.from_bytes(stream, 'big') == sum ct[j] << ((L-1-j)*b)
    ctx = fb.DeviceContext(KEY, 20)
    L = 3001
    q = _dev(np.random.RandomState(3).randint(0, 65536, L).astype(np.uint32))
    ct = ctx.encrypt(1, 2, fb.SCHEME_DOUBLE, q, fb.VectorSpan(L, 8))
    s = 0
    for v in _np(ct):
        s = (s << 20) + int(v)
    assert int.from_bytes(bytes(_np(ctx.wire_pack(ct))), "big") == s
    assert np.array_equal(_np(ctx.wire_unpack(ctx.wire_pack(ct), L)), _np(ct))


# ------------------------------------------------------------------------------------------- f2
def test_sparsify_golden(fb, gold):
    d, man = gold
    ctx = fb.DeviceContext(KEY, 32)
    remain = {}
    for c in [c for c in man["cases"] if c["name"].startswith("sp_")]:
        ci = c["name"].split("_")[1]
        if c["round"] == 0:
            remain[ci] = None
        sizes = c["sizes"]
        ends = np.cumsum(sizes)
        ks = [O.sparsify_k(c["sparsity"], s) for s in sizes]
        vals, idx, rem = ctx.topk_sparsify(_dev(d[c["name"] + "_x"]), ends, ks, residual=remain[ci])
        assert np.array_equal(_np(vals).view(np.uint32), d[c["name"] + "_values"].view(np.uint32)), c
        assert np.array_equal(_np(rem).view(np.uint32), d[c["name"] + "_remain"].view(np.uint32)), c
        assert idx.numel() == c["le"]
        # Client.sparsify returns _to_bytes(locations, base.bit_length())
        loc_bytes = ctx.wire_pack(idx.view(torch.uint64), c["bits"])
        assert np.array_equal(_np(loc_bytes), d[c["name"] + "_locbytes"]), c
        remain[ci] = rem


@pytest.mark.parametrize("sizes,sparsity", [([1000003], 0.01), ([5, 70000, 4096, 1, 123457], 0.01), ([300000, 300000], 0.25),
                                            ([4097], 1.0), ([17], 0.0001)])
def test_sparsify_vs_oracle_with_residual_rounds(fb, sizes, sparsity):
    rs = np.random.RandomState(len(sizes) * 7 + sizes[0] % 97)
    ctx = fb.DeviceContext(KEY, 32)
    ends = np.cumsum(sizes)
    ks = [O.sparsify_k(sparsity, s) for s in sizes]
    rem_d, rem_o = None, None
    for rnd in range(3):
        layers = [(rs.standard_normal(s) * 0.1).astype(np.float32) for s in sizes]
        want_v, rem_o, want_loc, _ = O.sparsify(layers, rem_o, sparsity)
        vals, idx, rem_d = ctx.topk_sparsify(_dev(np.concatenate(layers)), ends, ks, residual=rem_d)
        assert np.array_equal(_np(idx), want_loc)
        assert np.array_equal(_np(vals).view(np.uint32), np.concatenate(want_v).view(np.uint32))
        assert np.array_equal(_np(rem_d).view(np.uint32), np.concatenate(rem_o).view(np.uint32))


def test_sparsify_ties_zeros_and_specials(fb):
    # heavy ties (quantised magnitudes, signed zeros) and inf / nan: stable-argsort order, highest index first
    rs = np.random.RandomState(11)
    ctx = fb.DeviceContext(KEY, 32)
    n = 50000
    x = (rs.randint(-3, 4, n) * 0.25).astype(np.float32)
    x[rs.randint(0, n, 50)] = -0.0
    x[123] = np.inf; x[4567] = -np.inf; x[n - 1] = np.nan
    for k in (1, 2, 3, 1000, 20000, n - 1, n):
        want_v, want_r, want_loc, _ = O.sparsify([x], None, k / n + 1e-12)
        assert want_loc.shape[0] == k
        vals, idx, rem = ctx.topk_sparsify(_dev(x), [n], [k])
        assert np.array_equal(_np(idx), want_loc), k
        assert np.array_equal(_np(vals).view(np.uint32), want_v[0].view(np.uint32)), k
        assert np.array_equal(_np(rem).view(np.uint32), want_r[0].view(np.uint32)), k


def test_sparse_round_c4_shape_end_to_end(fb):
    # sparsify -> encode compact -> single-mask encrypt -> expand + zero fill -> sum -> sparse decrypt == sum of q
    total, n, b, n_jobs = 200000, 4, 32, 8
    rs = np.random.RandomState(5)
    ctx = fb.DeviceContext(KEY, b)
    k = O.sparsify_k(0.01, total)
    dense_sum = np.zeros(total, dtype=np.uint64)
    agg = None
    index_lists = []
    for c in range(n):
        x = (rs.standard_normal(total) * 0.1).astype(np.float32)
        vals, idx, _ = ctx.topk_sparsify(_dev(x), [total], [k])
        span = fb.VectorSpan(k, n_jobs)
        u = _dev(rs.random_sample(k))
        q = ctx.encode(vals, fb.CodecSpec(alpha=1.0, element_bits=16), fb.NoiseSpec(u=u), span)
        ct = ctx.encrypt(0, c, fb.SCHEME_SINGLE, q, span)
        zero = int(O.quantize(np.zeros(1, np.float32), np.zeros(1), 1.0, 16)[0])
        dense = ctx.sparse_expand(ct, idx, total, zero)
        agg = dense if agg is None else ctx.aggregate(torch.stack([agg.view(torch.int32), dense.view(torch.int32)]).view(torch.uint32))
        want_dense = np.full(total, zero, dtype=np.uint64)
        want_dense[_np(idx)] = _np(q)
        dense_sum += want_dense
        index_lists.append(idx)
    for c in range(n):
        ctx.sparse_apply_masks(0, [c], [-1], fb.VectorSpan(k, n_jobs), index_lists[c], agg)
    assert np.array_equal(_np(agg).astype(np.uint64), dense_sum & np.uint64(2 ** b - 1))


# ------------------------------------------------------------------------------------------- f3
def test_segment_stats_golden(fb, gold):
    d, _ = gold
    ctx = fb.DeviceContext(KEY, 32)
    ends = np.cumsum(d["stats_sizes"])
    w_out, stats = ctx.segment_stats(_dev(d["stats_w"]), ends, d["stats_shift"], out=torch.empty(int(ends[-1]), dtype=torch.float64, device="cuda"))
    assert np.array_equal(_np(w_out).view(np.uint64), d["stats_w_out"].view(np.uint64))     # w + past_mean: exact
    s = _np(stats)
    np.testing.assert_allclose(s[:, 0], d["stats_mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(s[:, 1], d["stats_std"], rtol=1e-12, atol=1e-15)
    # statistics only (no output vector) and in place give the same numbers
    _, s2 = ctx.segment_stats(_dev(d["stats_w"]), ends, d["stats_shift"])
    assert np.array_equal(_np(s2), s)
    w_in = _dev(d["stats_w"])
    _, s3 = ctx.segment_stats(w_in, ends, d["stats_shift"], inplace=True)
    assert np.array_equal(_np(s3), s) and np.array_equal(_np(w_in), _np(w_out))


def test_segment_stats_large_vs_numpy(fb):
    rs = np.random.RandomState(8)
    sizes = [3000001, 1, 999999, 4096 * 5]
    ends = np.cumsum(sizes)
    w = rs.standard_normal(int(ends[-1])) * 0.3 + 1.5
    shift = [0.0, -2.0, 0.25, 1e-3]
    ctx = fb.DeviceContext(KEY, 32)
    want_w, want = O.unnormalize_stats(w, ends, shift)
    got_w, stats = ctx.segment_stats(_dev(w), ends, shift, out=torch.empty(w.shape[0], dtype=torch.float64, device="cuda"))
    assert np.array_equal(_np(got_w).view(np.uint64), want_w.view(np.uint64))
    np.testing.assert_allclose(_np(stats), want, rtol=1e-12, atol=1e-15)
    # deterministic: same bits on a second run
    _, again = ctx.segment_stats(_dev(w), ends, shift)
    assert np.array_equal(_np(again), _np(stats))


# ------------------------------------------------------------------------------------------- mirrors
class _W(object):
    def __init__(self, layers):
        self._weights = dict(layers)
        self.walking_order = sorted(self._weights.keys(), key=str)


def test_reference_shaped_mirrors(fb, gold):
    """flashe_b200.weights._to_bytes/_from_bytes and aggregate.SparsifyingClient.sparsify reproduce the
    reference's return values (Python ints, reversed lists, compact layers, residual dict)."""
    from flashe_b200 import weights as W
    from flashe_b200.aggregate import SparsifyingClient
    d, man = gold
    for c in man["cases"]:
        if not c["name"].startswith("wire_") or c.get("ref_raises"):
            continue
        bits, L = c["bits"], c["L"]
        lo, hi = d[c["name"] + "_lo"], d[c["name"] + "_hi"]
        vals = [int(a) | (int(b) << 64) for a, b in zip(lo, hi)]
        s, l = W._to_bytes(np.array(vals, dtype=object), bits)
        assert l == L and s == int.from_bytes(bytes(d[c["name"] + "_bytes"]), "big"), c
        back = W._from_bytes(s, L, bits)
        back.reverse()
        assert back == vals, c
    keys = ["a_conv", "b_dense", "c_bias", "d_one", "e_big"]
    for ci in ("0", "1", "2"):
        cl = None
        for c in [c for c in man["cases"] if c["name"].startswith("sp_%s_" % ci)]:
            if cl is None:
                cl = SparsifyingClient(c["sparsity"])
            offs = np.cumsum([0] + c["sizes"])
            x = d[c["name"] + "_x"]
            w = _W({k: x[offs[i]:offs[i + 1]].copy() for i, k in enumerate(keys)})
            enc, le, bits, base = cl.sparsify(w)
            assert (le, bits, base) == (c["le"], c["bits"], c["base"])
            assert enc == int.from_bytes(bytes(d[c["name"] + "_locbytes"]), "big")
            got_v = np.concatenate([w._weights[k] for k in keys])
            got_r = np.concatenate([cl.remain_weights[k] for k in keys])
            assert np.array_equal(got_v.view(np.uint32), d[c["name"] + "_values"].view(np.uint32))
            assert np.array_equal(got_r.view(np.uint32), d[c["name"] + "_remain"].view(np.uint32))


# ------------------------------------------------------------------------------------------- online step after precompute
@pytest.mark.parametrize("bits,L,begin,count", [(32, 100003, 0, None), (32, 100003, 4000, 50001), (32, 100003, 4001, 50000),
                                                (20, 70001, 0, None), (64, 30001, 0, None), (120, 9001, 0, None), (32, 3, 0, None)])
def test_encode_add_premasked_equals_fused_encrypt(fb, bits, L, begin, count):
    """prepare_encrypt + online add (jzf_flashe.py:599-631, 480-486) == on-the-fly encode+encrypt, for the
    vectorised 4-byte path, its scalar tail, unaligned shards and the wide words; layers with own alphas."""
    rs = np.random.RandomState(L + bits)
    ctx = fb.DeviceContext(KEY, bits)
    cnt = L - begin if count is None else count
    span = fb.VectorSpan(L, 8, begin, cnt)
    seg_end = [L // 3, L // 3 + 5, L] if L > 20 else [L]
    codec = fb.CodecSpec(alpha=[0.3, 0.05, 0.7][:len(seg_end)], element_bits=16, seg_end=seg_end)
    x = _dev((rs.standard_normal(cnt) * 0.2).astype(np.float32))
    mask = ctx.masks(5, [3, 4], [1, -1], span)
    for noise in (fb.NoiseSpec(seed=77, stream=3), fb.NoiseSpec(u=_dev(rs.random_sample(cnt)))):
        want = ctx.encode_encrypt(5, 3, fb.SCHEME_DOUBLE, x, codec, noise, span)
        got = ctx.encode_add_premasked(x, codec, noise, mask, span)
        assert torch.equal(got.view(torch.int32), want.view(torch.int32))
