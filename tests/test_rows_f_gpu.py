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


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sparsify_many_random_layers(fb, seed):
    # runs of tiles that cross layer boundaries, partial tiles, aligned and unaligned layer starts, layers on both routes
    rs = np.random.RandomState(100 + seed)
    ctx = fb.DeviceContext(KEY, 32)
    sizes = [int(v) for v in rs.choice([1, 3, 17, 4095, 4096, 4097, 8192, 16384, 16385, 20000, 50001, 131072, 262147], size=24)]
    if seed == 2:
        sizes = [s - s % 4 + 4 for s in sizes]              # every layer 16-byte aligned: the ring path everywhere
    ends = np.cumsum(sizes)
    sparsity = [0.01, 0.003, 0.02][seed]
    ks = [O.sparsify_k(sparsity, s) for s in sizes]
    rem_d, rem_o = None, None
    for rnd in range(2):
        layers = [(rs.standard_normal(s) * rs.uniform(1e-3, 10.0)).astype(np.float32) for s in sizes]
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


def test_sparsify_small_and_large_layers_in_one_call(fb):
    # Layers whose radix passes run inside one block (k_topk_select_small: up to 4 M elements on the candidate route, 64 K on
    # the exact route) next to layers that take the multi-block passes (6 M and 5 M elements; 100 k at sparsity 0.5 = exact
    # route above the one-block limit): both kinds in the same launches, more than 8 layers (the one-block sample bound).
    rs = np.random.RandomState(31)
    ctx = fb.DeviceContext(KEY, 32)
    sizes = [6_000_000, 100_000, 3000, 5_000_001, 70_000, 9, 262_144, 40_000, 1_000_000]
    sparsity = [0.01, 0.5, 0.01, 0.002, 0.3, 0.5, 0.01, 0.02, 0.004]
    ends = np.cumsum(sizes)
    layers = [(rs.standard_normal(s) * rs.uniform(0.01, 3.0)).astype(np.float32) for s in sizes]
    ks = [O.sparsify_k(p, s) for p, s in zip(sparsity, sizes)]
    vals, idx, rem = ctx.topk_sparsify(_dev(np.concatenate(layers)), ends, ks)
    got_i, got_v, got_r = _np(idx), _np(vals), _np(rem)
    o = 0
    for li, (x, k) in enumerate(zip(layers, ks)):
        order = np.argsort(np.abs(x), kind="stable")[-k:]                 # jzf_aggregator.py:600-606: stable argsort, last k
        loc = np.sort(order)
        base = int(ends[li] - sizes[li])
        assert np.array_equal(got_i[o:o + k], loc + base), li
        assert np.array_equal(got_v[o:o + k].view(np.uint32), x[loc].view(np.uint32)), li
        r = x.copy(); r[loc] = 0.0
        assert np.array_equal(got_r[base:base + sizes[li]].view(np.uint32), r.view(np.uint32)), li
        o += k


@pytest.mark.parametrize("route", ["exact", "badbound"])
def test_sparsify_routes_agree(fb, route, monkeypatch):
    # The threshold is found either from a sampled candidate list (default for sparse selections) or from x itself.
    # FLASHE_TOPK_ROUTE=exact disables sampling on the host; =badbound forces a sample bound ABOVE the threshold, so
    # the device must notice the short candidate list and take the exact route by itself.  All three must agree bit
    # for bit with each other and with the reference's argsort order.
    rs = np.random.RandomState(21)
    ctx = fb.DeviceContext(KEY, 32)
    sizes = [3_000_001, 20_000, 4096, 250_000, 7, 1_000_000]
    ends = np.cumsum(sizes)
    layers = [(rs.standard_normal(s) * rs.uniform(0.01, 3.0)).astype(np.float32) for s in sizes]
    layers[3] = (layers[3] * 1e-3).astype(np.float32)
    layers[3][::5] = 0.125                                 # a fifth of one layer tied at its maximum: the candidate list overflows
    layers[5][:500_000] *= 40.0                            # structure a strided sample has to cope with
    x = _dev(np.concatenate(layers))
    res = _dev((rs.standard_normal(int(ends[-1])) * 0.01).astype(np.float32))
    ks = [O.sparsify_k(0.01, s) for s in sizes]
    want_v, want_r, want_loc, _ = O.sparsify(layers, list(np.split(_np(res), ends[:-1])), 0.01)
    monkeypatch.delenv("FLASHE_TOPK_ROUTE", raising=False)
    v0, i0, r0 = ctx.topk_sparsify(x, ends, ks, residual=res)
    assert np.array_equal(_np(i0), want_loc)
    assert np.array_equal(_np(v0).view(np.uint32), np.concatenate(want_v).view(np.uint32))
    assert np.array_equal(_np(r0).view(np.uint32), np.concatenate(want_r).view(np.uint32))
    monkeypatch.setenv("FLASHE_TOPK_ROUTE", route)
    v1, i1, r1 = ctx.topk_sparsify(x, ends, ks, residual=res)
    assert torch.equal(i0, i1) and torch.equal(v0.view(torch.int32), v1.view(torch.int32)) and torch.equal(r0.view(torch.int32), r1.view(torch.int32))


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
    """Bit-exact against what the reference's QuantizingClient.unnormalize left in past_layer_mean_list /
    past_layer_std_list, for float64 layers (pairwise order) and for object layers (sequential order)."""
    d, _ = gold
    ctx = fb.DeviceContext(KEY, 32)
    ends = np.cumsum(d["stats_sizes"])
    for order, tag in ((fb.SUM_PAIRWISE, "stats"), (fb.SUM_SEQUENTIAL, "stats_obj")):
        w_out, stats = ctx.segment_stats(_dev(d["stats_w"]), ends, d["stats_shift"], order=order,
                                         out=torch.empty(int(ends[-1]), dtype=torch.float64, device="cuda"))
        assert np.array_equal(_np(w_out).view(np.uint64), d[tag + "_w_out"].view(np.uint64))     # w + past_mean
        s = _np(stats)
        assert np.array_equal(s[:, 0].view(np.uint64), d[tag + "_mean"].view(np.uint64)), tag
        assert np.array_equal(s[:, 1].view(np.uint64), d[tag + "_std"].view(np.uint64)), tag
        # statistics only (no output vector) and in place give the same numbers
        _, s2 = ctx.segment_stats(_dev(d["stats_w"]), ends, d["stats_shift"], order=order)
        assert np.array_equal(_np(s2).view(np.uint64), s.view(np.uint64))
        w_in = _dev(d["stats_w"])
        _, s3 = ctx.segment_stats(w_in, ends, d["stats_shift"], inplace=True, order=order)
        assert np.array_equal(_np(s3).view(np.uint64), s.view(np.uint64)) and np.array_equal(_np(w_in), _np(w_out))


def test_segment_stats_bit_exact_vs_numpy_all_tree_shapes(fb):
    """numpy's pairwise recursion at every level of the device decomposition: layers below 8 elements, around
    the 128-element block, around the group (8192) and mid-node (262144) cuts, a 3 M-element layer, an empty
    layer; plus the sequential (object-array) order."""
    rs = np.random.RandomState(8)
    sizes = [3000001, 1, 999999, 4096 * 5, 0, 7, 8, 9, 127, 128, 129, 255, 257, 8191, 8192, 8193, 16385, 262143, 262144, 262145,
             524289, 2, 100003]
    ends = np.cumsum(sizes)
    w = rs.standard_normal(int(ends[-1])) * 0.3 + 1.5
    w[:1000] *= 1e6                                           # wide dynamic range: the order of the adds shows
    shift = [float(v) for v in rs.standard_normal(len(sizes)) * 0.01]
    ctx = fb.DeviceContext(KEY, 32)
    for order, name in ((fb.SUM_PAIRWISE, "pairwise"), (fb.SUM_SEQUENTIAL, "sequential")):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                    # numpy warns on the empty layer
            want_w, want = O.unnormalize_stats(w, ends, shift, order=name)
        got_w, stats = ctx.segment_stats(_dev(w), ends, shift, order=order, out=torch.empty(w.shape[0], dtype=torch.float64, device="cuda"))
        assert np.array_equal(_np(got_w).view(np.uint64), want_w.view(np.uint64))
        got = _np(stats)
        live = np.array(sizes) > 0
        assert np.array_equal(got[live].view(np.uint64), want[live].view(np.uint64)), name
        assert np.isnan(got[~live]).all()


def test_segment_stats_many_small_layers(fb):
    rs = np.random.RandomState(9)
    sizes = [int(v) for v in rs.randint(1, 700, size=400)]
    ends = np.cumsum(sizes)
    w = rs.standard_normal(int(ends[-1]))
    shift = [0.0] * len(sizes)
    ctx = fb.DeviceContext(KEY, 32)
    _, want = O.unnormalize_stats(w, ends, shift)
    _, stats = ctx.segment_stats(_dev(w), ends, shift)
    assert np.array_equal(_np(stats).view(np.uint64), want.view(np.uint64))


# ------------------------------------------------------------------------------------------- mirrors
class _W(object):
    def __init__(self, layers):
        self._weights = dict(layers)
        self.walking_order = sorted(self._weights.keys(), key=str)


def test_quantizing_client_whole_model_matches_reference(fb, gold):
    """The drop-in QuantizingClient + flatten / unflatten mirrors driven like the client code drives the
    reference's (jzf_aggregator.py:714-723, 896-904) on a five-layer model with the sparse sentinel layer,
    un-batched and batched: every quantised integer / 120-bit word, the flat vector, the decoded layers (dtype
    included: Python floats un-batched, float64 batched) and the refreshed per-layer mean / std lists —
    bit for bit against what the reference's own classes produced (tests/golden/make_golden_f.py)."""
    from flashe_b200.aggregate import flatten_weights, unflatten_weights
    from flashe_b200.secureprotol import QuantizingClient
    d, man = gold
    for c in [c for c in man["cases"] if c["name"].startswith("model_")]:
        tag, batch, n, keys = c["name"], c["batch"], c["n_clients"], c["keys"]
        shapes = {k: tuple(sh) for k, sh in zip(keys, c["shapes"])}
        x = d[tag + "_x"]
        offs = np.cumsum([0] + [int(np.prod(shapes[k])) for k in keys])
        w = _W({k: x[offs[i]:offs[i + 1]].reshape(shapes[k]).copy() for i, k in enumerate(keys)})
        qc = QuantizingClient(c["int_bits"], None, None, batch, 16, True, True)
        qc.num_clients = n
        qc.set_layer_size_list(w)
        qc.past_layer_std_list = [float(v) for v in d[tag + "_std_in"]]
        qc.past_layer_mean_list = [float(v) for v in d[tag + "_mean_in"]]
        np.random.seed(777 + int(batch))
        qc.normalize(w)
        qc.quantize(w)
        want_q = [int(lo) | (int(hi) << 64) for lo, hi in zip(d[tag + "_q_lo"], d[tag + "_q_hi"])]
        got_q = [int(v) for k in keys for v in np.asarray(w._weights[k], dtype=object).reshape(-1)]
        assert got_q == want_q, tag
        assert [int(np.asarray(w._weights[k]).size) for k in keys] == [int(v) for v in d[tag + "_q_sizes"]]
        w, shape_dict = flatten_weights(w)
        only = w.walking_order[0]
        assert len(w._weights[only]) == int(d[tag + "_flat_len"][0]) and "zzz" not in shape_dict
        w._weights[only] = w._weights[only] * n
        back = unflatten_weights(w, shape_dict)
        assert list(back.walking_order) == c["out_keys"]
        qc.unquantize(back)
        qc.unnormalize(back)
        got = np.concatenate([np.asarray(back._weights[k], dtype=np.float64).reshape(-1) for k in c["out_keys"]])
        assert np.array_equal(got.view(np.uint64), d[tag + "_out"].view(np.uint64)), tag
        assert [int(np.asarray(back._weights[k]).dtype == object) for k in c["out_keys"]] == [int(v) for v in d[tag + "_out_is_object"]]
        assert [tuple(back._weights[k].shape) for k in c["out_keys"]] == [shapes[k] for k in c["out_keys"]]
        assert np.array_equal(np.array([float(v) for v in qc.past_layer_mean_list]).view(np.uint64), d[tag + "_mean_out"].view(np.uint64)), tag
        assert np.array_equal(np.array([float(v) for v in qc.past_layer_std_list]).view(np.uint64), d[tag + "_std_out"].view(np.uint64)), tag


def test_batch_pack_layers_equals_per_layer_pack(fb):
    """The layer table version packs every layer by itself (zero padding per layer), in one launch."""
    ctx = fb.DeviceContext(KEY, 120)
    rs = np.random.RandomState(12)
    sizes = [1, 6, 7, 0, 600, 12, 5, 100003]
    ends = [int(v) for v in np.cumsum(sizes)]
    q = rs.randint(0, 65536, size=ends[-1]).astype(np.uint32)
    e, f = 16, 4                                              # lanes of 20 bits, 6 per word
    words = ctx.batch_pack_layers(_dev(q), ends, e, f)
    wend = ctx.batch_layout(ends, e, f)
    assert wend == [int(v) for v in np.cumsum([(n + 5) // 6 for n in sizes])]
    got = _np(words)
    b = wb = 0
    for n_, e_, we in zip(sizes, ends, wend):
        if n_:
            want = O.batch(q[b:e_], 120, e, f)
            assert np.array_equal(got[wb:we], want)
        b, wb = e_, we
    back = ctx.batch_unpack_layers(words, ends, e, f)
    assert np.array_equal(_np(back), q)


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
