"""GPU parity: the CUDA path (through the C ABI) against the committed reference outputs
(tests/golden) and against the CPU oracle on seeded inputs.  Bit-exact for every integer; decoded
float64 compared bit for bit (tolerance 0)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

KEY = bytes(range(32))


def _np(t):
    return t.cpu().numpy()


def _dev(a, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


@pytest.fixture(scope="module")
def fb():
    import flashe_b200
    return flashe_b200


def ctx_for(fb, bits, key=KEY):
    return fb.DeviceContext(key, bits)


# ---------------------------------------------------------------------------------------- AES
def test_prp_fips197_and_random_blocks(fb):
    ctx = ctx_for(fb, 32)
    pt = bytes.fromhex("00112233445566778899aabbccddeeff")
    assert ctx.prp_block(pt).hex() == "8ea2b7ca516745bfeafc49904b496089"
    rs = np.random.RandomState(5)
    for _ in range(20):
        blk = bytes(rs.randint(0, 256, 16).astype(np.uint8))
        if rs.rand() < 0.5:
            blk = blk[:8] + bytes(4) + blk[12:]      # word 2 == 0: the hoisted round-1 path
        assert ctx.prp_block(blk) == O.aes256_encrypt_block(KEY, blk)
    long_seed = bytes(224) + KEY                     # hosts carry a 256-byte padded copy of the seed
    assert fb.DeviceContext(long_seed, 32).prp_block(pt) == ctx.prp_block(pt)


# ---------------------------------------------------------------------------------------- masks
def test_masks_golden(fb, golden):
    for c in golden.cases("masks"):
        b = c["int_bits"]
        ctx = ctx_for(fb, b)
        span = fb.VectorSpan(c["L"], c["n_jobs"])
        got = _np(ctx.masks(c["iter"], [c["prf_idx"]], [1], span))
        assert np.array_equal(got, golden.words(c["name"], b)), c


@pytest.mark.parametrize("bits,n_jobs,L", [(20, 8, 100003), (32, 1024, 250000), (24, 3, 7777), (64, 8, 30011),
                                           (120, 8, 20001), (12, 8, 9999), (33, 5, 5000), (20, 64, 50)])
def test_masks_vs_oracle_multistream_and_shards(fb, bits, n_jobs, L):
    ctx = ctx_for(fb, bits)
    idx, sign = [3, 4, 9, 0], [1, -1, -1, 1]
    want = O.masks(KEY, bits, n_jobs, 7, idx, sign, L)
    got = _np(ctx.masks(7, idx, sign, fb.VectorSpan(L, n_jobs)))
    assert np.array_equal(got, want)
    # element-range shards (multi-GPU layout): 3 uneven pieces, arbitrary (unaligned) cuts
    cuts = [0, L // 3 + 1, (2 * L) // 3 + 2, L]
    for a, e in zip(cuts[:-1], cuts[1:]):
        part = _np(ctx.masks(7, idx, sign, fb.VectorSpan(L, n_jobs, a, e - a)))
        assert np.array_equal(part, want[a:e]), (a, e)
    # empty shard is a no-op
    assert ctx.masks(7, idx, sign, fb.VectorSpan(L, n_jobs, 5, 0)).numel() == 0


def test_counter_above_2_32_uses_generic_round1(fb):
    # chunk begins beyond 2^32: the AES counter's high word is non-zero
    ctx = ctx_for(fb, 32)
    L, n_jobs = (1 << 33) + 1000, 2
    begin = (1 << 32) + 517          # inside chunk 1 (starts at 2^32 + 500)
    got = _np(ctx.masks(1, [2], [1], fb.VectorSpan(L, n_jobs, begin, 4096)))
    want = O.masks(KEY, 32, n_jobs, 1, [2], [1], L, begin, 4096)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------- golden round trips
def test_roundtrip_golden(fb, golden):
    for c in golden.cases("roundtrip"):
        name, b, nj, L, n, it, scheme = (c[k] for k in ("name", "int_bits", "n_jobs", "L", "n_clients", "iter", "scheme"))
        ctx = ctx_for(fb, b)
        span = fb.VectorSpan(L, nj)
        sch = fb.SCHEME_DOUBLE if scheme == "double" else fb.SCHEME_SINGLE
        codec = fb.CodecSpec(alpha=c["alpha"], element_bits=c["element_bits"], n_clients=n)
        x, u = _dev(golden[name + "_x"]), _dev(golden[name + "_u"])
        q_ref = golden.words(name + "_q", 32).reshape(n, L)
        ct_ref = golden.words(name + "_ct", b).reshape(n, L)
        cts = ctx.empty_words(L, rows=n)
        for k in range(n):
            q_out = torch.empty(L, dtype=torch.uint32, device="cuda")
            ctx.encode_encrypt(it, k, sch, x[k], codec, fb.NoiseSpec(u=u[k]), span, out=cts[k], q_out=q_out)
            assert np.array_equal(_np(q_out), q_ref[k]), (name, k)
            assert np.array_equal(_np(ctx.encode(x[k], codec, fb.NoiseSpec(u=u[k]), span)), q_ref[k])
            assert np.array_equal(_np(ctx.encrypt(it, k, sch, _dev(q_ref[k].astype(ct_ref.dtype)), span)), ct_ref[k])
        assert np.array_equal(_np(cts), ct_ref), name
        # all clients in one launch, with and without stream sharing
        for share in (False, True):
            got = ctx.encode_encrypt_batch(it, 0, sch, x, codec, fb.NoiseSpec(u=u.reshape(-1)), span, share_streams=share)
            assert np.array_equal(_np(got), ct_ref), (name, share)
        agg_b = ctx.aggregate(cts, fb.AGG_ELEMENTWISE)
        agg_a = ctx.aggregate(cts, fb.AGG_PACKED)
        assert np.array_equal(_np(agg_b), golden.words(name + "_aggB", b)), name
        assert np.array_equal(_np(agg_a), golden.words(name + "_aggA", b)), name
        if scheme == "double":
            add, minus = [n], [0]
        else:
            add, minus = [], list(range(n))
        for tag, agg in (("B", agg_b), ("A", agg_a)):
            dec = ctx.decrypt(it, add, minus, agg, span)
            assert np.array_equal(_np(dec), golden.words(name + "_dec" + tag, b)), (name, tag)
        p_out = ctx.empty_words(L)
        decoded = ctx.decrypt_decode(it, add, minus, agg_b, codec, span, p_out=p_out)
        assert np.array_equal(_np(p_out), golden.words(name + "_decB", b))
        assert np.array_equal(_np(decoded).view(np.uint64), golden[name + "_decoded"].view(np.uint64)), name
        assert np.array_equal(_np(ctx.decode(p_out, codec, span)).view(np.uint64), golden[name + "_decoded"].view(np.uint64))


def test_dropout_golden(fb, golden):
    from flashe_b200.secureprotol.flashe import collapse_runs
    cases = golden.cases("dropout")
    c0 = cases[0]
    b, nj, L, n, it = (c0[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter"))
    ctx = ctx_for(fb, b)
    span = fb.VectorSpan(L, nj)
    ct = _dev(golden.words("drop_ct", b).reshape(n, L))
    for c in cases:
        add, minus = collapse_runs(c["survivors"])
        assert add == c["add"] and minus == c["minus"]
        rows = ct.view(torch.int32)[sorted(c["survivors"])].contiguous()   # (index_cuda has no uint32 kernel)
        agg = ctx.aggregate(rows)
        assert np.array_equal(_np(agg), golden.words(c["name"] + "_agg", b))
        assert np.array_equal(_np(ctx.decrypt(it, add, minus, agg, span)), golden.words(c["name"] + "_dec", b))


def test_precompute_golden(fb, golden):
    c = golden.cases("precompute")[0]
    b, nj, L, n, it, idx = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter", "idx"))
    ctx = ctx_for(fb, b)
    span = fb.VectorSpan(L, nj)
    assert np.array_equal(_np(ctx.masks(it, [idx], [1], span)), golden.words("pre_enc_add", b))
    assert np.array_equal(_np(ctx.masks(it, [idx + 1], [1], span)), golden.words("pre_enc_minus", b))
    assert np.array_equal(_np(ctx.masks(it, [n], [1], span)), golden.words("pre_dec_add", b))
    assert np.array_equal(_np(ctx.masks(it, [0], [1], span)), golden.words("pre_dec_minus", b))
    combined = ctx.masks(it, [idx, idx + 1], [1, -1], span)
    q = _dev(golden.words("pre_q", 32))
    assert np.array_equal(_np(ctx.add_premasked(q, combined, +1)), golden.words("pre_ct", b))
    assert np.array_equal(_np(ctx.add_premasked(_dev(golden.words("pre_ct", b)), combined, -1)), golden.words("pre_q", 32))


def test_sparse_single_golden(fb, golden):
    c = golden.cases("sparse")[0]
    b, nj, n, it, total = (c[k] for k in ("int_bits", "n_jobs", "n_clients", "iter", "total"))
    ctx = ctx_for(fb, b)
    idx = [_dev(golden["sparse_mask_%d" % k]) for k in range(n)]
    dense = []
    for k in range(n):
        q = _dev(golden.words("sparse_q_%d" % k, 32))
        ct = ctx.encrypt(it, k, fb.SCHEME_SINGLE, q, fb.VectorSpan(q.numel(), nj))
        assert np.array_equal(_np(ct), golden.words("sparse_ct_%d" % k, b))
        dense.append(ctx.sparse_expand(ct, idx[k], total, int(golden["sparse_zero"][k])))
    agg = ctx.aggregate(torch.stack([d.view(torch.int32) for d in dense]))
    assert np.array_equal(_np(agg), golden.words("sparse_agg", b))
    minus = ctx.zeros_words(total)
    for k in range(n):
        ctx.sparse_apply_masks(it, [k], [1], fb.VectorSpan(idx[k].numel(), nj), idx[k], minus)
    dec = ctx.add_premasked(agg, minus, -1)
    assert np.array_equal(_np(dec), golden.words("sparse_dec", b))
    # overlap counts behind dynamic_masking
    want = [len(set(_np(idx[i]).tolist()) & set(_np(idx[i + 1]).tolist())) for i in range(n - 1)]
    assert ctx.sparse_overlap(idx, total) == want


def test_batch120_golden(fb, golden):
    c = golden.cases("batch")[0]
    b, nj, L, n, it, e, f = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter", "element_bits", "factor"))
    nw = c["words"]
    ctx = ctx_for(fb, b)
    span = fb.VectorSpan(nw, nj)
    codec = fb.CodecSpec(alpha=c["alpha"], element_bits=e, n_clients=n)
    ct_ref = golden.words("batch_ct", b).reshape(n, nw, 2)
    w_ref = golden.words("batch_w", b).reshape(n, nw, 2)
    cts = ctx.empty_words(nw, rows=n)
    for k in range(n):
        q = ctx.encode(_dev(golden["batch_x"][k]), codec, fb.NoiseSpec(u=_dev(golden["batch_u"][k])), fb.VectorSpan(L, 1))
        w = ctx.batch_pack(q, e, f)
        assert np.array_equal(_np(w), w_ref[k])
        ctx.encrypt(it, k, fb.SCHEME_DOUBLE, w, span, out=cts[k])
    assert np.array_equal(_np(cts), ct_ref)
    agg = ctx.aggregate(cts)
    assert np.array_equal(_np(agg), golden.words("batch_agg", b))
    dec = ctx.decrypt(it, [n], [0], agg, span)
    assert np.array_equal(_np(dec), golden.words("batch_dec", b))
    unb = ctx.batch_unpack(dec, e, f)[:L].contiguous()
    assert np.array_equal(_np(unb), golden.words("batch_unb", 32))
    out = fb.DeviceContext(KEY, 32).decode(unb, codec, fb.VectorSpan(L, 1))
    assert np.array_equal(_np(out).view(np.uint64), golden["batch_decoded"].view(np.uint64))
    # the shipped dense path: packed-carry sum of the 120-bit words (jzf_aggregator.py:404-419)
    agg_a = ctx.aggregate(cts, fb.AGG_PACKED)
    assert np.array_equal(_np(agg_a), golden.words("batch_aggA", b))
    dec_a = ctx.decrypt(it, [n], [0], agg_a, span)
    assert np.array_equal(_np(dec_a), golden.words("batch_decA", b))


def test_batch120_fused_golden(fb, golden):
    """The shipped batch mode through the FUSED entries: float32 layer -> encode -> 6-lane pack -> mask in one
    launch per client (flashe_encode_encrypt with a lane-batching codec), and unmask -> unbatch -> decode in one
    launch (flashe_decrypt_decode on 128-bit words) — against the reference's own outputs."""
    c = golden.cases("batch")[0]
    b, nj, L, n, it, e, f = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter", "element_bits", "factor"))
    nw = c["words"]
    ctx = ctx_for(fb, b)
    span = fb.VectorSpan(nw, nj)
    codec = fb.CodecSpec(alpha=[c["alpha"]], element_bits=e, n_clients=n, seg_end=[L], batch_lane_bits=e + f)
    ct_ref = golden.words("batch_ct", b).reshape(n, nw, 2)
    cts = ctx.empty_words(nw, rows=n)
    for k in range(n):
        q = torch.empty(L, dtype=torch.int32, device="cuda").view(torch.uint32)
        ctx.encode_encrypt(it, k, fb.SCHEME_DOUBLE, _dev(golden["batch_x"][k]), codec, fb.NoiseSpec(u=_dev(golden["batch_u"][k])), span,
                           out=cts[k], q_out=q)
        assert np.array_equal(_np(q), golden.words("batch_q", 32).reshape(n, L)[k])
    assert np.array_equal(_np(cts), ct_ref)
    # all clients in one launch, with and without shared streams
    x_all, u_all = _dev(golden["batch_x"]), _dev(golden["batch_u"]).reshape(-1)
    for share in (False, True):
        got = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x_all, codec, fb.NoiseSpec(u=u_all), span, share_streams=share)
        assert np.array_equal(_np(got), ct_ref), share
    for mode, tag in ((fb.AGG_ELEMENTWISE, ""), (fb.AGG_PACKED, "A")):
        agg = ctx.aggregate(cts, mode)
        assert np.array_equal(_np(agg), golden.words("batch_agg" + tag, b))
        p = ctx.empty_words(nw)
        out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
        assert np.array_equal(_np(p), golden.words("batch_dec" + tag, b))
        if not tag:
            assert np.array_equal(_np(out).view(np.uint64), golden["batch_decoded"].view(np.uint64))


@pytest.mark.parametrize("n_jobs,sizes", [(8, [100, 7, 4097, 1, 3000, 6, 12]), (3, [250_001, 5, 120_000]), (16, [5]), (5, [2_000_000])])
def test_batch_fused_multi_layer_vs_oracle(fb, n_jobs, sizes):
    """Lane-batched model of several layers (each padded by itself): fused encode+pack+encrypt of 3 clients,
    fused decrypt+unbatch+decode, whole vector and two word-range shards, seeded and device noise — against the
    oracle's quantize / batch / encrypt / decrypt / unbatch / unquantize chain per layer."""
    b, e, n, it = 120, 16, 3, 5
    f = int(np.ceil(np.log2(n)))
    bs = b // (e + f)
    ends = [int(v) for v in np.cumsum(sizes)]
    total = ends[-1]
    alphas = [0.3 + 0.05 * i for i in range(len(sizes))]
    ctx = ctx_for(fb, b)
    codec = fb.CodecSpec(alpha=alphas, element_bits=e, n_clients=n, seg_end=ends, batch_lane_bits=e + f)
    wends = codec.word_ends(b)
    nw = wends[-1]
    assert wends == ctx.batch_layout(ends, e, f)
    span = fb.VectorSpan(nw, n_jobs)
    rs = np.random.RandomState(total % 1000)
    x = (rs.standard_normal((n, total)) * 0.2).astype(np.float32)
    u = rs.random_sample((n, total))
    O.set_threads(8)

    def oracle_words(xc, uc):
        parts, bgn = [], 0
        for en, a in zip(ends, alphas):
            parts.append(O.batch(O.quantize(xc[bgn:en], uc[bgn:en], a, e), b, e, f))
            bgn = en
        return np.concatenate(parts)

    want_ct = np.stack([O.encrypt(KEY, b, n_jobs, it, k, "double", oracle_words(x[k], u[k])) for k in range(n)])
    got = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(u=_dev(u).reshape(-1)), span)
    assert np.array_equal(_np(got), want_ct)
    # device noise: the documented Philox stream, one stream id per client
    got_dn = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(seed=9, stream=4), span)
    for k in range(n):
        uk = _np(ctx.rng_uniform(9, 4 + k, 0, total))
        assert np.array_equal(_np(got_dn[k]), O.encrypt(KEY, b, n_jobs, it, k, "double", oracle_words(x[k], uk))), k
    want_agg = O.aggregate(b, want_ct)
    want_p = O.decrypt(KEY, b, n_jobs, it, list(range(n)), "double", want_agg)
    want_out, bgn, wb = [], 0, 0
    for en, we, a in zip(ends, wends, alphas):
        lanes = O.unbatch(want_p[wb:we], b, e, f)[:en - bgn]
        want_out.append(O.unquantize(lanes, a, e, n))
        bgn, wb = en, we
    want_out = np.concatenate(want_out)
    agg = ctx.aggregate(got)
    p = ctx.empty_words(nw)
    out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
    assert np.array_equal(_np(p), want_p)
    assert np.array_equal(_np(out).view(np.uint64), want_out.view(np.uint64))
    # two word-range shards: element-indexed buffers start at the shard's first element
    if nw >= 4:
        cut = nw // 2 + 1
        for lo, cnt in ((0, cut), (cut, nw - cut)):
            sp = fb.VectorSpan(nw, n_jobs, lo, cnt)
            e0, ne = codec.span_elements(b, sp)
            xs = np.ascontiguousarray(x[:, e0:e0 + ne]); us = np.ascontiguousarray(u[:, e0:e0 + ne])
            part = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(xs), codec, fb.NoiseSpec(u=_dev(us).reshape(-1)), sp)
            assert np.array_equal(_np(part), want_ct[:, lo:lo + cnt]), (lo, cnt)
            o2 = ctx.decrypt_decode(it, [n], [0], _dev(want_agg[lo:lo + cnt]), codec, sp)
            assert np.array_equal(_np(o2).view(np.uint64), want_out[e0:e0 + ne].view(np.uint64)), (lo, cnt)


def test_packed_aggregate_wide_words_golden(fb, golden):
    for c in golden.cases("packed_wide"):
        b, n, L = c["int_bits"], c["n_clients"], c["L"]
        ctx = ctx_for(fb, b)
        cts = golden.words(c["name"] + "_ct", b).reshape(n, L, 2)
        got = ctx.aggregate(_dev(cts), fb.AGG_PACKED)
        assert np.array_equal(_np(got), golden.words(c["name"] + "_aggA", b)), c["name"]


@pytest.mark.parametrize("bits", [65, 96, 120, 127, 128])
def test_packed_aggregate_wide_adversarial_and_shards(fb, bits):
    """16-byte words: all-ones digits (every word passes its carry-in on, across tile boundaries),
    digits that generate and propagate, two element-range shards with the descriptor exchange."""
    ctx = ctx_for(fb, bits)
    rs = np.random.RandomState(bits)
    L, n = 3000, 7
    full = (1 << bits) - 1
    m64 = (1 << 64) - 1
    cts = rs.randint(0, 1 << 62, size=(n, L, 2), dtype=np.uint64)
    cts[:, :, 0] |= rs.randint(0, 4, size=(n, L), dtype=np.uint64) << np.uint64(62)
    cts[:, :, 1] &= np.uint64(full >> 64)
    cts[:, 500:2200] = 0
    cts[0, 500:2200, 0] = np.uint64(m64); cts[0, 500:2200, 1] = np.uint64(full >> 64)   # only propagate
    cts[1, 2199, 0] = 1                                                                  # one carry enters
    cts[:, 2500:2600, 0] = np.uint64(m64); cts[:, 2500:2600, 1] = np.uint64(full >> 64)  # generate and propagate
    want, cw = O.aggregate(bits, cts, "packed", return_carry=True)
    desc = torch.zeros(4, dtype=torch.int32, device="cuda")
    got = ctx.aggregate(_dev(cts), fb.AGG_PACKED, carry_out=desc)
    assert np.array_equal(_np(got), want)
    assert int(desc[0]) == cw
    for cut in (1000, 2199, 2200):
        hi_part, lo_part = np.ascontiguousarray(cts[:, cut:]), np.ascontiguousarray(cts[:, :cut])
        d_hi = torch.zeros(4, dtype=torch.int32, device="cuda")
        out_hi = ctx.aggregate(_dev(hi_part), fb.AGG_PACKED, carry_out=d_hi)
        out_lo = ctx.aggregate(_dev(lo_part), fb.AGG_PACKED)
        ctx.aggregate_carry_fixup(out_lo, int(d_hi[0]))
        assert np.array_equal(np.concatenate([_np(out_lo), _np(out_hi)]), want), cut
        out_lo2 = ctx.aggregate(_dev(lo_part), fb.AGG_PACKED, carry_in=int(d_hi[0]))
        assert np.array_equal(_np(out_lo2), want[:cut])
    # a shard made of propagate-only words: its carry out DEPENDS on the carry in (descriptor words 1-3)
    mid = np.ascontiguousarray(cts[:, 600:900])
    d_mid = torch.zeros(4, dtype=torch.int32, device="cuda")
    ctx.aggregate(_dev(mid), fb.AGG_PACKED, carry_out=d_mid)
    dm = [int(v) & 0xffffffff for v in d_mid.cpu()]
    assert dm[1] == 1 and dm[2] == 0 and dm[3] == 1          # carry_out = 0 + (carry_in >= 1)


def test_packed_aggregate_rejects_more_clients_than_digit_values(fb):
    ctx = ctx_for(fb, 8)
    cts = torch.zeros((300, 64), dtype=torch.int32, device="cuda").view(torch.uint32)
    with pytest.raises(fb._cabi.FlasheError):
        ctx.aggregate(cts, fb.AGG_PACKED)
    ctx.aggregate(cts, fb.AGG_ELEMENTWISE)


def test_sparse_apply_masks_rejects_bad_index_lists(fb):
    """Index lists come from other parties: out of range / unsorted / repeated entries must not reach the
    scatter (the reference raises IndexError at jzf_flashe.py:333)."""
    ctx = ctx_for(fb, 32)
    total = 1000
    dense = ctx.zeros_words(total)
    for bad in ([0, 5, 1000], [-1, 2, 3], [1, 3, 3, 4], [5, 4, 6]):
        idx = torch.tensor(bad, dtype=torch.int64, device="cuda")
        with pytest.raises(IndexError):
            ctx.sparse_apply_masks(0, [0], [1], fb.VectorSpan(len(bad), 8), idx, dense)
    # the kernel itself skips entries outside the dense vector
    idx = torch.tensor([3, 999, 1000, 5000], dtype=torch.int64, device="cuda")
    ctx.sparse_apply_masks(0, [0], [1], fb.VectorSpan(4, 8), idx, dense, validate=False)
    torch.cuda.synchronize()
    got = _np(dense)
    want = O.masks(KEY, 32, 8, 0, [0], [1], 4)
    assert got[3] == want[0] and got[999] == want[1] and np.count_nonzero(got) <= 2


def test_quant_edges_golden(fb, golden):
    c = golden.cases("quant_edges")[0]
    L = c["L"]
    ctx = ctx_for(fb, 32)
    q_ref = golden.words("qe_q", 32).reshape(3, L)
    x = _dev(golden["qe_x"])
    for k, (alpha, e) in enumerate(zip(c["alphas"], c["element_bits"])):
        q = ctx.encode(x, fb.CodecSpec(alpha=alpha, element_bits=e), fb.NoiseSpec(u=_dev(golden["qe_u"][k])), fb.VectorSpan(L, 1))
        assert np.array_equal(_np(q), q_ref[k]), k
    d = golden.cases("decode")[0]
    v = _dev(golden.words("qd_v", 32))
    out = ctx.decode(v, fb.CodecSpec(alpha=d["alpha"], element_bits=d["element_bits"], n_clients=d["n_clients"]), fb.VectorSpan(v.numel(), 1))
    assert np.array_equal(_np(out).view(np.uint64), golden["qd_out"].view(np.uint64))


# ---------------------------------------------------------------------------------------- seeded, vs oracle
@pytest.mark.parametrize("bits,n_jobs,L,n", [(20, 8, 300007, 3), (32, 1024, 400001, 4), (22, 16, 123457, 5)])
def test_full_path_vs_oracle(fb, bits, n_jobs, L, n):
    """encode+encrypt (layered alphas, device noise) -> aggregate (B and A) -> decrypt+decode."""
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    seg_end = [L // 5, L // 2, L]
    alphas = [0.59383450, 0.25, 1.0]
    codec = fb.CodecSpec(alpha=alphas, element_bits=16, n_clients=n, seg_end=seg_end)
    it = 11
    rs = np.random.RandomState(77)
    x = (rs.standard_normal((n, L)) * 0.2).astype(np.float32)
    x_d = _dev(x)
    cts = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x_d, codec, fb.NoiseSpec(seed=0xABCDEF0123, stream=100), span)
    # oracle consumes the very noise the device generated
    ct_want = []
    for k in range(n):
        u = _np(ctx.rng_uniform(0xABCDEF0123, 100 + k, 0, L))
        assert u.min() >= 0.0 and u.max() < 1.0
        q = np.empty(L, dtype=np.uint32)
        lo = 0
        for hi, a in zip(seg_end, alphas):
            q[lo:hi] = O.quantize(x[k, lo:hi], u[lo:hi], a, 16)
            lo = hi
        ct_want.append(O.encrypt(KEY, bits, n_jobs, it, k, "double", q))
    ct_want = np.stack(ct_want)
    assert np.array_equal(_np(cts), ct_want)
    for mode, name in ((fb.AGG_ELEMENTWISE, "elementwise"), (fb.AGG_PACKED, "packed")):
        agg = ctx.aggregate(cts, mode)
        agg_want = O.aggregate(bits, ct_want, name)
        assert np.array_equal(_np(agg), agg_want), name
    agg = ctx.aggregate(cts)
    p = ctx.empty_words(L)
    out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
    p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", O.aggregate(bits, ct_want))
    assert np.array_equal(_np(p), p_want)
    want = np.empty(L, dtype=np.float64)
    lo = 0
    for hi, a in zip(seg_end, alphas):
        want[lo:hi] = O.unquantize(p_want[lo:hi], a, 16, n)
        lo = hi
    assert np.array_equal(_np(out).view(np.uint64), want.view(np.uint64))
    # decoded sum is the sum of the clipped inputs up to quantisation error (sanity, not parity)
    clipped = sum(np.clip(x[k].astype(np.float64), -np.repeat(alphas, np.diff([0] + seg_end)), np.repeat(alphas, np.diff([0] + seg_end))) for k in range(n))
    step = 2 * np.repeat(alphas, np.diff([0] + seg_end)) / 65535
    assert np.all(np.abs(_np(out) - clipped) <= n * step * 1.01)


def test_encode_encrypt_batch_more_clients_than_one_launch_takes(fb):
    """200 clients in one call: the host mirror splits them over launches of at most 127 (FLASHE_MAX_STREAMS - 1)."""
    n, L, bits, n_jobs, it = 200, 4099, 32, 8, 1
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=0.5, element_bits=16, n_clients=n)
    x = (np.random.RandomState(5).standard_normal((n, L)) * 0.2).astype(np.float32)
    got = _np(ctx.encode_encrypt_batch(it, 3, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(seed=2, stream=10), span))
    for c in (0, 1, 126, 127, 128, 199):
        u = _np(ctx.rng_uniform(2, 10 + c, 0, L))
        assert np.array_equal(got[c], O.encrypt(KEY, bits, n_jobs, it, 3 + c, "double", O.quantize(x[c], u, 0.5, 16))), c


def test_cuda_graph_replay_of_a_round(fb):
    """DeviceContext.capture: the three kernels of a round recorded once, replayed on new inputs written into
    the same buffers — same bits as the call-by-call round and as the oracle."""
    L, n, bits, n_jobs, it = 100_003, 3, 20, 8, 1
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=0.4, element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=3, stream=0)
    x = torch.zeros((n, L), dtype=torch.float32, device="cuda")
    cts, agg = ctx.empty_words(L, rows=n), ctx.empty_words(L)
    out = torch.empty(L, dtype=torch.float64, device="cuda")

    def rnd():
        ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x, codec, noise, span, out=cts)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(it, [n], [0], agg, codec, span, out=out)

    replay = ctx.capture(rnd)
    xs = (np.random.RandomState(2).standard_normal((n, L)) * 0.2).astype(np.float32)
    x.copy_(_dev(xs))
    replay()
    torch.cuda.synchronize()
    got_ct, got_out = _np(cts).copy(), _np(out).copy()
    rnd()
    torch.cuda.synchronize()
    assert np.array_equal(got_ct, _np(cts)) and np.array_equal(got_out.view(np.uint64), _np(out).view(np.uint64))
    q = np.stack([O.quantize(xs[c], _np(ctx.rng_uniform(3, c, 0, L)), 0.4, 16) for c in range(n)])
    want = np.stack([O.encrypt(KEY, bits, n_jobs, it, c, "double", q[c]) for c in range(n)])
    assert np.array_equal(got_ct, want)
    p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", O.aggregate(bits, want))
    assert np.array_equal(got_out.view(np.uint64), O.unquantize(p_want, 0.4, 16, n).view(np.uint64))


@pytest.mark.parametrize("bits", [32, 20, 64, 120])
def test_sparse_apply_masks_batch_equals_per_client_calls(fb, bits):
    """flashe_sparse_apply_masks_batch (masks of every client into a workspace - one launch per run of equal list lengths -
    then one tiled accumulate) against one flashe_sparse_apply_masks per client and against the oracle's masks."""
    ctx = ctx_for(fb, bits)
    rs = np.random.RandomState(100 + bits)
    total, n_jobs, it = 150_001, 5, 3
    ks = [4000, 4000, 4000, 0, 777, 4000, 1]                               # runs of equal lengths, an empty list, odd lengths
    prf = [2, 3, 4, 5, 9, 10, 11]
    lists = [_dev(np.sort(rs.choice(total, size=k, replace=False)).astype(np.int64)) for k in ks]
    base = ctx.words_from_ints(np.array([int.from_bytes(rs.bytes(16), "little") & ((1 << bits) - 1) for _ in range(total)], dtype=object))
    for sign in (1, -1):
        want = base.clone()
        for c, ix in enumerate(lists):
            if ix.numel():
                ctx.sparse_apply_masks(it, [prf[c]], [sign], fb.VectorSpan(ix.numel(), n_jobs), ix, want)
        got = ctx.sparse_apply_masks_batch(it, prf, sign, n_jobs, lists, base.clone())
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int32), want.view(torch.int32)), sign
    # oracle: the term of client 4 (777 positions) by itself
    z = ctx.sparse_apply_masks_batch(it, [prf[4]], 1, n_jobs, [lists[4]], ctx.zeros_words(total))
    m = O.from_words(O.masks(KEY, bits, n_jobs, it, [prf[4]], [1], ks[4]), bits)
    zi = ctx.ints_from_words(z)
    idx = _np(lists[4])
    assert [int(zi[j]) for j in idx] == [int(v) for v in m]
    assert sum(int(v) for v in zi) == sum(int(v) for v in m)
    with pytest.raises(IndexError):
        ctx.sparse_apply_masks_batch(it, [1], 1, n_jobs, [_dev(np.array([5, 4], dtype=np.int64))], ctx.zeros_words(total))
    with pytest.raises(IndexError):                                        # unsorted list from a client: refused before the tiled sum
        ctx.sparse_sum([ctx.zeros_words(2)], [_dev(np.array([5, 4], dtype=np.int64))], total, [0], validate=True)
    with pytest.raises(IndexError):
        ctx.sparse_sum([ctx.zeros_words(1)], [_dev(np.array([total], dtype=np.int64))], total, [0], validate=True)


@pytest.mark.parametrize("n,total,kmax", [(70, 40_001, 900), (300, 20_003, 120), (32, 5_000, 5_000), (3, 100_000, 60_000)])
def test_sparse_tiled_paths_many_clients_dense_lists_unaligned(fb, n, total, kmax):
    """The tiled sparse kernels outside the shipped shape: more clients than travel in the kernel parameters (70: table in
    global memory), more than a block has threads (300: grouped walk, per-pair overlap kernels), index lists that cover
    the whole vector (32 x 5000: a tile's runs exceed the overlap kernel's shared-memory buffer; adjacent lists are equal),
    long runs per tile (3 x 60 %), and a dense vector that starts 4 bytes off a 16-byte boundary."""
    bits = 32
    ctx = ctx_for(fb, bits)
    rs = np.random.RandomState(n)
    ks = [kmax if kmax == total else int(rs.randint(1, kmax + 1)) for _ in range(n)]
    lists_np = [np.sort(rs.choice(total, size=k, replace=False)).astype(np.int64) for k in ks]
    lists = [_dev(a) for a in lists_np]
    vals_np = [rs.randint(0, 2 ** 32, size=k, dtype=np.uint64).astype(np.uint32) for k in ks]
    vals = [_dev(v) for v in vals_np]
    zeros = [int(z) for z in rs.randint(0, 2 ** 32, size=n, dtype=np.uint64)]
    want = np.full(total, sum(zeros) % 2 ** 32, dtype=np.uint64)
    for c in range(n):
        want[lists_np[c]] += vals_np[c].astype(np.uint64) + (2 ** 32 - zeros[c])
    want = (want % 2 ** 32).astype(np.uint32)
    buf = torch.zeros(total + 4, dtype=torch.int32, device="cuda").view(torch.uint32)
    for off in (0, 1):                                                     # 16-byte aligned / 4 bytes off
        out = buf[off:off + total]
        got = ctx.sparse_sum(vals, lists, total, zeros, out=out)
        assert np.array_equal(_np(got), want), off
    # every client's masks in one call == client by client
    base = _dev(rs.randint(0, 2 ** 32, size=total, dtype=np.uint64).astype(np.uint32))
    per_client = base.clone()
    prf = list(range(5, 5 + n))
    for c in range(n):
        ctx.sparse_apply_masks(7, [prf[c]], [-1], fb.VectorSpan(ks[c], 3), lists[c], per_client)
    one_call = ctx.sparse_apply_masks_batch(7, prf, -1, 3, lists, base.clone())
    assert torch.equal(one_call.view(torch.int32), per_client.view(torch.int32))
    ov = ctx.sparse_overlap(lists, total)
    assert ov == [int(np.intersect1d(lists_np[i], lists_np[i + 1], assume_unique=True).size) for i in range(n - 1)]


def test_dynamic_deal_concurrent_streams_and_graphs(fb):
    """The stream kernel's warps draw their work units from a ticket counter (one slot per stream, one per captured
    launch, reset by the launch's last warp).  Launches that overlap in time - two streams, two graphs replayed on two
    streams, the same context - must not disturb each other, and a slot must come back clean launch after launch:
    every result equals the one a lone launch on the default stream gave."""
    L, n, bits, n_jobs, it = 400_003, 4, 20, 16, 5
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=0.4, element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=3, stream=0)
    x = _dev((np.random.RandomState(8).standard_normal((n, L)) * 0.2).astype(np.float32))
    want_ct = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x, codec, noise, span)
    want_agg = ctx.aggregate(want_ct, fb.AGG_ELEMENTWISE)
    want_out = ctx.decrypt_decode(it, [n], [0], want_agg, codec, span)
    q = O.quantize(_np(x[1]), _np(ctx.rng_uniform(3, 1, 0, L)), 0.4, 16)
    assert np.array_equal(_np(want_ct[1]), O.encrypt(KEY, bits, n_jobs, it, 1, "double", q))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    bufs = [(ctx.empty_words(L, rows=n), ctx.empty_words(L), torch.empty(L, dtype=torch.float64, device="cuda")) for _ in streams]

    def rnd(k):
        cts, agg, out = bufs[k]
        ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x, codec, noise, span, out=cts)
        ctx.aggregate(cts, fb.AGG_ELEMENTWISE, out=agg)
        ctx.decrypt_decode(it, [n], [0], agg, codec, span, out=out)

    def check():
        torch.cuda.synchronize()
        for cts, agg, out in bufs:
            assert torch.equal(cts.view(torch.int32), want_ct.view(torch.int32))
            assert torch.equal(out.view(torch.int64), want_out.view(torch.int64))
            cts.zero_(); out.zero_()
        torch.cuda.synchronize()

    for _ in range(6):                                   # eager launches interleaved over three streams
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                rnd(k)
    check()
    replays = []
    for k, s in enumerate(streams):                      # one graph per stream, replayed concurrently
        with torch.cuda.stream(s):
            replays.append(ctx.capture(lambda k=k: rnd(k)))
    check()
    for _ in range(6):
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                replays[k]()
        rnd(0)                                           # ... with eager launches on the default stream in between
    check()


def test_decode_into_an_exported_peer_buffer(fb):
    """sharding.PeerGather on one rank: the decode kernel writes into memory allocated by flashe_peer_alloc (the
    buffer other GPUs would map and write into over NVLink; scripts/multi_gpu_check.py does that under torchrun)
    through a PeerSlice `out=`, shard by shard, and the owner's view equals the plain call bit for bit."""
    from flashe_b200.sharding import PeerGather, shard_bounds
    L, n, bits, n_jobs, it = 300_007, 4, 32, 8, 2
    ctx = ctx_for(fb, bits)
    codec = fb.CodecSpec(alpha=0.7, element_bits=16, n_clients=n)
    agg = _dev(np.random.RandomState(4).randint(0, 2 ** 32, L, dtype=np.uint64).astype(np.uint32))
    want = ctx.decrypt_decode(it, [n], [0], agg, codec, fb.VectorSpan(L, n_jobs))
    pg = PeerGather(ctx, L)
    for r in range(3):
        begin, count = shard_bounds(L, 3, r)
        span = fb.VectorSpan(L, n_jobs, begin=begin, count=count)
        ctx.decrypt_decode(it, [n], [0], agg[begin:begin + count].contiguous(), codec, span, out=pg.slice_of(begin, count))
    pg.finish()
    assert torch.equal(pg.tensor().view(torch.int64), want.view(torch.int64))
    with pytest.raises(ValueError):
        pg.slice_of(L - 3, 8)
    pg.close()


def test_single_scheme_many_streams(fb):
    """single masking decrypt subtracts one stream per survivor; > FLASHE_MAX_STREAMS needs chaining."""
    bits, n_jobs, L, n = 32, 8, 20000, 150
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    agg = _dev(np.random.RandomState(3).randint(0, 2 ** 32, L, dtype=np.uint64).astype(np.uint32))
    out = ctx.decrypt(3, [], list(range(128)), agg, span)
    out = ctx.decrypt(3, [], list(range(128, n)), out, span)
    want = O.decrypt(KEY, bits, n_jobs, 3, list(range(n)), "single", _np(agg))
    assert np.array_equal(_np(out), want)


def test_packed_aggregate_adversarial_and_shards(fb):
    """Carry chains: all-ones digits make every element pass its carry-in on (worst case for the
    look-ahead); element-range shards exchange carry descriptors."""
    for bits, dtype in ((20, np.uint32), (32, np.uint32), (40, np.uint64), (64, np.uint64)):
        ctx = ctx_for(fb, bits)
        top = (1 << bits) - 1
        rs = np.random.RandomState(bits)
        L, n = 5000, 7
        cts = rs.randint(0, 1 << min(bits, 62), size=(n, L), dtype=np.uint64).astype(dtype)
        cts[:, 1000:3500] = 0
        cts[0, 1000:3500] = top          # long run of digits that only propagate
        cts[1, 3499] = 1                 # ... and one carry injected at its end
        cts[:, 4000:4100] = top          # digits that generate and propagate
        want, cw = O.aggregate(bits, cts, "packed", return_carry=True)
        desc = torch.zeros(4, dtype=torch.int32, device="cuda")
        got = ctx.aggregate(_dev(cts), fb.AGG_PACKED, carry_out=desc)
        assert np.array_equal(_np(got), want), bits
        assert int(desc[0]) == cw
        # two shards, cut inside the propagate run
        cut = 2000
        hi_part = np.ascontiguousarray(cts[:, cut:])
        lo_part = np.ascontiguousarray(cts[:, :cut])
        d_hi = torch.zeros(4, dtype=torch.int32, device="cuda")
        out_hi = ctx.aggregate(_dev(hi_part), fb.AGG_PACKED, carry_out=d_hi)
        out_lo = ctx.aggregate(_dev(lo_part), fb.AGG_PACKED)          # carry_in unknown yet: 0
        c_hi = int(d_hi[0])                                            # last shard: carry_in really is 0
        ctx.aggregate_carry_fixup(out_lo, c_hi)
        assert np.array_equal(np.concatenate([_np(out_lo), _np(out_hi)]), want), bits
        # direct carry_in path gives the same
        out_lo2 = ctx.aggregate(_dev(lo_part), fb.AGG_PACKED, carry_in=c_hi)
        assert np.array_equal(_np(out_lo2), want[:cut])


def test_aggregate_unaligned_and_tail(fb):
    ctx = ctx_for(fb, 20)
    rs = np.random.RandomState(8)
    for L in (1, 3, 5, 1023, 4097):
        cts = rs.randint(0, 1 << 20, size=(9, L), dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(_np(ctx.aggregate(_dev(cts))), O.aggregate(20, cts))


# ---------------------------------------------------------------------------------------- the drop-in classes
def test_flashecipher_double_mask_many_dropout_runs(fb):
    """300 clients with every other one dropped: 150 runs = 300 PRF index terms, more than one launch
    takes (FLASHE_MAX_STREAMS); the drop-in chains the calls.  Also: a re-keyed cipher forgets what the
    old key prepared, and an iteration index that does not fit 4 bytes raises like the reference."""
    from flashe_b200.secureprotol import FlasheCipher
    b, nj, L, it = 32, 8, 777, 3
    survivors = list(range(0, 300, 2))
    agg = np.random.RandomState(3).randint(0, 1 << 32, size=L, dtype=np.uint64).astype(np.uint32)
    cipher = FlasheCipher(b, n_jobs=nj)
    cipher.generate_prp_seed(KEY)
    cipher.set_num_clients(300)
    cipher.set_iter_index(it)
    cipher.set_idx_list(raw_idx_list=survivors, mode="decrypt")
    assert len(cipher.index_prefix_for_add) + len(cipher.index_prefix_for_minus) == 300
    dec = cipher.decrypt(agg.astype(object))
    want = O.decrypt(KEY, b, nj, it, survivors, "double", agg)
    assert [int(v) for v in dec] == [int(v) for v in want]
    # re-keying drops the prepared buffers and the ring of the old key
    cipher.idx = 1
    cipher.set_num_params(L)
    cipher.prepare_encrypt()
    cipher.prepare_encrypt(rounds=3)
    assert cipher.next_iter_encrypt_prepared and cipher._ring is not None
    cipher.generate_prp_seed(bytes(range(1, 33)))
    assert cipher.next_iter_encrypt_prepared == {} and cipher._ring is None
    cipher.set_iter_index(it)
    ct = cipher.encrypt(agg.astype(object))
    want_ct = O.encrypt(bytes(range(1, 33)), b, nj, it, 1, "double", agg)
    assert [int(v) for v in ct] == [int(v) for v in want_ct]
    with pytest.raises(OverflowError):
        cipher.set_iter_index(1 << 32)
    with pytest.raises(OverflowError):
        cipher.set_iter_index(-1)


def test_prepare_encrypt_ring_refills_only_consumed_slots(fb):
    from flashe_b200.secureprotol import FlasheCipher
    from flashe_b200.device import launch_count
    b, nj, L = 20, 8, 5000
    cipher = FlasheCipher(b, n_jobs=nj)
    cipher.generate_prp_seed(KEY)
    cipher.idx = 2
    cipher.set_num_params(L)
    q = np.random.RandomState(4).randint(0, 65536, size=L).astype(np.uint32)
    cipher.prepare_encrypt(rounds=4)                     # iter_index = -1: rounds 0..3
    for it in range(6):
        cipher.set_iter_index(it)
        ct = cipher.encrypt(_dev(q))
        assert np.array_equal(_np(ct), O.encrypt(KEY, b, nj, it, 2, "double", q)), it
        l0 = launch_count()
        cipher.prepare_encrypt(rounds=4)                 # look-ahead: rounds it+1 .. it+4
        assert launch_count() - l0 == 1, "one consumed slot -> one round regenerated"


def test_flashecipher_dropin_matches_reference_outputs(fb, golden):
    """Driven exactly like the reference's notebook / aggregator drive jzf_flashe.FlasheCipher."""
    from flashe_b200.secureprotol import FlasheCipher
    c = [k for k in golden.cases("roundtrip") if k["name"] == "rt_b20_n3"][0]
    name, b, nj, L, n, it = (c[k] for k in ("name", "int_bits", "n_jobs", "L", "n_clients", "iter"))
    q_ref = golden.words(name + "_q", 32).reshape(n, L)
    ct_ref = golden.words(name + "_ct", b).reshape(n, L)
    cts = []
    for k in range(n):
        cipher = FlasheCipher(b, n_jobs=nj)
        assert cipher.encrypt(np.zeros(3, dtype=object)) is None          # no key yet
        cipher.generate_prp_seed(KEY)
        cipher.idx = k
        cipher.set_iter_index(it)
        assert cipher.encrypt([1, 2, 3]) is None                           # not an ndarray
        ct = cipher.encrypt(q_ref[k].astype(object))
        assert ct.dtype == object and isinstance(ct[0], int)
        assert [int(v) for v in ct] == [int(v) for v in ct_ref[k]]
        assert cipher.get_idx_list() == [k]
        cts.append(ct)
    from flashe_b200 import aggregate as agg_mod
    total_b = agg_mod.aggregate(cts, b, is_compressed=False)
    total_a = agg_mod.aggregate(cts, b, is_compressed=True)
    assert [int(v) for v in total_b] == [int(v) for v in golden.words(name + "_aggB", b)]
    assert [int(v) for v in total_a] == [int(v) for v in golden.words(name + "_aggA", b)]
    cipher = FlasheCipher(b, n_jobs=nj)
    cipher.generate_prp_seed(KEY)
    cipher.set_num_clients(n)
    cipher.set_iter_index(it)
    cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
    dec = cipher.decrypt(total_b)
    assert [int(v) for v in dec] == [int(v) for v in golden.words(name + "_decB", b)]
    from flashe_b200.secureprotol.quantize import _static_unquantize_padding_asymmetric
    out = _static_unquantize_padding_asymmetric(dec, c["alpha"], 16, n)
    assert out.dtype == object and isinstance(out[0], float)
    assert np.array_equal(out.astype(np.float64).view(np.uint64), golden[name + "_decoded"].view(np.uint64))


def test_flashecipher_precompute_and_dropout(fb, golden):
    from flashe_b200.secureprotol import FlasheCipher
    cases = golden.cases("dropout")
    c0 = cases[0]
    b, nj, L, n, it = (c0[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter"))
    q = golden.words("drop_q", 32).reshape(n, L)
    ct_ref = golden.words("drop_ct", b).reshape(n, L)
    # prepare_encrypt at iter-1 then encrypt at iter == on-the-fly (reference probe P9)
    cipher = FlasheCipher(b, n_jobs=nj)
    cipher.generate_prp_seed(KEY); cipher.idx = 2; cipher.set_num_clients(n); cipher.set_num_params(L)
    cipher.set_iter_index(it - 1)
    cipher.prepare_encrypt()
    assert 'add' in cipher.next_iter_encrypt_prepared
    cipher.set_iter_index(it)
    ct = cipher.encrypt(q[2].astype(object))
    assert [int(v) for v in ct] == [int(v) for v in ct_ref[2]]
    assert 'add' not in cipher.next_iter_encrypt_prepared          # consumed, as in the reference
    # decrypt with and without prepare_decrypt for every survivor set; the reference is wrong for
    # [1,2,3] and [0,2] WITH precompute — here precomputed == on-the-fly == sum of plaintexts
    for c in cases:
        for pre in (False, True):
            cipher = FlasheCipher(b, n_jobs=nj)
            cipher.generate_prp_seed(KEY); cipher.set_num_clients(n); cipher.set_num_params(L)
            cipher.set_iter_index(it)
            if pre:
                cipher.prepare_decrypt()
            cipher.set_idx_list(raw_idx_list=list(c["survivors"]), mode="decrypt")
            dec = cipher.decrypt(golden.words(c["name"] + "_agg", b).astype(object))
            assert [int(v) for v in dec] == [int(v) for v in golden.words(c["name"] + "_dec", b)], (c["survivors"], pre)
            assert cipher.next_iter_decrypt_prepared == {} and cipher.next_iter_decrypt_prepared_idx == {}


def test_flashecipher_sparse_single(fb, golden):
    from flashe_b200 import aggregate as agg_mod
    from flashe_b200.secureprotol import FlasheCipher
    c = golden.cases("sparse")[0]
    b, nj, n, it, total = (c[k] for k in ("int_bits", "n_jobs", "n_clients", "iter", "total"))
    masks = [[int(v) for v in golden["sparse_mask_%d" % k]] for k in range(n)]
    uploads = []
    for k in range(n):
        cipher = FlasheCipher(b, mask="single", n_jobs=nj)
        cipher.generate_prp_seed(KEY); cipher.idx = k; cipher.set_iter_index(it)
        ct = cipher.encrypt(golden.words("sparse_q_%d" % k, 32).astype(object))
        uploads.append(np.append(ct, [int(golden["sparse_zero"][k])]))      # jzf_aggregator.py:741-743
    dense = agg_mod.expand_to_dense(uploads, masks, total, b)
    agg = agg_mod.aggregate(dense, b)
    assert [int(v) for v in agg] == [int(v) for v in golden.words("sparse_agg", b)]
    fused = agg_mod.aggregate_sparse(uploads, masks, total, b)              # expand + reduce without the dense vectors
    assert [int(v) for v in fused] == [int(v) for v in golden.words("sparse_agg", b)]
    cipher = FlasheCipher(b, mask="single", n_jobs=nj)
    cipher.generate_prp_seed(KEY); cipher.set_iter_index(it)
    cipher.masks = masks; cipher.total = total
    cipher.set_idx_list(raw_idx_list=list(range(n)), mode="decrypt")
    dec = cipher.decrypt(agg)
    assert [int(v) for v in dec] == [int(v) for v in golden.words("sparse_dec", b)]
    d = agg_mod.dynamic_masking(masks, total)
    want_choice, sc, dc = O.dynamic_masking(masks)
    assert (d["choice"], d["single_cost"], d["double_cost"]) == (want_choice, sc, dc)
    dbl = FlasheCipher(b, mask="double", n_jobs=nj)
    dbl.generate_prp_seed(KEY); dbl.set_iter_index(it); dbl.masks = masks; dbl.total = total
    with pytest.raises(NotImplementedError):
        dbl.set_idx_list(raw_idx_list=[0, 1, 2], mode="decrypt")


class _W(object):
    """Minimal stand-in for JZFOrderDictWeights (framework/jzf_weights.py:431-477): `_weights` dict +
    `walking_order` = keys sorted as strings."""

    def __init__(self, d):
        self._weights = d
        self.walking_order = sorted(d.keys(), key=str)


def test_quantizing_client_seed_level_parity(fb, golden):
    """np.random.seed(s) + QuantizingClient.quantize == the reference's outputs for the same seed."""
    from flashe_b200.secureprotol import QuantizingClient
    c = golden.cases("batch")[0]
    b, L, n, e = c["int_bits"], c["L"], c["n_clients"], c["element_bits"]
    q_ref = golden.words("batch_q", 32).reshape(n, L)
    for batch in (False, True):
        qc = QuantizingClient(b if batch else 20, None, None, batch, e, True, True)
        qc.num_clients = n
        w = _W({"layer0": golden["batch_x"][0].copy()})
        qc.set_layer_size_list(w)
        qc.past_layer_std_list = [0.05]           # alpha = 5.938345 * 0.05, the fixture's alpha
        np.random.seed(6100)
        qc.quantize(w)
        if not batch:
            assert [int(v) for v in w._weights["layer0"]] == [int(v) for v in q_ref[0]]
        else:
            w_ref = golden.words("batch_w", b).reshape(n, c["words"], 2)[0]
            assert [int(v) for v in w._weights["layer0"]] == [int(lo) | (int(hi) << 64) for lo, hi in w_ref]
        # unquantize round trip of the golden aggregate
        if batch:
            w2 = _W({"layer0": np.array([int(lo) | (int(hi) << 64) for lo, hi in golden.words("batch_dec", b)], dtype=object)})
        else:
            w2 = _W({"layer0": golden.words("batch_unb", 32).astype(object)})
        qc.shape_list = [(L,)]
        qc.unquantize(w2)
        assert np.array_equal(np.asarray(w2._weights["layer0"], dtype=np.float64).view(np.uint64), golden["batch_decoded"].view(np.uint64))


@pytest.mark.parametrize("bits,n_jobs,L", [(32, 8, 1 << 20), (32, 16, 3_000_000), (28, 4, 600_000),
                                           (32, 7, 7 * 150_001), (32, 24, 2_000_003), (27, 3, 3 * 100_002)])
def test_aligned_fast_path_vs_oracle(fb, bits, n_jobs, L):
    """m = 4 takes the lane-local path: 128-bit accesses when the chunk starts on a 16-byte boundary of
    the buffers, 64/32-bit pieces otherwise (odd chunk lengths put chunk starts at every residue mod 4:
    the last three cases); every mode must agree with the oracle, for whole vectors and for shards cut
    on and off 4-element boundaries."""
    ctx = ctx_for(fb, bits)
    it, n = 9, 3
    rs = np.random.RandomState(123)
    x = (rs.standard_normal((n, L)) * 0.1).astype(np.float32)
    u = rs.random_sample((n, L))
    alpha = 0.59383450
    codec = fb.CodecSpec(alpha=alpha, element_bits=16, n_clients=n)
    q = np.stack([O.quantize(x[k], u[k], alpha, 16) for k in range(n)])
    ct_want = np.stack([O.encrypt(KEY, bits, n_jobs, it, k, "double", q[k]) for k in range(n)])
    full = fb.VectorSpan(L, n_jobs)
    # masks / apply / encode(+q_out, given noise) / batch / share
    assert np.array_equal(_np(ctx.masks(it, [1, 2], [1, -1], full)), O.masks(KEY, bits, n_jobs, it, [1, 2], [1, -1], L))
    assert np.array_equal(_np(ctx.encrypt(it, 1, fb.SCHEME_DOUBLE, _dev(q[1]), full)), ct_want[1])
    q_out = torch.empty(L, dtype=torch.uint32, device="cuda")
    got = ctx.encode_encrypt(it, 2, fb.SCHEME_DOUBLE, _dev(x[2]), codec, fb.NoiseSpec(u=_dev(u[2])), full, q_out=q_out)
    assert np.array_equal(_np(got), ct_want[2]) and np.array_equal(_np(q_out), q[2])
    for share in (False, True):
        got = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(u=_dev(u).reshape(-1)), full, share_streams=share)
        assert np.array_equal(_np(got), ct_want), share
    # device noise on the fast path equals rng_uniform
    got = ctx.encode_encrypt(it, 0, fb.SCHEME_DOUBLE, _dev(x[0]), codec, fb.NoiseSpec(seed=99, stream=5), full)
    ud = _np(ctx.rng_uniform(99, 5, 0, L))
    assert np.array_equal(_np(got), O.encrypt(KEY, bits, n_jobs, it, 0, "double", O.quantize(x[0], ud, alpha, 16)))
    # decrypt + decode
    agg = O.aggregate(bits, ct_want)
    p = ctx.empty_words(L)
    out = ctx.decrypt_decode(it, [n], [0], _dev(agg), codec, full, p_out=p)
    p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", agg)
    assert np.array_equal(_np(p), p_want)
    assert np.array_equal(_np(out).view(np.uint64), O.unquantize(p_want, alpha, 16, n).view(np.uint64))
    # shards: aligned cut, unaligned cut (falls back to the slab path at the edges), odd begin
    for a, e in ((0, L // 2), (L // 2, L), (L // 4 + 1, L // 2 + 3), (4 * 1001, 4 * 1001 + 70001)):
        sp = fb.VectorSpan(L, n_jobs, a, e - a)
        got = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(np.ascontiguousarray(x[:, a:e])), codec,
                                       fb.NoiseSpec(u=_dev(np.ascontiguousarray(u[:, a:e])).reshape(-1)), sp)
        assert np.array_equal(_np(got), ct_want[:, a:e]), (a, e)
        out = ctx.decrypt_decode(it, [n], [0], _dev(np.ascontiguousarray(agg[a:e])), codec, sp)
        assert np.array_equal(_np(out).view(np.uint64), O.unquantize(p_want[a:e], alpha, 16, n).view(np.uint64)), (a, e)


def test_device_noise_is_the_documented_philox_stream(fb):
    """flashe_rng_uniform == oracle restatement (Philox4x32-10, counter (j>>1, stream), res53), incl.
    odd begins, 64-bit seeds/streams and counters beyond 2^32."""
    ctx = ctx_for(fb, 32)
    for seed, stream, begin, cnt in [(0, 0, 0, 4097), (0x0123456789ABCDEF, 7, 5, 100001), (0xFFFFFFFFFFFFFFFF, 1 << 40, (1 << 33) + 3, 5000),
                                     (42, 63, 99_999_999, 2)]:
        got = _np(ctx.rng_uniform(seed, stream, begin, cnt))
        want = O.noise_uniform(seed, stream, begin, cnt)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (seed, stream, begin)
        got32 = _np(ctx.rng_uniform(seed, stream, begin, cnt, resolution=32))
        want32 = O.noise_uniform(seed, stream, begin, cnt, resolution=32)
        assert np.array_equal(got32.view(np.uint64), want32.view(np.uint64)), (seed, stream, begin)
        assert got32.min() >= 0.0 and got32.max() < 1.0


@pytest.mark.parametrize("bits,n_jobs,L", [(32, 16, 1_000_000), (32, 7, 7 * 50_001), (20, 8, 600_001), (24, 5, 300_000), (64, 8, 200_000), (16, 4, 100_000)])
def test_throughput_mode_noise_32_bit_resolution(fb, bits, n_jobs, L):
    """NoiseSpec(resolution=32): one Philox word per element.  Every fused path (aligned / shifted / slab, batch and
    shared streams, the premasked online step) must draw exactly the documented stream: ciphertexts equal the
    oracle's when it is fed rng_uniform(..., resolution=32)."""
    n, it, alpha = 3, 2, 0.37
    ctx = ctx_for(fb, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=alpha, element_bits=16, n_clients=n)
    x = (np.random.RandomState(L % 97).standard_normal((n, L)) * 0.2).astype(np.float32)
    noise = fb.NoiseSpec(seed=11, stream=5, resolution=32)
    want = []
    for c in range(n):
        u = O.noise_uniform(11, 5 + c, 0, L, resolution=32)
        q = O.quantize(x[c], u, alpha, 16)
        want.append(O.encrypt(KEY, bits, n_jobs, it, c, "double", O.to_words(q, bits) if bits > 32 else q))
    want = np.stack(want)
    for share in (False, True):
        got = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, noise, span, share_streams=share)
        assert np.array_equal(_np(got), want), share
    one = ctx.encode_encrypt(it, 1, fb.SCHEME_DOUBLE, _dev(x[1]), codec, fb.NoiseSpec(seed=11, stream=6, resolution=32), span)
    assert np.array_equal(_np(one), want[1])
    # a shard that starts on an odd element (pairs / quads of the noise stream are cut)
    lo, cnt = 12345, L // 3
    sp = fb.VectorSpan(L, n_jobs, lo, cnt)
    part = ctx.encode_encrypt(it, 2, fb.SCHEME_DOUBLE, _dev(np.ascontiguousarray(x[2, lo:lo + cnt])), codec,
                              fb.NoiseSpec(seed=11, stream=7, resolution=32), sp)
    assert np.array_equal(_np(part), want[2, lo:lo + cnt])
    if bits <= 32:
        masks = torch.stack([ctx.masks(it, [c, c + 1], [1, -1], span).view(torch.int32) for c in range(n)]).view(torch.uint32)
        online = ctx.encode_add_premasked_batch(_dev(x), codec, noise, masks, span)
        assert np.array_equal(_np(online), want)


def test_encode_division_and_floor_tricks_dense_sweep(fb):
    """The encode kernel divides through a host-computed reciprocal + FMA corrections and floors through
    a round-down add; sweep a dense grid of float32 inputs (every 997th bit pattern across the clip
    range, plus the clip edges, zeros and denormals) for many alphas and noise values against the
    oracle's plain C arithmetic."""
    ctx = ctx_for(fb, 32)
    rs = np.random.RandomState(123)
    for alpha in [0.5938345, 1.0, 0.1, 3.3e-4, 7.77e3, 1e-20, 2.5e15, float(np.float32(1.0) / 3)]:
        a32 = np.float32(alpha)
        hi = np.float32(a32 * np.float32(1.5)).view(np.uint32)
        pos = np.arange(0, int(hi), 997, dtype=np.uint32)
        bits = np.concatenate([pos, pos | np.uint32(0x80000000), np.array([0, 0x80000000, 1, 0x807FFFFF, 0x7F800000, 0xFF800000], dtype=np.uint32)])
        x = bits.view(np.float32)
        L = x.size
        for mode in ("zero", "almost_one", "random"):
            u = {"zero": np.zeros(L), "almost_one": np.full(L, 1.0 - 2.0 ** -53), "random": rs.random_sample(L)}[mode]
            span = fb.VectorSpan(L, 8)
            codec = fb.CodecSpec(alpha=float(alpha), element_bits=16)
            got = _np(ctx.encode(_dev(x), codec, fb.NoiseSpec(u=_dev(u)), span))
            want = O.quantize(x, u, float(alpha), 16)
            assert np.array_equal(got, want), (alpha, mode)
            # the fused kernel (quad path) agrees too
            ct = _np(ctx.encode_encrypt(3, 1, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(u=_dev(u)), span))
            assert np.array_equal(ct, O.encrypt(KEY, 32, 8, 3, 1, "double", want)), (alpha, mode)


def test_decode_division_sweep(fb):
    """decode divides through RN(1/den) + two FMA corrections (library division when a layer's scale is
    out of the guarded range): sweep element widths, client counts and alphas (incl. tiny / huge ones that
    take the fallback) against the oracle, float64 bit patterns, standalone and fused with decrypt."""
    rs = np.random.RandomState(77)
    L = 40_003
    for bits, wdt in ((32, np.uint32), (64, np.uint64)):
        ctx = ctx_for(fb, bits)
        span = fb.VectorSpan(L, 8)
        for e, n, alpha in ((16, 3, 0.5938345), (16, 64, 0.0123), (8, 1, 7.25), (20, 10, 1e-3), (24, 1000, 3.0e5), (12, 7, 1e-200),
                            (16, 5, 1e180), (1, 2, 0.1), (16, 4096, 2.0 ** -30), (23, 33, 1.0 / 3.0)):
            top = min((1 << e) * n, (1 << min(bits, 63)) - 1)
            v = rs.randint(0, top + 1, L, dtype=np.int64).astype(wdt)
            v[:4] = (0, 1, top, top - 1)
            codec = fb.CodecSpec(alpha=alpha, element_bits=e, n_clients=n)
            want = O.unquantize(v, alpha, e, n)
            got = _np(ctx.decode(_dev(v), codec, span))
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (bits, e, n, alpha)
            off = _np(ctx.decode(_dev(v)[1:], codec, fb.VectorSpan(L, 8, 1, L - 1)))   # not 16-byte aligned: scalar kernel
            assert np.array_equal(off.view(np.uint64), want[1:].view(np.uint64)), (bits, e, n, alpha)
    # fused with decrypt, lane-local and slab paths
    for bits, n_jobs in ((32, 8), (20, 8)):
        ctx = ctx_for(fb, bits)
        span = fb.VectorSpan(L, n_jobs)
        agg = rs.randint(0, 1 << 20, L, dtype=np.int64).astype(np.uint32)
        for e, n, alpha in ((16, 3, 0.5938345), (12, 7, 1e-200), (16, 12, 41.5)):
            codec = fb.CodecSpec(alpha=alpha, element_bits=e, n_clients=n)
            p_want = O.decrypt(KEY, bits, n_jobs, 4, list(range(n)), "double", agg)
            got = _np(ctx.decrypt_decode(4, [n], [0], _dev(agg), codec, span))
            assert np.array_equal(got.view(np.uint64), O.unquantize(p_want, alpha, e, n).view(np.uint64)), (bits, e, n, alpha)


@pytest.mark.parametrize("bits", [32, 20, 64, 120])
def test_sparse_sum_equals_expand_then_aggregate(fb, bits):
    """flashe_sparse_sum (fill with the sum of the zero words + one scatter-add per client) against the
    reference's own order of operations: expand_to_dense per client, then the element-wise reduce."""
    ctx = ctx_for(fb, bits)
    rs = np.random.RandomState(bits)
    total, n = 200_003, 7
    compacts, indexes, zeros, dense = [], [], [], []
    for c in range(n):
        k = int(rs.randint(0, 5000)) if c else 0                           # one client with an empty upload
        idx = np.sort(rs.choice(total, size=k, replace=False)).astype(np.int64)
        vals = [int.from_bytes(rs.bytes(16), "little") & ((1 << bits) - 1) for _ in range(k)]
        zero = int.from_bytes(rs.bytes(16), "little") & ((1 << bits) - 1)
        w = ctx.words_from_ints(np.array(vals, dtype=object)) if k else ctx.empty_words(0)
        ix = _dev(idx)
        compacts.append(w); indexes.append(ix); zeros.append(zero)
        dense.append(ctx.sparse_expand(w, ix, total, zero))
    signed = {torch.uint32: torch.int32, torch.uint64: torch.int64}
    want = dense[0]
    for d in dense[1:]:
        want = ctx.aggregate(torch.stack([want.view(signed[want.dtype]), d.view(signed[d.dtype])]).view(want.dtype))
    got = ctx.sparse_sum(compacts, indexes, total, zeros)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))
    # python big-int cross-check on a few positions
    gi = ctx.ints_from_words(got)
    for j in (0, int(indexes[1][0]) if indexes[1].numel() else 1, total - 1):
        s_ = 0
        for c in range(n):
            pos = np.searchsorted(_np(indexes[c]), j)
            hit = pos < indexes[c].numel() and int(indexes[c][pos]) == j
            s_ += int(ctx.ints_from_words(compacts[c][pos:pos + 1])[0]) if hit else zeros[c]
        assert int(gi[j]) == s_ % (1 << bits), j


@pytest.mark.parametrize("bits", [32, 20, 24, 64, 120])
def test_item_geometry_every_counter_offset(fb, bits):
    """Items are cut on multiples of 64 of the AES counter (chunk begin + block): with an odd chunk length
    and many chunks, chunk begins take every value mod 64 and mod 256, items straddle counter windows at
    every position, and chunks end with every partial-item length.  Masks, encrypt and shards against
    the oracle."""
    ctx = ctx_for(fb, bits)
    n_jobs = 131
    L = n_jobs * 5003 + 17                       # chunks of 5004 / 5003 elements
    it = 11
    want = O.masks(KEY, bits, n_jobs, it, [5, 6], [1, -1], L)
    full = fb.VectorSpan(L, n_jobs)
    assert np.array_equal(_np(ctx.masks(it, [5, 6], [1, -1], full)), want)
    for a, e in ((4 * 777, 4 * 777 + 100_000), (12345, 12345 + 77_777), (L - 50_001, L)):
        got = _np(ctx.masks(it, [5, 6], [1, -1], fb.VectorSpan(L, n_jobs, a, e - a)))
        assert np.array_equal(got, want[a:e]), (a, e)
    rs = np.random.RandomState(bits)
    if bits <= 32:
        q = rs.randint(0, 65536, L).astype(np.uint32)
    elif bits <= 64:
        q = rs.randint(0, 65536, L).astype(np.uint64)
    else:
        q = np.stack([rs.randint(0, 2 ** 62, L).astype(np.uint64), rs.randint(0, 2 ** 50, L).astype(np.uint64)], axis=1)
    assert np.array_equal(_np(ctx.encrypt(it, 5, fb.SCHEME_DOUBLE, _dev(q), full)), O.encrypt(KEY, bits, n_jobs, it, 5, "double", q))
