"""The CPU oracle against the committed outputs of the reference's own Python code
(tests/golden/flashe_golden.npz, made by tests/golden/make_golden.py) and FIPS-197 C.3."""
import numpy as np
import pytest

from oracle import oracle as O


def test_aes256_fips197_c3():
    key = bytes(range(32))
    pt = bytes.fromhex("00112233445566778899aabbccddeeff")
    assert O.aes256_encrypt_block(key, pt).hex() == "8ea2b7ca516745bfeafc49904b496089"


def test_key_reduction_matches_reference_rule():
    # jzf_flashe.py:285-292 left-pads the seed to 256 BYTES; jzf_aes.py:21-28 keeps the low 32.
    key = bytes(range(32))
    pt = bytes(16)
    assert O.aes256_encrypt_block(bytes(224) + key, pt) == O.aes256_encrypt_block(key, pt)
    assert O.aes256_encrypt_block(b"\x01", pt) == O.aes256_encrypt_block(bytes(31) + b"\x01", pt)


def test_chunk_bounds_cover():
    for L in (0, 1, 3, 8, 1000, 1003):
        for n in (1, 3, 8, 16):
            prev = 0
            for i in range(n):
                b, e = O.chunk_bounds(L, n, i)
                assert b == prev and e >= b
                prev = e
            assert prev == L


def test_masks(golden):
    for c in golden.cases("masks"):
        b = c["int_bits"]
        got = O.masks(golden.key, b, c["n_jobs"], c["iter"], [c["prf_idx"]], [1], c["L"])
        assert np.array_equal(got, golden.words(c["name"], b)), c
        # a sub-range must agree with the full vector (shards)
        if c["L"] > 10:
            j0, cnt = c["L"] // 3, c["L"] // 2
            sub = O.masks(golden.key, b, c["n_jobs"], c["iter"], [c["prf_idx"]], [1], c["L"], j0, cnt)
            assert np.array_equal(sub, got[j0:j0 + cnt])


def test_threads_do_not_change_results(golden):
    c = golden.cases("masks")[6]
    O.set_threads(5)
    try:
        got = O.masks(golden.key, c["int_bits"], c["n_jobs"], c["iter"], [c["prf_idx"]], [1], c["L"])
    finally:
        O.set_threads(1)
    assert np.array_equal(got, golden.words(c["name"], c["int_bits"]))


def test_roundtrip(golden):
    for c in golden.cases("roundtrip"):
        name, b, nj, L, n, it, scheme = (c[k] for k in ("name", "int_bits", "n_jobs", "L", "n_clients", "iter", "scheme"))
        xs, us = golden[name + "_x"], golden[name + "_u"]
        q_ref = golden.words(name + "_q", 32).reshape(n, L)
        ct_ref = golden.words(name + "_ct", b).reshape(n, L)
        cts = []
        for k in range(n):
            q = O.quantize(xs[k], us[k], c["alpha"], c["element_bits"])
            assert np.array_equal(q, q_ref[k]), (name, k)
            ct = O.encrypt(golden.key, b, nj, it, k, scheme, q.astype(ct_ref.dtype))
            assert np.array_equal(ct, ct_ref[k]), (name, k)
            cts.append(ct)
        cts = np.stack(cts)
        agg_b = O.aggregate(b, cts, "elementwise")
        agg_a = O.aggregate(b, cts, "packed")
        assert np.array_equal(agg_b, golden.words(name + "_aggB", b))
        assert np.array_equal(agg_a, golden.words(name + "_aggA", b))
        dec_b = O.decrypt(golden.key, b, nj, it, list(range(n)), scheme, agg_b)
        dec_a = O.decrypt(golden.key, b, nj, it, list(range(n)), scheme, agg_a)
        assert np.array_equal(dec_b, golden.words(name + "_decB", b))
        assert np.array_equal(dec_a, golden.words(name + "_decA", b))
        out = O.unquantize(dec_b, c["alpha"], c["element_bits"], n)
        assert np.array_equal(out.view(np.uint64), golden[name + "_decoded"].view(np.uint64))


def test_packed_aggregate_differs_from_elementwise(golden):
    c = golden.cases("roundtrip")[0]
    a = golden.words(c["name"] + "_aggA", c["int_bits"])
    b = golden.words(c["name"] + "_aggB", c["int_bits"])
    assert (a != b).sum() > len(a) // 2     # the carry leak of SURVEY §0.4 is real and reproduced


def test_dropout(golden):
    cases = golden.cases("dropout")
    c0 = cases[0]
    b, nj, L, n, it = (c0[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter"))
    ct = golden.words("drop_ct", b).reshape(n, L)
    q = golden.words("drop_q", 32).reshape(n, L)
    for c in cases:
        add, minus = O.collapse_runs(c["survivors"])
        assert add == c["add"] and minus == c["minus"], c
        agg = O.aggregate(b, ct[sorted(c["survivors"])], "elementwise")
        assert np.array_equal(agg, golden.words(c["name"] + "_agg", b))
        dec = O.decrypt(golden.key, b, nj, it, c["survivors"], "double", agg)
        assert np.array_equal(dec, golden.words(c["name"] + "_dec", b))
        assert np.array_equal(dec, q[sorted(c["survivors"])].sum(axis=0).astype(np.uint32))


def test_runs(golden):
    for row in golden.cases("runs")[0]["rows"]:
        add, minus = O.collapse_runs(row["survivors"])
        assert add == row["add"] and minus == row["minus"], row


def test_precompute_buffers(golden):
    c = golden.cases("precompute")[0]
    b, nj, L, n, it, idx = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter", "idx"))
    add = O.masks(golden.key, b, nj, it, [idx], [1], L)
    minus = O.masks(golden.key, b, nj, it, [idx + 1], [1], L)
    assert np.array_equal(add, golden.words("pre_enc_add", b))
    assert np.array_equal(minus, golden.words("pre_enc_minus", b))
    q = golden.words("pre_q", 32)
    assert np.array_equal((q + add - minus) & ((1 << b) - 1), golden.words("pre_ct", b))
    assert np.array_equal(O.encrypt(golden.key, b, nj, it, idx, "double", q), golden.words("pre_ct", b))
    assert np.array_equal(O.masks(golden.key, b, nj, it, [n], [1], L), golden.words("pre_dec_add", b))
    assert np.array_equal(O.masks(golden.key, b, nj, it, [0], [1], L), golden.words("pre_dec_minus", b))


def test_sparse_single(golden):
    c = golden.cases("sparse")[0]
    b, nj, n, it, total = (c[k] for k in ("int_bits", "n_jobs", "n_clients", "iter", "total"))
    idx = [golden["sparse_mask_%d" % k] for k in range(n)]
    dense = []
    for k in range(n):
        q = golden.words("sparse_q_%d" % k, 32)
        ct = O.encrypt(golden.key, b, nj, it, k, "single", q)
        assert np.array_equal(ct, golden.words("sparse_ct_%d" % k, b))
        dense.append(O.expand_to_dense(b, ct, idx[k], total, int(golden["sparse_zero"][k])))
    agg = O.aggregate(b, np.stack(dense), "elementwise")
    assert np.array_equal(agg, golden.words("sparse_agg", b))
    dec = O.sparse_single_decrypt(golden.key, b, nj, it, idx, total, agg)
    assert np.array_equal(dec, golden.words("sparse_dec", b))


def test_batch120(golden):
    c = golden.cases("batch")[0]
    b, nj, L, n, it, e, f = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter", "element_bits", "factor"))
    nw = c["words"]
    q_ref = golden.words("batch_q", 32).reshape(n, L)
    w_ref = golden.words("batch_w", b).reshape(n, nw, 2)
    ct_ref = golden.words("batch_ct", b).reshape(n, nw, 2)
    cts = []
    for k in range(n):
        q = O.quantize(golden["batch_x"][k], golden["batch_u"][k], c["alpha"], e)
        assert np.array_equal(q, q_ref[k])
        w = O.batch(q, b, e, f)
        assert np.array_equal(w, w_ref[k])
        ct = O.encrypt(golden.key, b, nj, it, k, "double", w)
        assert np.array_equal(ct, ct_ref[k])
        cts.append(ct)
    agg = O.aggregate(b, np.stack(cts), "elementwise")
    assert np.array_equal(agg, golden.words("batch_agg", b))
    dec = O.decrypt(golden.key, b, nj, it, list(range(n)), "double", agg)
    assert np.array_equal(dec, golden.words("batch_dec", b))
    unb = O.unbatch(dec, b, e, f)[:L]
    assert np.array_equal(unb, golden.words("batch_unb", 32))
    out = O.unquantize(unb, c["alpha"], e, n)
    assert np.array_equal(out.view(np.uint64), golden["batch_decoded"].view(np.uint64))
    # the shipped dense path sums the packed uploads (jzf_aggregator.py:404-419): 120-bit digits
    agg_a = O.aggregate(b, np.stack(cts), "packed")
    assert np.array_equal(agg_a, golden.words("batch_aggA", b))
    assert not np.array_equal(agg_a, agg)
    dec_a = O.decrypt(golden.key, b, nj, it, list(range(n)), "double", agg_a)
    assert np.array_equal(dec_a, golden.words("batch_decA", b))


def test_packed_aggregate_wide_words(golden):
    """int_bits 65..128: big-int sums of the packed wire integers with long carry ripples."""
    cases = golden.cases("packed_wide")
    assert len(cases) >= 6
    for c in cases:
        b, n, L = c["int_bits"], c["n_clients"], c["L"]
        cts = golden.words(c["name"] + "_ct", b).reshape(n, L, 2)
        assert np.array_equal(O.aggregate(b, cts, "packed"), golden.words(c["name"] + "_aggA", b)), c["name"]


def test_quant_edges_and_decode(golden):
    c = golden.cases("quant_edges")[0]
    L = c["L"]
    q_ref = golden.words("qe_q", 32).reshape(3, L)
    for k, (alpha, e) in enumerate(zip(c["alphas"], c["element_bits"])):
        q = O.quantize(golden["qe_x"], golden["qe_u"][k], alpha, e)
        assert np.array_equal(q, q_ref[k]), k
    d = golden.cases("decode")[0]
    out = O.unquantize(golden.words("qd_v", 32), d["alpha"], d["element_bits"], d["n_clients"])
    assert np.array_equal(out.view(np.uint64), golden["qd_out"].view(np.uint64))


def test_dynamic_masking_cost_model():
    # jzf_flashe_block.py:89-117
    a, b, c = [0, 1, 2, 3], [2, 3, 4], [9]
    choice, sc, dc = O.dynamic_masking([a, b, c])
    assert sc == 16 and dc == 32 - 2 * 2 and choice == "single"
    choice, sc, dc = O.dynamic_masking([a, a, a])
    assert sc == 24 and dc == 48 - 2 * 8 and choice == "single"


def test_python_port_matches_golden(golden):
    """oracle/flashe_port.py (the timed CPU baseline) against the reference's own outputs."""
    from oracle import flashe_port as P
    for name in ("rt_b20_n3", "rt_b32_n5"):
        c = [k for k in golden.cases("roundtrip") if k["name"] == name][0]
        b, nj, L, n, it = (c[k] for k in ("int_bits", "n_jobs", "L", "n_clients", "iter"))
        xs, us = golden[name + "_x"], golden[name + "_u"]
        cts, agg, dec, out = P.run_round(golden.key, b, it, list(xs), c["alpha"], 16, n_jobs=nj, us=list(us))
        ct_ref = golden.words(name + "_ct", b).reshape(n, L)
        for k in range(n):
            assert [int(v) for v in cts[k]] == [int(v) for v in ct_ref[k]]
        assert [int(v) for v in agg] == [int(v) for v in golden.words(name + "_aggB", b)]
        assert [int(v) for v in dec] == [int(v) for v in golden.words(name + "_decB", b)]
        assert np.array_equal(np.array([float(v) for v in out]).view(np.uint64), golden[name + "_decoded"].view(np.uint64))
        assert [int(v) for v in P.aggregate_packed(cts, b)] == [int(v) for v in golden.words(name + "_aggA", b)]
    for c in golden.cases("dropout"):
        assert P.collapse_runs(c["survivors"]) == (c["add"], c["minus"])


def test_philox4x32_10_random123_known_answers():
    """Random123 kat_vectors for philox4x32-10 pin the restated device noise generator."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = O.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_noise_uniform_is_res53_of_philox():
    u = O.noise_uniform(0x0123456789ABCDEF, 7, 5, 1001)
    assert u.dtype == np.float64 and (u >= 0).all() and (u < 1).all()
    # element j uses counter j>>1: elements 6 and 7 come from one Philox call
    o = O.philox4x32_10(np.array([3], dtype=np.uint64), np.array([0], dtype=np.uint64), np.array([7], dtype=np.uint64),
                        np.array([0], dtype=np.uint64), 0x89ABCDEF, 0x01234567)
    w = [int(x[0]) for x in o]
    assert u[1] == ((w[0] >> 5) * 67108864.0 + (w[1] >> 6)) / 9007199254740992.0
    assert u[2] == ((w[2] >> 5) * 67108864.0 + (w[3] >> 6)) / 9007199254740992.0


def test_reference_itself_agrees_with_the_oracle_when_staged():
    """oracle/_ref (the reference's own modules, staged by oracle/make_ref.py) driven through a full round
    equals the C oracle on the same seeded inputs — the pin of the oracle that needs no fixture."""
    from oracle import ref_driver as R
    if not R.available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py in the build container)")
    key, b, nj, L, n, it, alpha = bytes(range(32)), 20, 5, 20011, 4, 3, 0.25
    xs = [(np.random.RandomState(50 + c).standard_normal(L) * 0.1).astype(np.float32) for c in range(n)]
    seeds = [60 + c for c in range(n)]
    survivors = [0, 1, 3]
    qs, cts, agg, dec, out = R.run_round(key, b, it, xs, alpha, 16, n_jobs=nj, seeds=seeds, inline_pool=True, survivors=survivors)
    for c in range(n):
        np.random.seed(seeds[c])
        q = O.quantize(xs[c], np.random.random(L), alpha, 16)
        assert np.array_equal(q, np.asarray(qs[c], dtype=object).astype(np.uint32))
        assert np.array_equal(O.encrypt(key, b, nj, it, c, "double", q), np.asarray(cts[c], dtype=object).astype(np.uint32))
    a = O.aggregate(b, np.stack([np.asarray(cts[c], dtype=object).astype(np.uint32) for c in survivors]))
    assert np.array_equal(a, np.asarray(agg, dtype=object).astype(np.uint32))
    d = O.decrypt(key, b, nj, it, survivors, "double", a)
    assert np.array_equal(d, np.asarray(dec, dtype=object).astype(np.uint32))
    assert np.array_equal(O.unquantize(d, alpha, 16, len(survivors)).view(np.uint64), np.asarray(out, dtype=np.float64).view(np.uint64))
