"""BASELINE.json configs C1-C5 at their FULL sizes on the GPU.

C1 and C2 are checked bit for bit against the oracle in full.  C3-C5 are too large for the CPU oracle to
redo in seconds, so they combine (i) oracle checks of whole reference chunks / whole compact vectors
sampled from the full-size run with (ii) size-independent properties evaluated on the whole vectors
on the device: decrypt(aggregate(encrypt(q_c))) == sum_c q_c mod 2^b, precomputed == on-the-fly,
shard concatenation == whole vector.  Tolerance 0 everywhere (decoded float64 compared as bits).
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

KEY = bytes(range(32))
ALPHA = float(5.938345 * 0.1)


def _np(t):
    return t.cpu().numpy()


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda")


@pytest.fixture(scope="module")
def fb():
    import flashe_b200
    return flashe_b200


def _i64(words):
    """uint32 word tensor -> int64 tensor (torch has no uint32 arithmetic)."""
    return words.view(torch.int32).to(torch.int64) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ C1
def test_c1_reference_cpu_case_1m_3_clients_bit_exact(fb):
    """BASELINE config 1 exactly: 1M-element float32 gradient, 3 clients, int_bits 20, double masking —
    the reference's own CPU-runnable case.  Every ciphertext, BOTH server sums (element-wise and the
    packed-carry sum of the wire integers), the decrypted integers and the decoded float64 against the C
    oracle; and, when the reference's own modules are staged (oracle/_ref, built by oracle/make_ref.py in
    the build container), against jzf_flashe.FlasheCipher / jzf_quantize themselves on the same inputs."""
    L, n, bits, n_jobs, it = 1_000_000, 3, 20, 8, 0
    ctx = fb.DeviceContext(KEY, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    xs = [(np.random.RandomState(1000 + c).standard_normal(L) * 0.1).astype(np.float32) for c in range(n)]
    seeds = [2000 + c for c in range(n)]
    u = np.empty((n, L), dtype=np.float64)
    for c in range(n):
        np.random.seed(seeds[c])
        u[c] = np.random.random(L)
    q_out = torch.empty((n, L), dtype=torch.int32, device="cuda").view(torch.uint32)
    cts = ctx.empty_words(L, rows=n)
    for c in range(n):                     # client by client, as the parties would (one launch each)
        ctx.encode_encrypt(it, c, fb.SCHEME_DOUBLE, _dev(xs[c]), codec, fb.NoiseSpec(u=_dev(u[c])), span, out=cts[c], q_out=q_out[c])
    O.set_threads(8)
    q = np.stack([O.quantize(xs[c], u[c], ALPHA, 16) for c in range(n)])
    ct_want = np.stack([O.encrypt(KEY, bits, n_jobs, it, c, "double", q[c]) for c in range(n)])
    assert np.array_equal(_np(q_out), q)
    assert np.array_equal(_np(cts), ct_want)
    got = {}
    for mode, name in ((fb.AGG_ELEMENTWISE, "elementwise"), (fb.AGG_PACKED, "packed")):
        agg = ctx.aggregate(cts, mode)
        agg_want = O.aggregate(bits, ct_want, name)
        assert np.array_equal(_np(agg), agg_want), name
        p = ctx.empty_words(L)
        out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
        p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", agg_want)
        assert np.array_equal(_np(p), p_want), name
        assert np.array_equal(_np(out).view(np.uint64), O.unquantize(p_want, ALPHA, 16, n).view(np.uint64)), name
        got[name] = (_np(agg), _np(p), _np(out))
    assert np.array_equal(got["elementwise"][1].astype(np.int64), q.astype(np.int64).sum(axis=0))
    from oracle import ref_driver as R
    if R.available():                      # the reference itself (in-process pool: this process holds a CUDA context)
        qs_r, cts_r, agg_r, dec_r, out_r = R.run_round(KEY, bits, it, xs, ALPHA, 16, n_jobs=n_jobs, seeds=seeds, inline_pool=True)
        assert np.array_equal(np.stack([np.asarray(v, dtype=object).astype(np.uint32) for v in cts_r]), _np(cts))
        assert np.array_equal(np.asarray(agg_r, dtype=object).astype(np.uint32), got["elementwise"][0])
        assert np.array_equal(np.asarray(dec_r, dtype=object).astype(np.uint32), got["elementwise"][1])
        assert np.array_equal(np.asarray(out_r, dtype=np.float64).view(np.uint64), got["elementwise"][2].view(np.uint64))


# ------------------------------------------------------------------------------------------------ C2
def test_c2_cnn_sized_gradient_10_clients_bit_exact(fb):
    """2.5M-element gradient, 10 clients, int_bits 20, double masking: every ciphertext, both server
    sums, the decrypted integers and the decoded floats against the oracle, noise = the seeded
    np.random.random stream the reference would draw."""
    L, n, bits, n_jobs, it = 2_500_000, 10, 20, 8, 0
    ctx = fb.DeviceContext(KEY, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    x = np.stack([(np.random.RandomState(1000 + c).standard_normal(L) * 0.1).astype(np.float32) for c in range(n)])
    u = np.empty((n, L), dtype=np.float64)
    for c in range(n):
        np.random.seed(2000 + c)
        u[c] = np.random.random(L)
    cts = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(u=_dev(u).reshape(-1)), span)
    O.set_threads(8)
    q = np.stack([O.quantize(x[c], u[c], ALPHA, 16) for c in range(n)])
    ct_want = np.stack([O.encrypt(KEY, bits, n_jobs, it, c, "double", q[c]) for c in range(n)])
    assert np.array_equal(_np(cts), ct_want)
    # shared-stream launch produces the same ciphertexts
    cts2 = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, _dev(x), codec, fb.NoiseSpec(u=_dev(u).reshape(-1)), span,
                                    share_streams=True)
    assert torch.equal(cts.view(torch.int32), cts2.view(torch.int32))
    for mode, name in ((fb.AGG_ELEMENTWISE, "elementwise"), (fb.AGG_PACKED, "packed")):
        agg = ctx.aggregate(cts, mode)
        agg_want = O.aggregate(bits, ct_want, name)
        assert np.array_equal(_np(agg), agg_want), name
        p = ctx.empty_words(L)
        out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
        p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", agg_want)
        assert np.array_equal(_np(p), p_want), name
        assert np.array_equal(_np(out).view(np.uint64), O.unquantize(p_want, ALPHA, 16, n).view(np.uint64)), name
    # element-wise: the plaintext sum exactly
    p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", O.aggregate(bits, ct_want))
    assert np.array_equal(p_want.astype(np.int64), q.astype(np.int64).sum(axis=0))


# ------------------------------------------------------------------------------------------------ C3
def test_c3_precompute_16_rounds_25m_with_dropout(fb):
    """Mask precomputation for 16 future iterations x 25M elements, double masking, 2 of 10 clients
    dropping per round.  (a) ring ciphertext == on-the-fly ciphertext on all 25M elements, every round;
    (b) ring slots == oracle masks on whole reference chunks (n_jobs = 1024 -> 24,414-element chunks);
    (c) per round: decrypt of the survivors' aggregate with the run-collapsed index sets == sum of the
    survivors' plaintexts on all 25M elements, and == the oracle on sampled chunks."""
    L, n, bits, n_jobs, rounds = 25_000_000, 10, 20, 1024, 16
    ctx = fb.DeviceContext(KEY, bits)
    span = fb.VectorSpan(L, n_jobs)
    me = 3                                                   # the client that precomputes
    ring = fb.MaskRing.for_encrypt(ctx, me, span, rounds, "double").fill(0)
    assert ring.nbytes == rounds * L * 4                     # 1.6 GB: combined term, one word per element
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    q = torch.randint(0, 65536, (n, L), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
    d, r = divmod(L, n_jobs)
    chunk_ids = [0, r - 1, r, n_jobs // 2, n_jobs - 1]       # incl. the (d+1)/d boundary and both ends
    mask20 = (1 << bits) - 1
    O.set_threads(8)
    for t in range(rounds):
        # (b) oracle masks on whole chunks
        for k in chunk_ids[:3] if t % 4 else chunk_ids:
            j0, j1 = O.chunk_bounds(L, n_jobs, k)
            want = O.masks(KEY, bits, n_jobs, t, [me, me + 1], [1, -1], L, j0=j0, cnt=j1 - j0)
            assert np.array_equal(_np(ring.peek(t)[j0:j1]), want), (t, k)
        # (a) online step with the precomputed mask == on-the-fly encrypt
        ct_fly = ctx.encrypt(t, me, fb.SCHEME_DOUBLE, q[me], span)
        ct_ring = ring.encrypt(t, q[me])
        assert torch.equal(ct_fly.view(torch.int32), ct_ring.view(torch.int32)), t
        assert not ring.has(t)                               # consumed, like the reference's buffers
        # (c) dropout
        rs = np.random.RandomState(3000 + t)
        dropped = {0: [0, 5], 1: [n - 1, 2], 2: [0, n - 1]}.get(t, sorted(rs.choice(n, 2, replace=False).tolist()))
        alive = [c for c in range(n) if c not in dropped]
        cts = ctx.empty_words(L, rows=len(alive))
        for row, c in enumerate(alive):
            if c == me:
                cts[row].copy_(ct_ring)
            else:
                ctx.encrypt(t, c, fb.SCHEME_DOUBLE, q[c], span, out=cts[row])
        agg = ctx.aggregate(cts)
        add, minus = O.collapse_runs(alive)
        p = ctx.decrypt(t, add, minus, agg, span)
        want_sum = torch.zeros(L, dtype=torch.int64, device="cuda")
        for c in alive:
            want_sum += _i64(q[c])
        assert torch.equal(_i64(p), want_sum & mask20), (t, dropped)
        k = chunk_ids[t % len(chunk_ids)]
        j0, j1 = O.chunk_bounds(L, n_jobs, k)
        p_oracle = O.decrypt(KEY, bits, n_jobs, t, alive, "double", _np(agg[j0:j1]), L=L, j0=j0)
        assert np.array_equal(_np(p[j0:j1]), p_oracle), (t, k)
        del cts, agg, p, want_sum
    # refill: the ring rolls forward to rounds 16..31
    ring.fill(rounds, 2)
    assert ring.has(rounds) and ring.has(rounds + 1) and not ring.has(rounds + 2)
    j0, j1 = O.chunk_bounds(L, n_jobs, 7)
    assert np.array_equal(_np(ring.peek(rounds + 1)[j0:j1]),
                          O.masks(KEY, bits, n_jobs, rounds + 1, [me, me + 1], [1, -1], L, j0=j0, cnt=j1 - j0))


def test_c3_flashecipher_prepare_rounds_and_prepared_decrypt_under_dropout(fb):
    """The drop-in class on device tensors: prepare_encrypt(rounds=16) + prepare_decrypt() with
    survivor sets that drop the edge clients (where the reference itself decrypts wrongly, SURVEY
    §0.5) — precomputed == on-the-fly == plaintext sum."""
    from flashe_b200.secureprotol import FlasheCipher
    L, n, bits, n_jobs = 2_000_000, 10, 20, 16
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    q = torch.randint(0, 65536, (n, L), dtype=torch.int32, device="cuda", generator=g).view(torch.uint32)
    ciphers = []
    for c in range(n):
        ci = FlasheCipher(bits, n_jobs=n_jobs)
        ci.idx = c
        ci.set_num_clients(n)
        ci.generate_prp_seed(KEY)
        ci.set_num_params(L)
        ciphers.append(ci)
    assert ciphers[0].iter_index == -1                       # as at cipher creation (jzf_flashe_block.py:227-231)
    ciphers[0].prepare_encrypt(rounds=16)                    # rounds 0..15 for client 0
    ctx = fb.DeviceContext(KEY, bits)
    for t, dropped in enumerate([[], [0], [n - 1], [0, n - 1], [4, 5], [1, 8]]):
        alive = [c for c in range(n) if c not in dropped]
        cts = []
        for c in alive:
            ciphers[c].set_iter_index(t)
            ct = ciphers[c].encrypt(q[c])
            fly = ctx.encrypt(t, c, fb.SCHEME_DOUBLE, q[c], fb.VectorSpan(L, n_jobs))
            assert torch.equal(ct.view(torch.int32), fly.view(torch.int32)), (t, c)
            cts.append(ct)
        agg = ctx.aggregate(torch.stack([c_.view(torch.int32) for c_ in cts]).view(torch.uint32))
        dec = ciphers[alive[0]]
        dec.prepare_decrypt()                                # F(t,n) - F(t,0) ahead of the download
        dec.set_idx_list(alive, mode="decrypt")
        p = dec.decrypt(agg)
        want = torch.zeros(L, dtype=torch.int64, device="cuda")
        for c in alive:
            want += _i64(q[c])
        assert torch.equal(_i64(p), want & ((1 << bits) - 1)), (t, dropped)


# ------------------------------------------------------------------------------------------------ C4
def test_c4_sparse_top1pct_of_50m_32_clients(fb):
    """Index-sparse path: total = 50M, k = 500k per client, 32 clients, int_bits 32, masking scheme
    single.  Compact ciphertexts against the oracle in full (16M masks); expand_to_dense + element-wise
    sum + per-index unmasking against the plaintext dense sum on all 50M elements; overlap counts of
    dynamic_masking against numpy."""
    total, k, n, bits, n_jobs, it = 50_000_000, 500_000, 32, 32, 16, 5
    ctx = fb.DeviceContext(KEY, bits)
    O.set_threads(8)
    index, qs, zeros = [], [], []
    for c in range(n):
        rs = np.random.RandomState(4000 + c)
        idx = np.unique(rs.randint(0, total, size=k + k // 50))
        idx = np.sort(rs.permutation(idx)[:k]).astype(np.int64)
        assert idx.size == k
        index.append(idx)
        qs.append(rs.randint(0, 65536, size=k).astype(np.uint32))
        zeros.append(int(O.quantize(np.zeros(1, np.float32), np.array([rs.rand()]), 1.0, 16)[0]))   # 'zzz' layer, alpha 1.0
    span = fb.VectorSpan(k, n_jobs)
    dense_sum = torch.zeros(total, dtype=torch.int64, device="cuda")
    acc = ctx.zeros_words(total)
    index_d = []
    for c in range(n):
        ct = ctx.encrypt(it, c, fb.SCHEME_SINGLE, _dev(qs[c]), span)
        assert np.array_equal(_np(ct), O.encrypt(KEY, bits, n_jobs, it, c, "single", qs[c])), c
        idx_d = _dev(index[c])
        index_d.append(idx_d)
        dense = ctx.sparse_expand(ct, idx_d, total, zeros[c])              # expand_to_dense, jzf_aggregator.py:150-165
        if c == 0:                                                         # one client's dense vector against the oracle
            assert np.array_equal(_np(dense), O.expand_to_dense(bits, _np(ct), index[c], total, zeros[c]))
        acc = ctx.aggregate(torch.stack([acc.view(torch.int32), dense.view(torch.int32)]).view(torch.uint32))
        plain = torch.full((total,), zeros[c], dtype=torch.int64, device="cuda")
        plain[idx_d] = _dev(qs[c].astype(np.int64))
        dense_sum += plain
        del dense, plain
    # decrypt: for every client regenerate F(t, c) over its compact positions and subtract at its indices
    p = acc.clone()
    for c in range(n):
        ctx.sparse_apply_masks(it, [c], [-1], span, index_d[c], p)
    assert torch.equal(_i64(p), dense_sum & 0xFFFFFFFF)
    # the same through the one-call forms: tiled sparse sum on the server, every client's masks in one call on the client
    cts = [ctx.encrypt(it, c, fb.SCHEME_SINGLE, _dev(qs[c]), span) for c in range(n)]
    fused = ctx.sparse_sum(cts, index_d, total, zeros)
    assert torch.equal(fused.view(torch.int32), acc.view(torch.int32))
    ctx.sparse_apply_masks_batch(it, list(range(n)), -1, n_jobs, index_d, fused)
    assert torch.equal(_i64(fused), dense_sum & 0xFFFFFFFF)
    # cost model inputs
    ov = ctx.sparse_overlap(index_d, total)
    want = [int(np.intersect1d(index[i], index[i + 1], assume_unique=True).size) for i in range(n - 1)]
    assert ov == want


# ------------------------------------------------------------------------------------------------ C5
def test_c5_100m_elements_64_clients(fb):
    """The benchmark workload itself: L = 100M float32, 64 clients, int_bits 32, double masking, device
    noise.  (i) plaintext-sum property on all 100M elements; (ii) ciphertexts, aggregate, decrypted
    integers and decoded floats of whole reference chunks (n_jobs = 1024 -> 97,657-element chunks)
    against the oracle for 4 clients; (iii) two element-range shards reproduce the whole vector."""
    L, n, bits, n_jobs, it, seed = 100_000_000, 64, 32, 1024, 1, 0x5EED
    ctx = fb.DeviceContext(KEY, bits)
    span = fb.VectorSpan(L, n_jobs)
    codec = fb.CodecSpec(alpha=ALPHA, element_bits=16, n_clients=n)
    noise = fb.NoiseSpec(seed=seed, stream=0)
    x = torch.empty((n, L), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda")
    for c in range(n):
        g.manual_seed(1000 + c)
        x[c].normal_(0.0, 0.1, generator=g)
    cts = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, x, codec, noise, span)
    agg = ctx.aggregate(cts)
    p = ctx.empty_words(L)
    out = ctx.decrypt_decode(it, [n], [0], agg, codec, span, p_out=p)
    # (i) sum of the quantised plaintexts (same noise streams), all elements
    qsum = torch.zeros(L, dtype=torch.int64, device="cuda")
    qbuf = torch.empty(L, dtype=torch.uint32, device="cuda")
    for c in range(n):
        ctx.encode(x[c], codec, fb.NoiseSpec(seed=seed, stream=c), span, out=qbuf)
        qsum += _i64(qbuf)
    assert torch.equal(_i64(p), qsum & 0xFFFFFFFF)
    assert int(qsum.max()) <= n * 65536
    # (ii) oracle on whole chunks
    O.set_threads(8)
    d, r = divmod(L, n_jobs)
    for k in (0, r - 1, r, n_jobs - 1):
        j0, j1 = O.chunk_bounds(L, n_jobs, k)
        cnt = j1 - j0
        for c in (0, 1, 31, 63):
            u = O.noise_uniform(seed, c, j0, cnt)
            qc = O.quantize(_np(x[c, j0:j1]), u, ALPHA, 16)
            want = O.encrypt(KEY, bits, n_jobs, it, c, "double", qc, L=L, j0=j0)
            assert np.array_equal(_np(cts[c, j0:j1]), want), (k, c)
        agg_want = O.aggregate(bits, _np(cts[:, j0:j1]))
        assert np.array_equal(_np(agg[j0:j1]), agg_want)
        p_want = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", agg_want, L=L, j0=j0)
        assert np.array_equal(_np(p[j0:j1]), p_want)
        assert np.array_equal(_np(out[j0:j1]).view(np.uint64), O.unquantize(p_want, ALPHA, 16, n).view(np.uint64))
    # (iii) element-range shards (the multi-GPU layout) of 3 clients == the whole-vector result
    from flashe_b200.sharding import shard_bounds
    for world in (2, 8):
        for rank in (0, world - 1):
            b0, cnt = shard_bounds(L, world, rank)
            sp = fb.VectorSpan(L, n_jobs, b0, cnt)
            xs = x[:3, b0:b0 + cnt].contiguous()
            part = ctx.encode_encrypt_batch(it, 0, fb.SCHEME_DOUBLE, xs, codec, noise, sp)
            assert torch.equal(part.view(torch.int32), cts[:3, b0:b0 + cnt].contiguous().view(torch.int32)), (world, rank)
            o2 = ctx.decrypt_decode(it, [n], [0], agg[b0:b0 + cnt].contiguous(), codec, sp)
            assert torch.equal(o2.view(torch.int64), out[b0:b0 + cnt].view(torch.int64)), (world, rank)
