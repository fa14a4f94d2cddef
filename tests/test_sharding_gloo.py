"""N>1 host logic on CPU: world_size-2 gloo processes shard a vector by element range.  The oracle
stands in for the device kernels (test infrastructure only) so that the span arithmetic, the carry
descriptor exchange and the shard re-assembly are exercised without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flashe_b200.sharding import T_NEVER, resolve_carry_ins, shard_bounds
from oracle import oracle as O

KEY = bytes(range(32))


def test_shard_bounds_cover_and_align():
    for L in (1, 7, 1000, 100_000_003):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                b, c = shard_bounds(L, world, r)
                assert b == prev and (b % 4 == 0 or r == 0 or c == 0 or b == prev)
                prev = b + c
            assert prev == L


def _descriptor(cts, bits, n):
    """Python statement of the 4-word descriptor flashe_aggregate writes (carry_in = 0)."""
    out, c0 = O.aggregate(bits, cts, "packed", carry_in=0, return_carry=True)
    # walk from the END composing cin -> A + (cin >= T)
    A, T, started, resolved = 0, T_NEVER, False, False
    for j in reversed(range(cts.shape[1])):
        s = int(cts[:, j].astype(object).sum())
        H, lo = s >> bits, s & ((1 << bits) - 1)
        thr = (1 << bits) - lo if lo else T_NEVER
        f = (H, thr if thr < 0x7fffffff else T_NEVER)
        if not started:
            A, T = f
        else:                           # f is applied after (outer), (A,T) first (inner)
            lo_hit, hi_hit = A >= f[1], (T != T_NEVER) and (A + 1 >= f[1])
            if T == T_NEVER or lo_hit == hi_hit:
                A, T = f[0] + (1 if lo_hit else 0), T_NEVER
            else:
                A = f[0]
        started = True
        if T == T_NEVER:
            resolved = True
            break
    return out, [c0, 0 if resolved else 1, A, T]


def test_resolve_carry_ins_matches_whole_vector_sum():
    rs = np.random.RandomState(2)
    for bits in (20, 32):
        n, L = 5, 64
        cts = rs.randint(0, 1 << bits, size=(n, L), dtype=np.uint64).astype(np.uint32)
        cts[:, 20:45] = 0
        cts[0, 20:45] = (1 << bits) - 1          # a propagate-only run spanning whole shards
        cts[1, 44] = 1
        want = O.aggregate(bits, cts, "packed")
        for cuts in ([0, 16, 32, 48, 64], [0, 24, 28, 40, 64], [0, 44, 45, 46, 64]):
            outs, descs = [], []
            for a, e in zip(cuts[:-1], cuts[1:]):
                o, d = _descriptor(np.ascontiguousarray(cts[:, a:e]), bits, n)
                outs.append(o); descs.append(d)
            cins = resolve_carry_ins(descs)
            fixed = []
            for (a, e), cin in zip(zip(cuts[:-1], cuts[1:]), cins):
                fixed.append(O.aggregate(bits, np.ascontiguousarray(cts[:, a:e]), "packed", carry_in=cin))
            assert np.array_equal(np.concatenate(fixed), want), (bits, cuts)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flashe_b200.sharding import all_gather_descriptors, gather_shards
        bits, n_jobs, L, n, it = 20, 8, 4001, 3, 4
        rs = np.random.RandomState(11)
        q_all = rs.randint(0, 65536, size=(n, L)).astype(np.uint32)
        begin, count = shard_bounds(L, world, rank)
        # per-client encrypt of THIS rank's element range, global counters via (L, n_jobs, begin)
        cts = np.stack([O.encrypt(KEY, bits, n_jobs, it, c, "double", q_all[c, begin:begin + count], L=L, j0=begin)
                        for c in range(n)])
        # element-wise sum + decrypt need no exchange at all
        agg = O.aggregate(bits, cts)
        dec = O.decrypt(KEY, bits, n_jobs, it, list(range(n)), "double", agg, L=L, j0=begin)
        assert np.array_equal(dec, q_all[:, begin:begin + count].sum(axis=0).astype(np.uint32))
        # packed sum: one descriptor all-gather, then a local fix-up
        out0, desc = _descriptor(cts, bits, n)
        descs = all_gather_descriptors(torch.tensor(desc, dtype=torch.int64))
        cin = resolve_carry_ins(descs)[rank]
        mine = O.aggregate(bits, cts, "packed", carry_in=cin)
        counts = [shard_bounds(L, world, r)[1] for r in range(world)]
        whole = gather_shards(torch.from_numpy(mine.view(np.int32)), counts).numpy().view(np.uint32)
        full_ct = np.stack([O.encrypt(KEY, bits, n_jobs, it, c, "double", q_all[c]) for c in range(n)])
        assert np.array_equal(whole, O.aggregate(bits, full_ct, "packed"))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_round():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
