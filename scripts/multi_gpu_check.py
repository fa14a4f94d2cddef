#!/usr/bin/env python
"""Multi-GPU parity check on real GPUs, one rank per GPU over NCCL (run under torchrun).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multi_gpu_check.py

Every rank runs one round on ITS element range of the same job (encode+encrypt for all clients,
element-wise sum, packed-carry sum with the NCCL descriptor exchange, decrypt+decode); the shards
are then gathered with NCCL and every rank compares them bit for bit against the same round run
whole on its own GPU.  The CPU twin of this test is tests/test_sharding_gloo.py (gloo, oracle
arithmetic); here the arithmetic is the CUDA library and the exchange is NCCL over NVLink.
Prints one JSON line on rank 0; exits non-zero on any mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from flashe_b200.device import (AGG_ELEMENTWISE, AGG_PACKED, SCHEME_DOUBLE, CodecSpec, DeviceContext, NoiseSpec,  # noqa: E402
                                VectorSpan)
from flashe_b200.sharding import PeerGather, aggregate_packed_sharded, gather_shards, shard_bounds  # noqa: E402

KEY = bytes(range(32))


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    results = {}
    ok = True
    for bits, L, n, n_jobs, it in ((32, 4_000_003, 8, 16, 3), (20, 1_000_001, 5, 8, 1), (32, 1 << 20, 64, 1024, 0)):
        ctx = DeviceContext(KEY, bits, dev)
        g = torch.Generator(device=dev); g.manual_seed(77)            # same data on every rank
        x = torch.randn(n, L, generator=g, device=dev, dtype=torch.float32) * 0.1
        codec = CodecSpec(alpha=5.938345 * 0.1, element_bits=16, n_clients=n)
        noise = NoiseSpec(seed=9, stream=0)
        whole = VectorSpan(total_len=L, n_jobs=n_jobs)
        cts_w = ctx.encode_encrypt_batch(it, 0, SCHEME_DOUBLE, x, codec, noise, whole)
        # adversarial carries for the packed sum: a propagate-only run across the shard cuts
        begin, count = shard_bounds(L, world, rank)
        cuts = [shard_bounds(L, world, r)[0] for r in range(world)] + [L]
        ones = (1 << bits) - 1
        ones = ones - (1 << 32) if ones >= (1 << 31) else ones       # as an int32 bit pattern
        cw = cts_w.view(torch.int32)
        for cpos in cuts[1:-1]:
            cw[:, cpos - 3:cpos + 3] = 0
            cw[0, cpos - 3:cpos + 3] = ones
            cw[1, cpos + 2] = 1
        agg_w = ctx.aggregate(cts_w, AGG_ELEMENTWISE)
        aggp_w = ctx.aggregate(cts_w, AGG_PACKED)
        out_w = ctx.decrypt_decode(it, [n], [0], agg_w, codec, whole)

        span = VectorSpan(total_len=L, n_jobs=n_jobs, begin=begin, count=count)
        xs = x[:, begin:begin + count].contiguous()
        cts_s = ctx.encode_encrypt_batch(it, 0, SCHEME_DOUBLE, xs, codec, noise, span)
        cs = cts_s.view(torch.int32)
        for cpos in cuts[1:-1]:                                        # the same edits, in shard coordinates
            lo, hi = max(cpos - 3, begin), min(cpos + 3, begin + count)
            if lo < hi:
                cs[:, lo - begin:hi - begin] = 0
                cs[0, lo - begin:hi - begin] = ones
            if begin <= cpos + 2 < begin + count:
                cs[1, cpos + 2 - begin] = 1
        agg_s = ctx.aggregate(cts_s, AGG_ELEMENTWISE)
        aggp_s = aggregate_packed_sharded(ctx, cts_s)                 # NCCL: 16-byte descriptor all-gather + fix-up
        out_s = ctx.decrypt_decode(it, [n], [0], agg_s, codec, span)
        counts = [shard_bounds(L, world, r)[1] for r in range(world)]
        if world > 1:
            agg_g = gather_shards(agg_s, counts)
            aggp_g = gather_shards(aggp_s, counts)
            out_g = gather_shards(out_s, counts)
        else:
            agg_g, aggp_g, out_g = agg_s, aggp_s, out_s
        # the same decode with every rank's kernel storing its shard straight into rank 0's vector (peer memory over
        # NVLink, sharding.PeerGather): no gather step at all
        pg = PeerGather(ctx, L, root=0)
        ctx.decrypt_decode(it, [n], [0], agg_s, codec, span, out=pg.slice_of(begin, count))
        pg.finish()
        peer_ok = bool(torch.equal(pg.tensor().view(torch.int64), out_w.view(torch.int64))) if rank == 0 else True
        if world > 1 and L >= 4_000_000:                                # device-timed, max over ranks
            def timed(fn, steps=10):
                fn(); torch.cuda.synchronize(); dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            peer_slice = pg.slice_of(begin, count)
            results["gather_ms_b%d_L%d" % (bits, L)] = {
                "decode_then_nccl_all_gather": timed(lambda: gather_shards(ctx.decrypt_decode(it, [n], [0], agg_s, codec, span), counts)),
                "decode_into_peer_memory": timed(lambda: ctx.decrypt_decode(it, [n], [0], agg_s, codec, span, out=peer_slice)),
                "decode_local_only": timed(lambda: ctx.decrypt_decode(it, [n], [0], agg_s, codec, span))}
        pg.close()
        checks = {
            "decoded_gathered_through_peer_memory": peer_ok,
            "ciphertext_shard": bool(torch.equal(cts_s.view(torch.int32), cts_w[:, begin:begin + count].contiguous().view(torch.int32))),
            "aggregate_elementwise": bool(torch.equal(agg_g.view(torch.int32), agg_w.view(torch.int32))),
            "aggregate_packed_carry": bool(torch.equal(aggp_g.view(torch.int32), aggp_w.view(torch.int32))),
            "packed_differs_from_elementwise": bool(not torch.equal(aggp_w.view(torch.int32), agg_w.view(torch.int32))),
            "decoded_float64_bits": bool(torch.equal(out_g.view(torch.int64), out_w.view(torch.int64))),
        }
        results["b%d_L%d_n%d" % (bits, L, n)] = checks
        ok = ok and all(checks.values())
        ctx.close()
    if world > 1:
        # the gather at BASELINE config 5's size (100 M elements, float64 result = 800 MB on the owner), timed alone
        L, n, bits = 100_000_000, 64, 32
        ctx = DeviceContext(KEY, bits, dev)
        begin, count = shard_bounds(L, world, rank)
        counts = [shard_bounds(L, world, r)[1] for r in range(world)]
        span = VectorSpan(total_len=L, n_jobs=16, begin=begin, count=count)
        codec = CodecSpec(alpha=5.938345 * 0.1, element_bits=16, n_clients=n)
        agg_s = torch.randint(-2 ** 31, 2 ** 31 - 1, (count,), device=dev, dtype=torch.int32).view(torch.uint32)
        pg = PeerGather(ctx, L, root=0)
        peer_slice = pg.slice_of(begin, count)

        def timed(fn, steps=5):
            fn(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        local = ctx.decrypt_decode(0, [n], [0], agg_s, codec, span)
        ctx.decrypt_decode(0, [n], [0], agg_s, codec, span, out=peer_slice)
        pg.finish()
        same = bool(torch.equal(pg.tensor()[begin:begin + count].view(torch.int64), local.view(torch.int64))) if rank == 0 else True
        ok = ok and same
        results["gather_ms_100M"] = {
            "decode_then_nccl_all_gather": timed(lambda: gather_shards(ctx.decrypt_decode(0, [n], [0], agg_s, codec, span), counts)),
            "decode_into_peer_memory": timed(lambda: ctx.decrypt_decode(0, [n], [0], agg_s, codec, span, out=peer_slice)),
            "decode_local_only": timed(lambda: ctx.decrypt_decode(0, [n], [0], agg_s, codec, span)),
            "bytes_to_owner_per_writer": count * 8, "root_shard_equal": same}
        pg.close()
        ctx.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "all_ranks_ok": bool(flag.item()), "rank0": results}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
