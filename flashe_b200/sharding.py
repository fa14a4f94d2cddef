"""Element-range sharding of one gradient vector over the GPUs of a box (one process per GPU).

Every element's mask depends only on (key, iter, client, element index, L, n_jobs), so encode,
encrypt, the element-wise server sum, decrypt and decode shard with NO data-path collective: rank g
works on elements [begin_g, begin_g + count_g) of the whole vector and passes that range as the
`flashe_span` of every call.  Two optional exchanges exist:

  * the packed-carry server sum (jzf_aggregator.py:404-419) lets carries cross shard boundaries:
    each rank aggregates its shard with carry_in = 0 and emits a 4-word carry descriptor; one
    all-gather of those descriptors (16 bytes per rank) tells every rank its true carry-in, which a
    fix-up kernel ripples into the shard (in practice it touches one element);
  * gathering the shards into one tensor (when a caller wants the whole result on one GPU): either an
    NCCL all-gather afterwards (gather_shards), or no collective at all — PeerGather maps the owner's
    vector into every rank (flashe_peer_open) and the decode kernel of each rank stores its shard
    straight into it over NVLink, element by element as the PRF work proceeds.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

T_NEVER = 0xFFFFFFFF


def shard_bounds(total_len: int, world: int, rank: int, align: int = 4) -> Tuple[int, int]:
    """Contiguous element range of `rank`: equal pieces cut at multiples of `align` elements (16-byte
    aligned rows for 4-byte words); the last rank takes the remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    per = (total_len // world) // align * align
    begin = rank * per
    count = per if rank + 1 < world else total_len - begin
    return begin, count


def resolve_carry_ins(descs: Sequence[Sequence[int]]) -> List[int]:
    """descs[g] = (c0, depends, A, T) of shard g as written by flashe_aggregate(..., carry_in=0,
    carry_out=...): c0 = carry out of the shard's first element assuming carry_in = 0; when `depends`
    is set the exact transfer function of the shard is carry_out = A + (carry_in >= T).
    Shards are ordered by element range; the carry flows from the LAST shard towards shard 0.
    Returns the carry-in of every shard."""
    cins = [0] * len(descs)
    carry = 0
    for g in reversed(range(len(descs))):
        cins[g] = carry
        c0, depends, a, t = (int(v) & 0xFFFFFFFF for v in descs[g])
        if depends:
            carry = a + (1 if (t != T_NEVER and carry >= t) else 0)
        else:
            carry = c0
    return cins


def all_gather_descriptors(desc: torch.Tensor, group=None) -> List[List[int]]:
    """All-gather one 4-word descriptor per rank (NCCL for CUDA tensors, gloo for CPU tensors)."""
    world = dist.get_world_size(group)
    d = desc.to(torch.int64).reshape(4)
    bucket = [torch.empty_like(d) for _ in range(world)]
    dist.all_gather(bucket, d, group=group)
    return [[int(v) & 0xFFFFFFFF for v in b.cpu().tolist()] for b in bucket]


def aggregate_packed_sharded(ctx, cts_shard: torch.Tensor, group=None) -> torch.Tensor:
    """Packed-carry server sum of this rank's shard of the [n, L] ciphertext matrix, exact across
    shard boundaries.  cts_shard: words [n, count_g] on ctx.device; every shard must be non-empty."""
    from .device import AGG_PACKED
    desc = torch.zeros(4, dtype=torch.int32, device=ctx.device)
    out = ctx.aggregate(cts_shard, AGG_PACKED, carry_in=0, carry_out=desc)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        descs = all_gather_descriptors(desc, group)
        cin = resolve_carry_ins(descs)[dist.get_rank(group)]
        if cin:
            ctx.aggregate_carry_fixup(out, cin)
    return out


def gather_shards(shard: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """Re-assemble element-range shards into the whole vector on every rank (NCCL all-gather)."""
    world = dist.get_world_size(group)
    if world == 1:
        return shard
    tail = list(shard.shape[1:])
    pad = max(counts)
    view = shard.view(torch.int64) if shard.element_size() == 8 else shard.view(torch.int32)
    buf = torch.zeros([pad] + tail, dtype=view.dtype, device=shard.device)
    buf[:shard.shape[0]] = view
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    whole = torch.cat([p[:c] for p, c in zip(parts, counts)])
    return whole.view(shard.dtype)


class PeerGather(object):
    """The whole decoded vector on ONE GPU without a gather step: rank `root` owns `total_len` float64 (device memory
    exported with flashe_peer_alloc), every other rank maps it (flashe_peer_open) and hands `slice_of(begin, count)`
    as the `out=` of DeviceContext.decrypt_decode / decode — the kernel's stores then land in the owner's memory
    over NVLink while it is still computing the rest of its shard.  `finish()` synchronises the writers and the
    owner; `tensor()` (root only) views the owner's memory as a torch tensor.

    The reference has no counterpart: its arbiter is a single CPU process (proc/jzf_aggregator.py:404-430)."""

    def __init__(self, ctx, total_len, root=0, group=None, itemsize=8):
        self.ctx, self.total_len, self.root, self.group, self.itemsize = ctx, int(total_len), root, group, itemsize
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._owned = self._mapped = None
        handle = [None]
        if self.rank == root:
            self._owned, handle[0] = ctx.peer_alloc(self.total_len * itemsize)
        if self.world > 1:
            dist.broadcast_object_list(handle, src=root, group=group)
            if self.rank != root:
                self._mapped = ctx.peer_open(handle[0])
        self.base = self._owned if self.rank == root else self._mapped

    def slice_of(self, begin, count):
        from .device import PeerSlice
        if begin < 0 or begin + count > self.total_len:
            raise ValueError("slice outside the gathered vector")
        return PeerSlice(self.base + begin * self.itemsize, count * self.itemsize)

    def finish(self):
        """Every writer's stream has completed and the owner may read."""
        torch.cuda.synchronize(self.ctx.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def tensor(self, dtype=torch.float64):
        """Root only: a copy-free torch view of the owner's memory (valid until close())."""
        if self.rank != self.root:
            raise RuntimeError("only the owning rank can view the gathered vector")
        import ctypes
        n = self.total_len * self.itemsize
        iface = {"shape": (n,), "typestr": "|u1", "data": (self._owned, False), "version": 3, "strides": None}
        holder = type("PeerMemory", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.ctx.device).view(dtype)

    def close(self):
        if self._mapped is not None:
            self.ctx.peer_close(self._mapped)
            self._mapped = None
        if self.world > 1:
            dist.barrier(group=self.group)            # nobody maps the buffer any more
        if self._owned is not None:
            self.ctx.peer_free(self._owned)
            self._owned = None
