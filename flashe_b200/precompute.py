"""Mask precomputation buffers (reference: FlasheCipher.prepare_encrypt / prepare_decrypt and the
`next_iter_{encrypt,decrypt}_prepared` buffers, federatedml/secureprotol/jzf_flashe.py:596-666).

The reference looks ONE round ahead and keeps the add / minus streams as two object arrays.  On the
GPU the idle-time budget buys many rounds: `MaskRing` holds the COMBINED term
(sum_k sign_k F(t, prf_k)) mod 2^b of `rounds` future rounds in one [rounds, L] device buffer
(BASELINE config 3: 16 rounds x 25 M elements = 1.6 GB per client at 4-byte words, half of what the
reference's two-array layout would need).  A slot is consumed by the round it was made for and then
refilled for round t + rounds, so the buffer is a ring indexed by t mod rounds.

Contract: precomputed ciphertext == on-the-fly ciphertext, bit for bit.
"""
from typing import Optional, Sequence

import torch

from .device import CodecSpec, DeviceContext, NoiseSpec, VectorSpan


class MaskRing(object):

    def __init__(self, ctx: DeviceContext, prf_idx: Sequence[int], sign: Sequence[int], span: VectorSpan, rounds: int):
        if rounds < 1:
            raise ValueError("rounds must be >= 1")
        self.ctx = ctx
        self.prf_idx = [int(i) for i in prf_idx]
        self.sign = [int(s) for s in sign]
        self.span = span
        self.rounds = int(rounds)
        self.buf = ctx.empty_words(span.n, rows=self.rounds)
        self.slot_iter = [None] * self.rounds          # which round each slot currently holds

    @classmethod
    def for_encrypt(cls, ctx, idx, span, rounds, scheme="double"):
        """Client `idx`: F(t, idx) - F(t, idx+1) (double, jzf_flashe.py:601-605) or F(t, idx) (single)."""
        if scheme == "double":
            return cls(ctx, [idx, idx + 1], [1, -1], span, rounds)
        return cls(ctx, [idx], [1], span, rounds)

    @property
    def nbytes(self):
        return self.buf.numel() * self.buf.element_size()

    def fill(self, iter_from: int, count: Optional[int] = None):
        """Generate the masks of rounds iter_from .. iter_from+count-1 (default: the whole ring) into
        their slots t mod rounds.  Consecutive slots are written by one flashe_precompute call."""
        count = self.rounds if count is None else int(count)
        if count > self.rounds:
            raise ValueError("cannot hold more than %d rounds" % self.rounds)
        t = int(iter_from)
        left = count
        while left > 0:
            slot = t % self.rounds
            run = min(left, self.rounds - slot)
            self.ctx.precompute(t, run, self.prf_idx, self.sign, self.span, out=self.buf[slot:slot + run])
            for r in range(run):
                self.slot_iter[slot + r] = t + r
            t += run
            left -= run
        return self

    def has(self, it: int) -> bool:
        return self.slot_iter[int(it) % self.rounds] == int(it)

    def peek(self, it: int) -> torch.Tensor:
        if not self.has(it):
            raise KeyError("round %d is not in the ring" % it)
        return self.buf[int(it) % self.rounds]

    def take(self, it: int, refill: bool = False) -> torch.Tensor:
        """Mask of round `it`; the slot is marked consumed (the reference deletes its buffers after
        use, jzf_flashe.py:483-486).  With refill=True the caller promises to be done with the returned
        view before the next fill (same stream order) and the slot is regenerated for it + rounds."""
        m = self.peek(it)
        self.slot_iter[int(it) % self.rounds] = None
        if refill:
            m = m.clone()
            self.fill(int(it) + self.rounds, 1)
        return m

    # ------------------------------------------------------------------ the online step
    def encrypt(self, it: int, q: torch.Tensor, out=None) -> torch.Tensor:
        """ct = (q + mask_t) mod 2^b (consumption of next_iter_encrypt_prepared, jzf_flashe.py:480-486)."""
        return self.ctx.add_premasked(q, self.take(it), +1, out=out)

    def encode_encrypt(self, it: int, x: torch.Tensor, codec: CodecSpec, noise: NoiseSpec, out=None) -> torch.Tensor:
        """Fused quantise + add of the precomputed mask: 12 bytes of HBM traffic per element."""
        return self.ctx.encode_add_premasked(x, codec, noise, self.take(it), self.span, out=out)
