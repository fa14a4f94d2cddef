"""Tensor-native host API over the C ABI (torch supplies device memory and the current stream).

This is the fast entry the benchmark and the drop-in classes share: every method enqueues CUDA
kernels from libflashe_b200.so on torch's current stream of the context's device and returns torch
tensors.  There is no CPU path here: constructing a DeviceContext without a CUDA device or without
the built library raises.

Word tensors: int_bits <= 32 -> torch.uint32 [L]; <= 64 -> torch.uint64 [L]; <= 128 -> torch.uint64
[L, 2] (lo, hi).  Any tensor of the right byte size and contiguity is accepted as input (e.g. int32).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _cabi
from ._cabi import AGG_ELEMENTWISE, AGG_PACKED, SCHEME_DOUBLE, SCHEME_SINGLE  # noqa: F401


@dataclass(frozen=True)
class VectorSpan:
    """Which elements of which chunked vector (include/flashe_b200.h: flashe_span)."""
    total_len: int
    n_jobs: int
    begin: int = 0
    count: Optional[int] = None

    def c(self):
        cnt = self.total_len - self.begin if self.count is None else self.count
        return _cabi.Span(self.total_len, self.begin, cnt, self.n_jobs, 0)

    @property
    def n(self):
        return self.total_len - self.begin if self.count is None else self.count


@dataclass
class CodecSpec:
    """Encode/decode parameters (include/flashe_b200.h: flashe_codec).  `alpha` may be one Python
    float (single layer) or a list with `seg_end` giving each layer's end offset in the flat vector."""
    alpha: object
    element_bits: int = 16
    n_clients: int = 1
    seg_end: Optional[Sequence[int]] = None
    batch_lane_bits: int = 0     # lane batching (int_bits 65..128): lanes of element_bits + ceil(log2(num_clients)) bits;
                                 # spans then count WORDS and seg_end (required) counts ELEMENTS

    def c(self, total_len):
        alphas = [float(self.alpha)] if not isinstance(self.alpha, (list, tuple)) else [float(a) for a in self.alpha]
        if self.batch_lane_bits and self.seg_end is None:
            raise ValueError("lane batching needs seg_end (layer ends in elements)")
        ends = [int(total_len)] if self.seg_end is None else [int(e) for e in self.seg_end]
        if len(alphas) != len(ends):
            raise ValueError("alpha and seg_end must have the same length")
        self._ends = (C.c_uint64 * len(ends))(*ends)        # keep alive for the call
        self._alphas = (C.c_double * len(alphas))(*alphas)
        return _cabi.Codec(self.element_bits, self.n_clients, len(ends), int(self.batch_lane_bits), self._ends, self._alphas)

    def word_ends(self, int_bits):
        """Lane batching: cumulative words per layer (every layer padded to a multiple of int_bits // lane)."""
        bs = int_bits // self.batch_lane_bits
        out, prev, words = [], 0, 0
        for e in self.seg_end:
            words += (int(e) - prev + bs - 1) // bs
            out.append(words)
            prev = int(e)
        return out

    def span_elements(self, int_bits, span):
        """Elements covered by the words [span.begin, span.begin + span.n): (first element, count).  Without
        lane batching words are elements."""
        if not self.batch_lane_bits:
            return span.begin, span.n
        bs = int_bits // self.batch_lane_bits
        wends = self.word_ends(int_bits)

        def elem_of(word):
            prev_e = prev_w = 0
            for e, w in zip(self.seg_end, wends):
                if word < w:
                    return min(prev_e + (word - prev_w) * bs, int(e))
                prev_e, prev_w = int(e), w
            return int(self.seg_end[-1])
        first = elem_of(span.begin)
        last = elem_of(span.begin + span.n)
        return first, last - first


@dataclass
class NoiseSpec:
    """Stochastic-rounding noise: a float64 device tensor `u` (parity with np.random.random) or the
    device generator keyed by (seed, stream)."""
    u: Optional[torch.Tensor] = None
    seed: int = 0
    stream: int = 0
    resolution: int = 53      # device generator: 53 (numpy's construction, two words per element) or 32 (throughput mode)

    def c(self):
        if self.resolution not in (53, 32):
            raise ValueError("noise resolution must be 53 or 32")
        return _cabi.Noise(self.u.data_ptr() if self.u is not None else None, self.seed, self.stream,
                           _cabi.NOISE_32 if self.resolution == 32 else _cabi.NOISE_53, 0)


def _iter32(it):
    """The round counter is 4 big-endian bytes of the AES input (jzf_flashe.py:304: iter_index.to_bytes(4, 'big')
    raises OverflowError outside [0, 2^32)); same here instead of wrapping silently."""
    it = int(it)
    if not 0 <= it < (1 << 32):
        raise OverflowError("iter_index %d does not fit 4 bytes" % it)
    return it


def _i32(xs):
    return (C.c_int32 * max(1, len(xs)))(*[int(x) for x in xs])


class PeerSlice(object):
    """A range of a peer buffer (memory of ANOTHER process's GPU, mapped by flashe_peer_open): accepted wherever an
    `out=` float64 / word tensor is, so that a kernel stores its result straight into the owner's vector over
    NVLink.  Only the address and the size are known here — torch never sees the memory."""

    def __init__(self, address, nbytes):
        self.address, self.nbytes = int(address), int(nbytes)

    def data_ptr(self):
        return self.address


class DeviceContext(object):
    """One (key, int_bits, device) context = flashe_ctx."""

    def __init__(self, seed: bytes, int_bits: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("flashe_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _cabi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("flashe_b200 contexts live on CUDA devices")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)  # make sure the primary context exists
        self.int_bits = int(int_bits)
        seed = bytes(seed)
        h = C.c_void_p()
        buf = (C.c_uint8 * len(seed)).from_buffer_copy(seed)
        _cabi.check(self.lib.flashe_ctx_create(buf, len(seed), self.int_bits, self.device.index, C.byref(h)))
        self._h = h
        self.word_bytes = self.lib.flashe_word_bytes(self.int_bits)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.flashe_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ CUDA graphs
    def capture(self, fn, warmup=2):
        """Record everything `fn()` enqueues (calls of this or other contexts on torch's current stream) into a
        CUDA graph and return a zero-argument callable that replays it: one graph launch instead of one kernel
        launch plus host-side argument marshalling per call — what small vectors (BASELINE configs 1-2, where a
        round is three ~10-30 us kernels) are bound by.  `fn` must write into pre-allocated `out=` tensors and use
        fixed arguments (iteration index, index lists): they are baked into the graph.  Layer tables with more
        than 48 layers are uploaded with a synchronising copy and cannot be captured."""
        for _ in range(max(1, warmup)):
            fn()
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        self._graphs = getattr(self, "_graphs", []) + [graph]     # keep alive as long as the context
        return graph.replay

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty_words(self, n, rows=None):
        shape = [n] if rows is None else [rows, n]
        if self.word_bytes == 4:
            return torch.empty(shape, dtype=torch.uint32, device=self.device)
        if self.word_bytes == 8:
            return torch.empty(shape, dtype=torch.uint64, device=self.device)
        return torch.empty(shape + [2], dtype=torch.uint64, device=self.device)

    def zeros_words(self, n, rows=None):
        shape = [n] if rows is None else [rows, n]
        if self.word_bytes == 4:
            return torch.zeros(shape, dtype=torch.int32, device=self.device).view(torch.uint32)
        if self.word_bytes == 8:
            return torch.zeros(shape, dtype=torch.int64, device=self.device).view(torch.uint64)
        return torch.zeros(shape + [2], dtype=torch.int64, device=self.device).view(torch.uint64)

    def words_from_ints(self, values):
        """Sequence / object array of Python ints -> device word tensor (reduced mod 2^int_bits)."""
        import numpy as np
        v = np.asarray(values, dtype=object).reshape(-1) & ((1 << self.int_bits) - 1)
        if self.word_bytes == 4:
            host = torch.from_numpy(v.astype(np.uint32))
        elif self.word_bytes == 8:
            host = torch.from_numpy(v.astype(np.uint64))
        else:
            m64 = (1 << 64) - 1
            host = torch.from_numpy(np.stack([(v & m64).astype(np.uint64), (v >> 64).astype(np.uint64)], axis=1))
        return host.to(self.device)

    def ints_from_words(self, t):
        """Device word tensor -> object array of Python ints."""
        a = t.cpu().numpy()
        if self.word_bytes == 16:
            a = a.reshape(-1, 2)
            return a[:, 0].astype(object) | (a[:, 1].astype(object) << 64)
        return a.reshape(-1).astype(object)

    def _check_words(self, t, n, what):
        if t.device != self.device:
            raise ValueError("%s must live on %s" % (what, self.device))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % what)
        if t.numel() * t.element_size() != n * self.word_bytes:
            raise ValueError("%s must hold %d words of %d bytes" % (what, n, self.word_bytes))
        return t

    def _check(self, t, dtype, n, what):
        if isinstance(t, PeerSlice):
            if t.nbytes != n * torch.empty(0, dtype=dtype).element_size() or t.address % 16:
                raise ValueError("%s: peer slice must hold %d elements of %s and be 16-byte aligned" % (what, n, dtype))
            return t
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous() or t.numel() != n:
            raise ValueError("%s must be a contiguous %s tensor of %d elements on %s" % (what, dtype, n, self.device))
        return t

    # ------------------------------------------------------------------ primitives
    def prp_block(self, block16: bytes) -> bytes:
        out = (C.c_uint8 * 16)()
        _cabi.check(self.lib.flashe_prp_block(self._h, (C.c_uint8 * 16).from_buffer_copy(bytes(block16)), out, self._stream()))
        return bytes(out)

    def masks(self, it, prf_idx, sign, span: VectorSpan, out=None):
        out = self.empty_words(span.n) if out is None else self._check_words(out, span.n, "out")
        _cabi.check(self.lib.flashe_masks(self._h, _iter32(it), _i32(prf_idx), _i32(sign), len(prf_idx),
                                          C.byref(span.c()), out.data_ptr(), self._stream()))
        return out

    def precompute(self, iter_from, n_rounds, prf_idx, sign, span: VectorSpan, out=None):
        """Combined keystreams of rounds iter_from .. iter_from+n_rounds-1 -> words [n_rounds, span.n]."""
        out = self.empty_words(span.n, rows=n_rounds) if out is None else self._check_words(out, n_rounds * span.n, "out")
        _cabi.check(self.lib.flashe_precompute(self._h, _iter32(iter_from), n_rounds, _i32(prf_idx), _i32(sign),
                                               len(prf_idx), C.byref(span.c()), out.data_ptr(), span.n, self._stream()))
        return out

    def apply_masks(self, it, prf_idx, sign, words, span: VectorSpan, out=None):
        self._check_words(words, span.n, "words")
        out = torch.empty_like(words) if out is None else self._check_words(out, span.n, "out")
        _cabi.check(self.lib.flashe_apply_masks(self._h, _iter32(it), _i32(prf_idx), _i32(sign), len(prf_idx),
                                                C.byref(span.c()), words.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def encrypt(self, it, idx, scheme, q, span: VectorSpan, out=None):
        self._check_words(q, span.n, "q")
        out = torch.empty_like(q) if out is None else self._check_words(out, span.n, "out")
        _cabi.check(self.lib.flashe_encrypt(self._h, _iter32(it), idx, scheme, C.byref(span.c()), q.data_ptr(),
                                            out.data_ptr(), self._stream()))
        return out

    def decrypt(self, it, add_idx, minus_idx, agg, span: VectorSpan, out=None):
        self._check_words(agg, span.n, "agg")
        out = torch.empty_like(agg) if out is None else self._check_words(out, span.n, "out")
        _cabi.check(self.lib.flashe_decrypt(self._h, _iter32(it), _i32(add_idx), len(add_idx), _i32(minus_idx),
                                            len(minus_idx), C.byref(span.c()), agg.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def add_premasked(self, words, mask, sign=1, out=None):
        n = words.numel() * words.element_size() // self.word_bytes
        self._check_words(words, n, "words"); self._check_words(mask, n, "mask")
        out = torch.empty_like(words) if out is None else self._check_words(out, n, "out")
        _cabi.check(self.lib.flashe_add_premasked(self._h, words.data_ptr(), mask.data_ptr(), sign, n, out.data_ptr(), self._stream()))
        return out

    def encode(self, x, codec: CodecSpec, noise: NoiseSpec, span: VectorSpan, out=None):
        self._check(x, torch.float32, span.n, "x")
        if noise.u is not None:
            self._check(noise.u, torch.float64, span.n, "noise.u")
        out = torch.empty(span.n, dtype=torch.uint32, device=self.device) if out is None else out
        cc, nc = codec.c(span.total_len), noise.c()
        _cabi.check(self.lib.flashe_encode(self._h, C.byref(span.c()), x.data_ptr(), C.byref(cc), C.byref(nc),
                                           out.data_ptr(), self._stream()))
        return out

    def encode_encrypt(self, it, idx, scheme, x, codec: CodecSpec, noise: NoiseSpec, span: VectorSpan, out=None, q_out=None):
        """Fused encode + encrypt.  With codec.batch_lane_bits (int_bits 65..128) the span counts WORDS and x
        holds the span's ELEMENTS (the flattened layers): encode -> lane pack -> mask in one launch."""
        ne = codec.span_elements(self.int_bits, span)[1]
        self._check(x, torch.float32, ne, "x")
        if noise.u is not None:
            self._check(noise.u, torch.float64, ne, "noise.u")
        out = self.empty_words(span.n) if out is None else self._check_words(out, span.n, "out")
        cc, nc = codec.c(span.total_len), noise.c()
        _cabi.check(self.lib.flashe_encode_encrypt(self._h, _iter32(it), idx, scheme, C.byref(span.c()), x.data_ptr(),
                                                   C.byref(cc), C.byref(nc), out.data_ptr(),
                                                   q_out.data_ptr() if q_out is not None else None, self._stream()))
        return out

    def encode_encrypt_batch(self, it, idx0, scheme, x, codec: CodecSpec, noise: NoiseSpec, span: VectorSpan, out=None,
                             share_streams=False):
        """x: float32 [n_clients, span.n]; returns words [n_clients, span.n]."""
        ne = codec.span_elements(self.int_bits, span)[1]
        if x.dim() != 2 or x.shape[1] != ne or x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
            raise ValueError("x must be a contiguous float32 [n_clients, %d] tensor on %s" % (ne, self.device))
        n = x.shape[0]
        if noise.u is not None:
            self._check(noise.u, torch.float64, n * ne, "noise.u")
        out = self.empty_words(span.n, rows=n) if out is None else self._check_words(out, n * span.n, "out")
        cc = codec.c(span.total_len)
        # one launch takes at most FLASHE_MAX_STREAMS - 1 clients (their PRF index terms live in the kernel's
        # parameter block): more clients go out in consecutive launches, same outputs
        step = _cabi.MAX_STREAMS - 1
        for c0 in range(0, n, step):
            c1 = min(n, c0 + step)
            ns = NoiseSpec(u=None if noise.u is None else noise.u[c0 * ne:c1 * ne], seed=noise.seed, stream=noise.stream + c0,
                           resolution=noise.resolution)
            nc = ns.c()
            _cabi.check(self.lib.flashe_encode_encrypt_batch(self._h, _iter32(it), idx0 + c0, c1 - c0, scheme, C.byref(span.c()),
                                                             x[c0:c1].data_ptr(), ne, C.byref(cc), C.byref(nc), ne,
                                                             out[c0:c1].data_ptr(), span.n, 1 if share_streams else 0, self._stream()))
        return out

    def encode_add_premasked(self, x, codec: CodecSpec, noise: NoiseSpec, mask, span: VectorSpan, out=None):
        self._check(x, torch.float32, span.n, "x")
        self._check_words(mask, span.n, "mask")
        out = self.empty_words(span.n) if out is None else self._check_words(out, span.n, "out")
        cc, nc = codec.c(span.total_len), noise.c()
        _cabi.check(self.lib.flashe_encode_add_premasked(self._h, C.byref(span.c()), x.data_ptr(), C.byref(cc), C.byref(nc),
                                                         mask.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def encode_add_premasked_batch(self, x, codec: CodecSpec, noise: NoiseSpec, masks, span: VectorSpan, out=None):
        """All clients' online step in one launch: x float32 [n, count], masks / out words [n, count]."""
        if x.dim() != 2 or x.shape[1] != span.n or x.dtype != torch.float32 or not x.is_contiguous() or x.device != self.device:
            raise ValueError("x must be a contiguous float32 [n_clients, count] tensor on %s" % self.device)
        n = x.shape[0]
        self._check_words(masks, n * span.n, "masks")
        if noise.u is not None:
            self._check(noise.u, torch.float64, n * span.n, "noise.u")
        out = self.empty_words(span.n, rows=n) if out is None else self._check_words(out, n * span.n, "out")
        cc, nc = codec.c(span.total_len), noise.c()
        _cabi.check(self.lib.flashe_encode_add_premasked_batch(self._h, C.byref(span.c()), n, x.data_ptr(), span.n, C.byref(cc), C.byref(nc),
                                                               span.n, masks.data_ptr(), span.n, out.data_ptr(), span.n, self._stream()))
        return out

    def aggregate(self, cts, mode=AGG_ELEMENTWISE, carry_in=0, out=None, carry_out=None):
        """cts: words [n, L] (+[,2]); returns words [L]."""
        n = cts.shape[0]
        L = cts.shape[1]
        self._check_words(cts, n * L, "cts")
        out = self.empty_words(L) if out is None else self._check_words(out, L, "out")
        _cabi.check(self.lib.flashe_aggregate(self._h, cts.data_ptr(), L, n, L, mode, carry_in, out.data_ptr(),
                                              carry_out.data_ptr() if carry_out is not None else None, self._stream()))
        return out

    def aggregate_carry_fixup(self, out, carry_in):
        n = out.numel() * out.element_size() // self.word_bytes
        _cabi.check(self.lib.flashe_aggregate_carry_fixup(self._h, out.data_ptr(), n, carry_in, self._stream()))
        return out

    def decode(self, v, codec: CodecSpec, span: VectorSpan, out=None):
        self._check_words(v, span.n, "v")
        out = torch.empty(span.n, dtype=torch.float64, device=self.device) if out is None else out
        cc = codec.c(span.total_len)
        _cabi.check(self.lib.flashe_decode(self._h, C.byref(span.c()), v.data_ptr(), C.byref(cc), out.data_ptr(), self._stream()))
        return out

    def decrypt_decode(self, it, add_idx, minus_idx, agg, codec: CodecSpec, span: VectorSpan, out=None, p_out=None):
        """Fused decrypt + decode.  With codec.batch_lane_bits the aggregate holds lane-batched words and `out`
        one float64 per ELEMENT of the span (unmask -> unbatch -> decode in one launch)."""
        self._check_words(agg, span.n, "agg")
        ne = codec.span_elements(self.int_bits, span)[1]
        out = torch.empty(ne, dtype=torch.float64, device=self.device) if out is None else self._check(out, torch.float64, ne, "out")
        cc = codec.c(span.total_len)
        _cabi.check(self.lib.flashe_decrypt_decode(self._h, _iter32(it), _i32(add_idx), len(add_idx), _i32(minus_idx),
                                                   len(minus_idx), C.byref(span.c()), agg.data_ptr(), C.byref(cc),
                                                   out.data_ptr(), p_out.data_ptr() if p_out is not None else None,
                                                   self._stream()))
        return out

    # ------------------------------------------------------------------ peer buffers (multi-GPU gather, sharding.PeerGather)
    def peer_alloc(self, nbytes):
        """-> (address, 64-byte handle): device memory of this GPU that other processes can map (flashe_peer_alloc)."""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        _cabi.check(self.lib.flashe_peer_alloc(self._h, int(nbytes), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def peer_open(self, handle: bytes):
        ptr = C.c_void_p()
        _cabi.check(self.lib.flashe_peer_open(self._h, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(ptr)))
        return ptr.value

    def peer_close(self, address):
        _cabi.check(self.lib.flashe_peer_close(self._h, C.c_void_p(address)))

    def peer_free(self, address):
        _cabi.check(self.lib.flashe_peer_free(self._h, C.c_void_p(address)))

    def rng_uniform(self, seed, stream, begin, count, out=None, resolution=53):
        out = torch.empty(count, dtype=torch.float64, device=self.device) if out is None else out
        _cabi.check(self.lib.flashe_rng_uniform(self._h, seed, stream, _cabi.NOISE_32 if resolution == 32 else _cabi.NOISE_53,
                                                begin, count, out.data_ptr(), self._stream()))
        return out

    def batch_pack(self, q, element_bits, factor):
        self._check(q, torch.uint32, q.numel(), "q")
        bs = self.int_bits // (element_bits + factor)
        nw = (q.numel() + bs - 1) // bs
        out = self.empty_words(nw)
        _cabi.check(self.lib.flashe_batch_pack(self._h, q.data_ptr(), q.numel(), element_bits, factor, out.data_ptr(), self._stream()))
        return out

    def batch_unpack(self, words, element_bits, factor):
        nw = words.shape[0]
        self._check_words(words, nw, "words")
        bs = self.int_bits // (element_bits + factor)
        out = torch.empty(nw * bs, dtype=torch.uint32, device=self.device)
        _cabi.check(self.lib.flashe_batch_unpack(self._h, words.data_ptr(), nw, element_bits, factor, out.data_ptr(), self._stream()))
        return out

    def batch_layout(self, seg_end, element_bits, factor):
        """word_end[s] of the per-layer lane batching (each layer padded to a multiple of batch_size)."""
        n = len(seg_end)
        ends = (C.c_uint64 * n)(*[int(e) for e in seg_end])
        out = (C.c_uint64 * n)()
        _cabi.check(self.lib.flashe_batch_layout(self.int_bits, element_bits, factor, ends, n, out))
        return [int(x) for x in out]

    def batch_pack_layers(self, q, seg_end, element_bits, factor):
        """Whole quantised model (uint32 [total], layers ending at seg_end) -> words [sum ceil(size/bs)]."""
        self._check(q, torch.uint32, int(seg_end[-1]), "q")
        n = len(seg_end)
        ends = (C.c_uint64 * n)(*[int(e) for e in seg_end])
        out = self.empty_words(self.batch_layout(seg_end, element_bits, factor)[-1])
        _cabi.check(self.lib.flashe_batch_pack_layers(self._h, q.data_ptr(), ends, n, element_bits, factor, out.data_ptr(), self._stream()))
        return out

    def batch_unpack_layers(self, words, seg_end, element_bits, factor):
        """Inverse of batch_pack_layers: words -> uint32 [total] (padding lanes dropped)."""
        n = len(seg_end)
        nw = self.batch_layout(seg_end, element_bits, factor)[-1]
        self._check_words(words, nw, "words")
        ends = (C.c_uint64 * n)(*[int(e) for e in seg_end])
        out = torch.empty(int(seg_end[-1]), dtype=torch.uint32, device=self.device)
        _cabi.check(self.lib.flashe_batch_unpack_layers(self._h, words.data_ptr(), ends, n, element_bits, factor, out.data_ptr(), self._stream()))
        return out

    def sparse_expand(self, compact, index, total, zero: int):
        k = index.numel()
        self._check(index, torch.int64, k, "index")
        self._check_words(compact, k, "compact")
        out = self.empty_words(total)
        z = int(zero).to_bytes(self.word_bytes, "little")
        zb = (C.c_uint8 * self.word_bytes).from_buffer_copy(z)
        _cabi.check(self.lib.flashe_sparse_expand(self._h, compact.data_ptr(), index.data_ptr(), k, total, zb, out.data_ptr(), self._stream()))
        return out

    def _check_index_list(self, ix, total):
        """The tiled sparse kernels rely on sorted unique index lists inside [0, total) (the reference sends
        `sorted(locations)`, jzf_aggregator.py:605-618; its fancy indexing raises IndexError outside the range)."""
        if ix.numel():
            bad = bool((ix[0] < 0) | (ix[-1] >= total)) or (ix.numel() > 1 and not bool((ix[1:] > ix[:-1]).all()))
            if bad:
                raise IndexError("sparse index list must be sorted, unique and inside [0, %d)" % total)

    def sparse_sum(self, compacts, indexes, total, zeros, out=None, validate=False):
        """Sum over clients of expand_to_dense(compact_c, index_c, zero_c) mod 2^b without the n dense
        vectors: every tile of the result is built in shared memory from the clients' runs of sorted indices
        (flashe_sparse_sum).  validate=True checks the lists on the device first (lists from other parties)."""
        n = len(compacts)
        if n < 1 or len(indexes) != n or len(zeros) != n:
            raise ValueError("compacts, indexes and zeros must have the same non-zero length")
        for a, ix in zip(compacts, indexes):
            self._check(ix, torch.int64, ix.numel(), "index")
            self._check_words(a, ix.numel(), "compact")
            if validate:
                self._check_index_list(ix, total)
        out = self.empty_words(total) if out is None else self._check_words(out, total, "out")
        cp = (C.c_void_p * n)(*[t.data_ptr() for t in compacts])
        ip = (C.c_void_p * n)(*[t.data_ptr() for t in indexes])
        ks = (C.c_uint64 * n)(*[t.numel() for t in indexes])
        zb = b"".join(int(z).to_bytes(self.word_bytes, "little") for z in zeros)
        zbuf = (C.c_uint8 * len(zb)).from_buffer_copy(zb)
        _cabi.check(self.lib.flashe_sparse_sum(self._h, cp, ip, ks, zbuf, n, total, out.data_ptr(), self._stream()))
        return out

    def sparse_apply_masks(self, it, prf_idx, sign, span: VectorSpan, index, dense, validate=True):
        """dense[index[i]] += sum_k sign_k F(it, prf_k)[i] over the compact positions.  `index` comes from
        other parties (the arbiter's mask lists): validate=True checks on the device that it is sorted,
        unique and inside [0, len(dense)) and raises IndexError otherwise, as the reference's fancy
        indexing would (jzf_flashe.py:333); the kernel itself skips out-of-range entries."""
        self._check(index, torch.int64, span.n, "index")
        total = dense.numel() * dense.element_size() // self.word_bytes
        if dense.device != self.device or not dense.is_contiguous():
            raise ValueError("dense must be contiguous on %s" % self.device)
        if validate and span.n:
            bad = bool((index[0] < 0) | (index[-1] >= total)) or (span.n > 1 and not bool((index[1:] > index[:-1]).all()))
            if bad:
                raise IndexError("sparse index list must be sorted, unique and inside [0, %d)" % total)
        _cabi.check(self.lib.flashe_sparse_apply_masks(self._h, _iter32(it), _i32(prf_idx), _i32(sign), len(prf_idx),
                                                       C.byref(span.c()), index.data_ptr(), dense.data_ptr(), total, self._stream()))
        return dense

    def sparse_apply_masks_batch(self, it, prf_idx, sign, n_jobs, index_lists, dense, validate=True):
        """dense[index_c[i]] += sign * F(it, prf_idx[c])[i] for every client c in ONE call (jzf_flashe.py:315-343): the
        masks of all clients are generated over their compact positions (chunk rule on each list's own length) and
        added tile by tile of `dense`.  Same result as one sparse_apply_masks per client."""
        n = len(index_lists)
        if n != len(prf_idx) or n < 1:
            raise ValueError("one prf index per index list")
        total = dense.numel() * dense.element_size() // self.word_bytes
        if dense.device != self.device or not dense.is_contiguous():
            raise ValueError("dense must be contiguous on %s" % self.device)
        for ix in index_lists:
            self._check(ix, torch.int64, ix.numel(), "index")
            if validate and ix.numel():
                bad = bool((ix[0] < 0) | (ix[-1] >= total)) or (ix.numel() > 1 and not bool((ix[1:] > ix[:-1]).all()))
                if bad:
                    raise IndexError("sparse index list must be sorted, unique and inside [0, %d)" % total)
        ptrs = (C.c_void_p * n)(*[ix.data_ptr() for ix in index_lists])
        ks = (C.c_uint64 * n)(*[ix.numel() for ix in index_lists])
        _cabi.check(self.lib.flashe_sparse_apply_masks_batch(self._h, _iter32(it), _i32(prf_idx), int(sign), n, ks, int(n_jobs), ptrs,
                                                             dense.data_ptr(), total, self._stream()))
        return dense

    def sparse_overlap(self, index_lists, total):
        n = len(index_lists)
        if n < 2:
            return []
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in index_lists])
        ks = (C.c_uint64 * n)(*[t.numel() for t in index_lists])
        out = (C.c_uint64 * (n - 1))()
        _cabi.check(self.lib.flashe_sparse_overlap(self._h, ptrs, ks, n, total, out, self._stream()))
        return [int(v) for v in out]

    # ------------------------------------------------------------------ wire format, sparsify, statistics
    @staticmethod
    def _word_bytes_of(t):
        """Storage width of one element of a word tensor: uint32/int32 -> 4, uint64/int64 -> 8, [L, 2] uint64 -> 16."""
        if t.dim() == 2 and t.shape[1] == 2 and t.element_size() == 8:
            return 16
        if t.element_size() not in (4, 8):
            raise ValueError("word tensors hold 4-, 8- or 16-byte elements")
        return t.element_size()

    def wire_pack(self, words, bits=None, out=None):
        """_to_bytes (jzf_weights.py:45-84) as a big-endian byte string: uint8 [ceil(L*bits/8)]."""
        bits = self.int_bits if bits is None else int(bits)
        wb = self._word_bytes_of(words)
        L = words.shape[0]
        if words.device != self.device or not words.is_contiguous():
            raise ValueError("words must be contiguous on %s" % self.device)
        nb = C.c_uint64()
        _cabi.check(self.lib.flashe_wire_nbytes(bits, L, C.byref(nb)))
        out = torch.empty(nb.value, dtype=torch.uint8, device=self.device) if out is None else self._check(out, torch.uint8, nb.value, "out")
        _cabi.check(self.lib.flashe_wire_pack(self._h, words.data_ptr(), wb, L, bits, out.data_ptr(), self._stream()))
        return out

    def wire_unpack(self, data, count, bits=None, word_bytes=None, out=None):
        """_from_bytes + reverse (jzf_weights.py:98-137, 224): uint8 stream -> word tensor of `count` elements."""
        bits = self.int_bits if bits is None else int(bits)
        wb = (4 if bits <= 32 else (8 if bits <= 64 else 16)) if word_bytes is None else int(word_bytes)
        nb = C.c_uint64()
        _cabi.check(self.lib.flashe_wire_nbytes(bits, count, C.byref(nb)))
        self._check(data, torch.uint8, nb.value, "data")
        if out is None:
            out = (torch.empty(count, dtype=torch.uint32, device=self.device) if wb == 4 else
                   torch.empty(count, dtype=torch.uint64, device=self.device) if wb == 8 else
                   torch.empty((count, 2), dtype=torch.uint64, device=self.device))
        _cabi.check(self.lib.flashe_wire_unpack(self._h, data.data_ptr(), count, bits, wb, out.data_ptr(), self._stream()))
        return out

    def topk_sparsify(self, x, seg_end, k, residual=None, residual_out=None):
        """Client.sparsify (jzf_aggregator.py:578-623) on a flat float32 vector of concatenated layers.
        Returns (values float32 [sum k], index int64 [sum k] global ascending, new residual float32 [L])."""
        L = x.numel()
        self._check(x, torch.float32, L, "x")
        if residual is not None:
            self._check(residual, torch.float32, L, "residual")
        residual_out = torch.empty(L, dtype=torch.float32, device=self.device) if residual_out is None else self._check(residual_out, torch.float32, L, "residual_out")
        ends = (C.c_uint64 * len(seg_end))(*[int(e) for e in seg_end])
        ks = (C.c_uint64 * len(k))(*[int(v) for v in k])
        K = sum(int(v) for v in k)
        values = torch.empty(K, dtype=torch.float32, device=self.device)
        index = torch.empty(K, dtype=torch.int64, device=self.device)
        _cabi.check(self.lib.flashe_topk_sparsify(self._h, x.data_ptr(), residual.data_ptr() if residual is not None else None, L,
                                                  ends, ks, len(seg_end), values.data_ptr(), index.data_ptr(),
                                                  residual_out.data_ptr(), self._stream()))
        return values, index, residual_out

    def segment_stats(self, w, seg_end, shift=None, out=None, inplace=False, order=_cabi.SUM_PAIRWISE):
        """unnormalize (jzf_quantize.py:549-564): w + shift per layer, then (np.mean, np.std) per layer,
        bit-exact in numpy's summation order (order = SUM_PAIRWISE for float64 ndarrays, SUM_SEQUENTIAL for
        object arrays of Python floats).  Returns (shifted w or None, stats float64 [nseg, 2] on the device)."""
        L = w.numel()
        self._check(w, torch.float64, L, "w")
        nseg = len(seg_end)
        ends = (C.c_uint64 * nseg)(*[int(e) for e in seg_end])
        sh = None if shift is None else (C.c_double * nseg)(*[float(v) for v in shift])
        if inplace:
            out = w
        stats = torch.empty((nseg, 2), dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.flashe_segment_stats(self._h, w.data_ptr(), out.data_ptr() if out is not None else None, L, ends, sh,
                                                  nseg, int(order), stats.data_ptr(), self._stream()))
        return out, stats


def launch_count():
    return int(_cabi.load().flashe_launch_count())
