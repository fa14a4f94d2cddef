"""Wire bit-packing of weight vectors (reference: federatedml/framework/jzf_weights.py:45-137 and
JZFTransferableWeights.compress / decompress, :155-231), on the GPU.

`_to_bytes` / `_from_bytes` keep the reference's names, argument order and return shapes (a Python int
and the element count; a list in REVERSED element order that decompress() then reverses, :224).  The
device-resident forms (`pack` / `unpack`) keep the stream as a uint8 tensor, which is what leaves the
GPU: int_bits/8 bytes per element instead of a word.

The reference raises for vectors no longer than one lcm(num_bits, 8) batch (`s` is still None at
jzf_weights.py:79); here every length packs, with the same formula
s = sum_j a[j] << ((L-1-j)*num_bits).
"""
import numpy as np
import torch

from .device import DeviceContext

_ctx_cache = {}


def _ctx(device=None):
    key = str(device)
    if key not in _ctx_cache:
        _ctx_cache[key] = DeviceContext(b"\x00", 32, device)       # packing holds no key
    return _ctx_cache[key]


def _word_bytes(bits):
    return 4 if bits <= 32 else (8 if bits <= 64 else 16)


def _to_device_words(ctx, flatten_array, num_bits):
    if isinstance(flatten_array, torch.Tensor):
        return flatten_array
    v = np.asarray(flatten_array, dtype=object).reshape(-1)
    wb = _word_bytes(num_bits)
    if wb == 4:
        host = torch.from_numpy(v.astype(np.uint32))
    elif wb == 8:
        host = torch.from_numpy(v.astype(np.uint64))
    else:
        m64 = (1 << 64) - 1
        host = torch.from_numpy(np.stack([(v & m64).astype(np.uint64), (v >> 64).astype(np.uint64)], axis=1))
    return host.to(ctx.device)


def pack(words, num_bits, device=None):
    """Device word tensor (or sequence of ints) -> uint8 device tensor, big-endian bytes of the wire integer."""
    ctx = _ctx(device)
    return ctx.wire_pack(_to_device_words(ctx, words, num_bits), num_bits)


def unpack(stream, length, num_bits, device=None):
    """uint8 device tensor -> device word tensor of `length` elements, in element order."""
    return _ctx(device).wire_unpack(stream, int(length), num_bits)


def _to_bytes(flatten_array, num_bits, device=None):
    """jzf_weights.py:45-84: returns (s, l) with s the packed Python integer."""
    l = len(flatten_array)
    if l == 0:
        return 0, 0
    stream = pack(flatten_array, int(num_bits), device)
    return int.from_bytes(stream.cpu().numpy().tobytes(), "big"), l


def _from_bytes(s, l, num_bits, device=None):
    """jzf_weights.py:98-137: list of the l fields of s, LAST element first (the caller reverses)."""
    ctx = _ctx(device)
    num_bits, l = int(num_bits), int(l)
    if l == 0:
        return []
    nbytes = (l * num_bits + 7) // 8
    data = torch.from_numpy(np.frombuffer(int(s).to_bytes(nbytes, "big"), dtype=np.uint8).copy()).to(ctx.device)
    words = ctx.wire_unpack(data, l, num_bits).cpu().numpy()
    if words.ndim == 2:
        vals = [int(lo) | (int(hi) << 64) for lo, hi in words]
    else:
        vals = [int(v) for v in words]
    vals.reverse()
    return vals
