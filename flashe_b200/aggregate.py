"""The arbiter's arithmetic for the FLASHE path (reference:
federatedml/framework/homo/procedure/jzf_aggregator.py:150-165, 404-430 and
jzf_flashe_block.py:89-117), on the GPU.  Messaging, retries and partitioning around it are FATE
plumbing and out of scope.

Arrays are 1-D numpy object arrays of Python ints (what the reference's JZFOrderDictWeights hold) or
torch CUDA word tensors (flashe_b200.device layout); the latter stay on the device."""
import numpy as np
import torch

from .device import AGG_ELEMENTWISE, AGG_PACKED, DeviceContext
from .weights import _to_bytes

_ctx_cache = {}


def _ctx(int_bits, device=None):
    key = (int_bits, str(device))
    if key not in _ctx_cache:
        _ctx_cache[key] = DeviceContext(b"\x00", int_bits, device)   # the server holds no key
    return _ctx_cache[key]


def _stack(ctx, models):
    if isinstance(models, torch.Tensor):
        return models, True
    if isinstance(models[0], torch.Tensor):
        return torch.stack([m.view(torch.int64 if ctx.word_bytes > 4 else torch.int32) for m in models]), True
    rows = [ctx.words_from_ints(m) for m in models]
    return torch.stack([r.view(torch.int64 if ctx.word_bytes > 4 else torch.int32) for r in rows]), False


def aggregate(models, int_bits, is_compressed=False, device=None):
    """total = reduce(lambda x, y: (x + y) % mod, models) — jzf_aggregator.py:404-430.

    is_compressed=False: element-wise mod 2^int_bits (the decompressed / sparse branch, :421-430).
    is_compressed=True : the clients' packed wire integers are added mod 2^(int_bits*L) (:406-419);
                         carries out of element j leak into element j-1 and are reproduced exactly.
    `models`: list of n vectors (object arrays or device word tensors) or one [n, L] word tensor."""
    ctx = _ctx(int_bits, device)
    cts, on_device = _stack(ctx, models)
    out = ctx.aggregate(cts, AGG_PACKED if is_compressed else AGG_ELEMENTWISE)
    return out if on_device else ctx.ints_from_words(out)


def expand_to_dense(models, masks, total, int_bits, device=None):
    """jzf_aggregator.py:150-165: each upload is the compact ciphertext followed by that client's
    plaintext quantised zero; scatter the compact part to `mask` and fill the rest with the zero."""
    ctx = _ctx(int_bits, device)
    out = []
    for a, ma in zip(models, masks):
        a = np.asarray(a, dtype=object)
        zero, compact = int(a[-1]), a[:-1]
        index = torch.as_tensor(np.asarray(ma, dtype=np.int64)).to(ctx.device)
        dense = ctx.sparse_expand(ctx.words_from_ints(compact), index, int(total), zero)
        out.append(ctx.ints_from_words(dense))
    return out


def aggregate_sparse(models, masks, total, int_bits, device=None):
    """expand_to_dense (jzf_aggregator.py:150-165) of every upload followed by the element-wise reduce
    (:421-430), fused: the n dense vectors are never built.  Same result as
    aggregate(expand_to_dense(models, masks, total, int_bits), int_bits)."""
    ctx = _ctx(int_bits, device)
    compacts, indexes, zeros = [], [], []
    for a, ma in zip(models, masks):
        a = np.asarray(a, dtype=object)
        zeros.append(int(a[-1]))
        compacts.append(ctx.words_from_ints(a[:-1]))
        indexes.append(torch.as_tensor(np.asarray(ma, dtype=np.int64)).to(ctx.device))
    return ctx.ints_from_words(ctx.sparse_sum(compacts, indexes, int(total), zeros, validate=True))   # (the masks come from the clients)


def dynamic_masking(masks, total, device=None):
    """jzf_flashe_block.py:89-117: single = 2*sum|mask|; double = 2*single - 2*sum_i |mask_i ∩ mask_{i+1}|;
    choose single when single <= double.  Returns the dict the arbiter broadcasts plus the costs."""
    ctx = _ctx(32, device)
    lists = [torch.as_tensor(np.asarray(m, dtype=np.int64)).to(ctx.device) for m in masks]
    single_cost = 2 * sum(int(t.numel()) for t in lists)
    double_cost = 2 * single_cost - 2 * sum(ctx.sparse_overlap(lists, int(total)))
    choice = "single" if single_cost <= double_cost else "double"
    return {"choice": choice, "masks": masks, "single_cost": single_cost, "double_cost": double_cost}


def flatten_weights(weights):
    """Client.flatten_weights (jzf_aggregator.py:625-650): the layers, in walking order, become ONE flat
    vector stored under the first key; returns (weights, shape_dict) — the reference keeps shape_dict on the
    client object; the sentinel layer 'zzz' has no shape entry (:637-641).  One concatenation instead of the
    reference's repeated np.append (which re-copies the growing vector per layer)."""
    shape_dict = {}
    keys = list(weights.walking_order)
    if not keys:
        return weights, shape_dict
    parts = [np.array([])]                       # the reference starts from an empty float64 array (:626-627)
    for k in keys:
        layer = weights._weights[k]
        if k != "zzz":
            shape_dict[k] = layer.shape
        parts.append(layer.flatten())
        del weights._weights[k]
    weights._weights[keys[0]] = np.concatenate(parts)
    weights.walking_order = sorted(weights._weights.keys(), key=str)
    return weights, shape_dict


def unflatten_weights(weights, shape_dict):
    """Client.unflatten_weights (jzf_aggregator.py:652-671): split the single flat vector back into the layers
    of shape_dict (views of the flat vector, like the reference's slices)."""
    only_key = None
    for k in weights.walking_order:
        only_key = k
        break
    flat = weights._weights[only_key]
    for k, shape in shape_dict.items():
        size = int(np.prod(shape))
        weights._weights[k] = flat[:size].reshape(shape)
        flat = flat[size:]
    weights.walking_order = sorted(weights._weights.keys(), key=str)
    return weights


def flatten_layers(layers):
    """Device form of flatten_weights: a list of per-layer CUDA tensors -> (flat tensor, seg_end), the segment
    table every kernel of the path takes (flashe_codec.seg_end, flashe_batch_pack_layers, flashe_segment_stats,
    flashe_topk_sparsify), so that a caller can hand over per-layer tensors."""
    flat = torch.cat([t.reshape(-1) for t in layers])
    ends, tot = [], 0
    for t in layers:
        tot += t.numel()
        ends.append(tot)
    return flat, ends


def unflatten_layers(flat, seg_end, shapes=None):
    """Inverse of flatten_layers: views of `flat` per layer (reshaped when `shapes` is given)."""
    out, b = [], 0
    for i, e in enumerate(seg_end):
        t = flat[b:e]
        out.append(t.reshape(shapes[i]) if shapes is not None else t)
        b = e
    return out


class SparsifyingClient(object):
    """The sparsification state and step of the reference's aggregator Client
    (jzf_aggregator.py:560-623): `_sparsity`, `remain_weights` (per-layer residuals carried across
    rounds), `shape_dict_used_for_sparsification`.  `weights` is the reference's
    JZFOrderDictWeights-shaped object (`walking_order`, `_weights` dict of float32 ndarrays).

    All layers are handed to the GPU as one flat vector with a segment table (no per-layer launches of
    argsort): radix-select top-k per layer, ordered compaction, residual update in the same pass."""

    def __init__(self, sparsity=1.0, device=None):
        self._sparsity = sparsity
        self.remain_weights = None
        self.shape_dict_used_for_sparsification = None
        self.device = device
        self._remain_dev = None          # flat float32 residual, kept on the device between rounds

    def sparsify(self, weights):
        """Returns (encoded_locations, le, bits, base) and replaces every layer of `weights` by its compact
        top-k values, exactly like jzf_aggregator.py:578-623."""
        ctx = _ctx(32, self.device)
        keys = list(weights.walking_order)
        shapes = {k: weights._weights[k].shape for k in keys}
        sizes = [int(np.prod(shapes[k])) for k in keys]
        ends = np.cumsum(sizes)
        ks = [max(1, int(np.floor(self._sparsity * np.int64(n)))) for n in sizes]       # :598
        flat = np.concatenate([np.asarray(weights._weights[k], dtype=np.float32).reshape(-1) for k in keys])
        values, index, self._remain_dev = ctx.topk_sparsify(torch.from_numpy(flat).to(ctx.device), ends, ks,
                                                            residual=self._remain_dev)
        v = values.cpu().numpy()
        rem = self._remain_dev.cpu().numpy()
        self.remain_weights = {}
        o = b = 0
        for k, n, kk in zip(keys, sizes, ks):
            weights._weights[k] = v[o:o + kk]
            self.remain_weights[k] = rem[b:b + n]
            o += kk
            b += n
        if self.shape_dict_used_for_sparsification is None:
            self.shape_dict_used_for_sparsification = shapes
        base = int(ends[-1]) if len(ends) else 0
        bits = base.bit_length()
        encoded_locations, le = _to_bytes(index.view(torch.uint64), bits, self.device)
        return encoded_locations, le, bits, base
