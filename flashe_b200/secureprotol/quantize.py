"""Encode / decode mirror of federatedml/secureprotol/jzf_quantize.py for the live (padding,
asymmetric) variant, running on the GPU through the C ABI.

Module-level functions keep the reference's names and argument order so call sites and tests read the
same; `QuantizingClient` keeps the reference's constructor and methods (jzf_quantize.py:336-564).

Stochastic rounding draws `np.random.random(size)` on the host in the same order the reference does
(one draw per layer, jzf_quantize.py:64), so a seeded `np.random.seed(s)` run is bit-identical to the
reference; the numbers are copied to the device and consumed by the encode kernel.  The tensor-native
API (flashe_b200.device) can instead use the on-device counter-based generator.
"""
import numpy as np
import torch

from ..device import CodecSpec, DeviceContext, NoiseSpec, VectorSpan
from .aciq import ACIQ

_codec_ctx = {}


def _ctx(int_bits=32, device=None):
    """Encode/decode do not use the key; one keyless context per (int_bits, device) serves them."""
    key = (int_bits, str(device))
    if key not in _codec_ctx:
        _codec_ctx[key] = DeviceContext(b"\x00", int_bits, device)
    return _codec_ctx[key]


def _static_quantize_padding_asymmetric(value, alpha, int_bits, device=None):
    """jzf_quantize.py:55-67.  value: float32 ndarray; alpha: Python float; int_bits = element bits.
    Returns an object array of Python ints, like the reference."""
    ctx = _ctx(32, device)
    flat = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
    u = np.random.random(flat.shape)                       # same global-RNG draw as the reference
    span = VectorSpan(total_len=flat.size, n_jobs=1)
    x_d = torch.from_numpy(flat).to(ctx.device)
    u_d = torch.from_numpy(u).to(ctx.device)
    q = ctx.encode(x_d, CodecSpec(alpha=float(alpha), element_bits=int_bits), NoiseSpec(u=u_d), span)
    return q.cpu().numpy().astype(object).reshape(np.shape(value))


def _static_unquantize_padding_asymmetric(value, alpha, int_bits, num_clients, device=None):
    """jzf_quantize.py:102-107 — float64 arithmetic on integers; returns an object array of Python
    floats for object input (as the reference's object-array arithmetic does), float64 otherwise."""
    v = np.asarray(value)
    ctx = _ctx(64, device)
    words = torch.from_numpy(v.astype(object).astype(np.uint64) if v.dtype == object else v.astype(np.uint64)).to(ctx.device)
    span = VectorSpan(total_len=words.numel(), n_jobs=1)
    out = ctx.decode(words.reshape(-1), CodecSpec(alpha=float(alpha), element_bits=int_bits, n_clients=int(num_clients)), span)
    out = out.cpu().numpy().reshape(v.shape)
    return out.astype(object) if v.dtype == object else out


def _static_batching_padding_asymmetric(array, int_bits, element_bits, factor, device=None):
    """jzf_quantize.py:162-185 — lanes of element_bits+factor bits, first element most significant."""
    ctx = _ctx(int_bits, device)
    q = torch.from_numpy(np.asarray(array).astype(object).astype(np.uint32)).to(ctx.device)
    w = ctx.batch_pack(q, element_bits, factor).cpu().numpy()
    return w[:, 0].astype(object) | (w[:, 1].astype(object) << 64)


def _static_unbatching_padding_asymmetric(array, int_bits, element_bits, factor, device=None):
    """jzf_quantize.py:234-251 — returns an integer ndarray of len(array)*batch_size lanes."""
    ctx = _ctx(int_bits, device)
    v = np.asarray(array).astype(object)
    m64 = (1 << 64) - 1
    words = torch.from_numpy(np.stack([(v & m64).astype(np.uint64), (v >> 64).astype(np.uint64)], axis=1)).to(ctx.device)
    return ctx.batch_unpack(words, element_bits, factor).cpu().numpy().astype(np.int64)


class QuantizingBase(object):
    def __init__(self, int_bits, batch, element_bits, secure):
        self.int_bits = int_bits
        self.num_clients = None
        self.iter = 0
        self.batch = batch
        self.r_max_list = None
        self.alpha_list = None
        self.shape_list = None
        self.secure = secure
        self.layer_size_list = None
        self.element_bits = element_bits

    def set_iter(self, iter):
        self.iter = iter


class QuantizingClient(QuantizingBase):
    """jzf_quantize.py:336-564.  `from_arbiter` / `to_arbiter` are the FATE transfer variables; only
    receive_num_clients() and the non-secure branch of send_layer_size_list() touch them."""

    def __init__(self, int_bits, from_arbiter, to_arbiter, batch, element_bits, padding, secure, device=None):
        super(QuantizingClient, self).__init__(int_bits, batch, element_bits, secure)
        self.from_arbiter = from_arbiter
        self.to_arbiter = to_arbiter
        self.padding = padding
        self.device = device
        self.expected_mean_for_first_round = 0.0
        self.expected_std_for_first_round = 1.0
        if secure:
            self.past_layer_mean_list = []
            self.past_layer_std_list = []

    def receive_num_clients(self):
        self.num_clients = self.from_arbiter.get(idx=0, suffix=(self.iter, 'num_clients'))
        return self.num_clients

    def _remember_layers(self, weights):
        self.layer_size_list = [weights._weights[k].size for k in weights.walking_order]
        for _ in self.layer_size_list:
            self.past_layer_mean_list.append(self.expected_mean_for_first_round)
            self.past_layer_std_list.append(self.expected_std_for_first_round)

    def send_layer_size_list(self, weights):   # guest (jzf_quantize.py:357-378)
        if not self.secure:
            raise NotImplementedError("secure=False is deprecated in the reference (jzf_quantize.py:414-423)")
        self._remember_layers(weights)

    def set_layer_size_list(self, weights):    # host (jzf_quantize.py:380-392)
        self._remember_layers(weights)

    # ------------------------------------------------------------------ whole-model helpers
    @staticmethod
    def _layers(weights):
        keys = list(weights.walking_order)
        arrays = [np.asarray(weights._weights[k]) for k in keys]
        sizes = [int(a.size) for a in arrays]
        return keys, arrays, sizes, [int(v) for v in np.cumsum(sizes)]

    def quantize(self, weights):
        """jzf_quantize.py:394-491: per layer alpha = ACIQ(element_bits)*std of that layer in the last
        global model (0 -> 0.1; the sparse sentinel layer 'zzz' uses 1.0), then encode (+ lane batch).

        The whole model goes to the device ONCE: the layers are concatenated (what flatten_weights does one
        step later, jzf_aggregator.py:625-650), their alphas ride in the segment table of flashe_encode, the
        rounding noise is drawn layer by layer from numpy's global generator in the reference's order
        (jzf_quantize.py:64), and batching packs every layer by itself through the same table
        (flashe_batch_pack_layers)."""
        if not self.secure or not self.padding:
            raise NotImplementedError("only secure=True, padding=True is live in the reference")
        aciq = ACIQ(self.element_bits)
        alpha_list = []
        for i, _ in enumerate(self.layer_size_list):
            alpha = aciq.get_alpha_gaus_direct(self.past_layer_std_list[i])
            if alpha == 0:
                alpha = 0.1
            alpha_list.append(alpha)
        self.r_max_list, self.alpha_list = [], []
        if self.batch:
            self.shape_list = []
        factor = int(np.ceil(np.log2(self.num_clients)))
        keys, arrays, sizes, ends = self._layers(weights)
        if not keys or ends[-1] == 0:
            return weights
        alphas = []
        for layer_cnt, k in enumerate(keys):
            if k == 'zzz':
                alpha = 1.0
            else:
                alpha = alpha_list[layer_cnt]
                self.r_max_list.append(alpha * self.num_clients)
                self.alpha_list.append(alpha)
            alphas.append(float(alpha))
            if self.batch:
                self.shape_list.append(arrays[layer_cnt].shape)
        ctx = _ctx(self.int_bits if self.batch else 32, self.device)
        x = np.concatenate([np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in arrays])
        u = np.concatenate([np.random.random(a.reshape(-1).shape) for a in arrays])    # the reference's draws, in its order
        live = [i for i, n in enumerate(sizes) if n]                                   # empty layers own no table row
        codec = CodecSpec(alpha=[alphas[i] for i in live], element_bits=self.element_bits, seg_end=[ends[i] for i in live])
        q = ctx.encode(torch.from_numpy(x).to(ctx.device), codec, NoiseSpec(u=torch.from_numpy(u).to(ctx.device)),
                       VectorSpan(total_len=ends[-1], n_jobs=1))
        if self.batch:
            wend = ctx.batch_layout(ends, self.element_bits, factor)
            words = ctx.batch_pack_layers(q, ends, self.element_bits, factor).cpu().numpy()
            ints = words[:, 0].astype(object) | (words[:, 1].astype(object) << 64)
            b = 0
            for k, e in zip(keys, wend):
                weights._weights[k] = ints[b:e]
                b = e
        else:
            ints = q.cpu().numpy().astype(object)
            b = 0
            for k, a, e in zip(keys, arrays, ends):
                weights._weights[k] = ints[b:e].reshape(a.shape)
                b = e
        return weights

    def unquantize(self, weights):
        """jzf_quantize.py:493-540, the whole model in one unbatch + one decode launch."""
        factor = int(np.ceil(np.log2(self.num_clients)))
        keys, arrays, sizes, ends = self._layers(weights)
        if not keys:
            return weights
        alphas = [float(self.alpha_list[i]) for i in range(len(keys))]
        is_obj = [a.dtype == object for a in arrays]
        if self.batch:
            shapes = [tuple(self.shape_list[i]) for i in range(len(keys))]
            esizes = [int(np.prod(sh)) for sh in shapes]
            eends = [int(v) for v in np.cumsum(esizes)]
            ctx = _ctx(self.int_bits, self.device)
            wend = ctx.batch_layout(eends, self.element_bits, factor)
            if wend != ends:
                raise ValueError("batched layers do not match shape_list")
            v = np.concatenate([a.reshape(-1).astype(object) for a in arrays])
            m64 = (1 << 64) - 1
            words = torch.from_numpy(np.stack([(v & m64).astype(np.uint64), (v >> 64).astype(np.uint64)], axis=1)).to(ctx.device)
            lanes = ctx.batch_unpack_layers(words, eends, self.element_bits, factor)
            is_obj = [False] * len(keys)      # the reference's unbatch returns an integer ndarray: float64 arithmetic follows
            vals, vends = (lanes.view(torch.int32).to(torch.int64) & 0xFFFFFFFF).view(torch.uint64), eends
        else:
            shapes = [a.shape for a in arrays]
            flat = np.concatenate([a.reshape(-1) for a in arrays])
            vals = torch.from_numpy(flat.astype(object).astype(np.uint64) if flat.dtype == object else flat.astype(np.uint64))
            vends = ends
        if not vends or vends[-1] == 0:
            return weights
        ctx64 = _ctx(64, self.device)
        live = [i for i in range(len(keys)) if (vends[i] - (vends[i - 1] if i else 0))]
        codec = CodecSpec(alpha=[alphas[i] for i in live], element_bits=self.element_bits, n_clients=int(self.num_clients),
                          seg_end=[vends[i] for i in live])
        out = ctx64.decode(vals.to(ctx64.device).reshape(-1), codec, VectorSpan(total_len=vends[-1], n_jobs=1)).cpu().numpy()
        b = 0
        for i, k in enumerate(keys):
            layer = out[b:vends[i]].reshape(shapes[i])
            weights._weights[k] = layer.astype(object) if is_obj[i] else layer
            b = vends[i]
        return weights

    def normalize(self, weights):
        """jzf_quantize.py:542-547 (host side, O(layers) scalars)."""
        for layer_cnt, k in enumerate(weights.walking_order):
            weights._weights[k] -= self.past_layer_mean_list[layer_cnt]
        return weights

    def unnormalize(self, weights):
        """jzf_quantize.py:549-564: add the mean back and refresh the per-layer mean / std that define the
        next round's alpha — one flashe_segment_stats call over the whole model, bit-exact with np.mean /
        np.std in the order numpy uses for the layer's dtype (pairwise for float64 ndarrays, left to right for
        the object arrays of Python floats the un-batched mode holds)."""
        from .._cabi import SUM_PAIRWISE, SUM_SEQUENTIAL
        keys, arrays, sizes, ends = self._layers(weights)
        if not keys:
            return weights
        ctx = _ctx(32, self.device)
        for want_obj in (False, True):
            sel = [i for i, a in enumerate(arrays) if (a.dtype == object) == want_obj]
            if not sel:
                continue
            flat = np.concatenate([arrays[i].reshape(-1).astype(np.float64) for i in sel]) if sel else np.empty(0)
            sends = [int(v) for v in np.cumsum([sizes[i] for i in sel])]
            shift = [float(self.past_layer_mean_list[i]) for i in sel]
            w_d = torch.from_numpy(flat).to(ctx.device)
            _, stats = ctx.segment_stats(w_d, sends, shift, inplace=True, order=SUM_SEQUENTIAL if want_obj else SUM_PAIRWISE)
            w_h, st = w_d.cpu().numpy(), stats.cpu().numpy()
            b = 0
            for j, i in enumerate(sel):
                layer = w_h[b:sends[j]].reshape(arrays[i].shape)
                weights._weights[keys[i]] = layer.astype(object) if want_obj else layer
                # np.mean of an object array is a Python float, of a float64 array an np.float64; np.std: np.float64
                self.past_layer_mean_list[i] = float(st[j, 0]) if want_obj else np.float64(st[j, 0])
                self.past_layer_std_list[i] = np.float64(st[j, 1])
                b = sends[j]
        return weights
