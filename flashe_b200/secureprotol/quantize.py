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

    def quantize(self, weights):
        """jzf_quantize.py:394-491: per layer alpha = ACIQ(element_bits)*std of that layer in the last
        global model (0 -> 0.1; the sparse sentinel layer 'zzz' uses 1.0), then encode (+ lane batch)."""
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
        layer_cnt = 0
        for k in weights.walking_order:
            if k == 'zzz':
                alpha = 1.0
            else:
                alpha = alpha_list[layer_cnt]
                self.r_max_list.append(alpha * self.num_clients)
                self.alpha_list.append(alpha)
            layer_weights = weights._weights[k]
            shape = layer_weights.shape
            elements = _static_quantize_padding_asymmetric(layer_weights.flatten(), float(alpha), self.element_bits, self.device)
            if self.batch:
                self.shape_list.append(shape)
                weights._weights[k] = _static_batching_padding_asymmetric(elements, self.int_bits, self.element_bits,
                                                                          factor, self.device)
            else:
                weights._weights[k] = elements.astype(object).reshape(shape)
            layer_cnt += 1
        return weights

    def unquantize(self, weights):
        """jzf_quantize.py:493-540."""
        factor = int(np.ceil(np.log2(self.num_clients)))
        layer_cnt = 0
        for k in weights.walking_order:
            alpha = self.alpha_list[layer_cnt]
            layer_weights = weights._weights[k]
            flat = layer_weights.flatten()
            if self.batch:
                shape = self.shape_list[layer_cnt]
                size = int(np.prod(shape))
                flat = _static_unbatching_padding_asymmetric(flat, self.int_bits, self.element_bits, factor, self.device)[:size]
            else:
                shape = layer_weights.shape
            ret = _static_unquantize_padding_asymmetric(flat, alpha, self.element_bits, self.num_clients, self.device)
            weights._weights[k] = ret.reshape(shape)
            layer_cnt += 1
        return weights

    def normalize(self, weights):
        """jzf_quantize.py:542-547 (host side, O(layers) scalars)."""
        for layer_cnt, k in enumerate(weights.walking_order):
            weights._weights[k] -= self.past_layer_mean_list[layer_cnt]
        return weights

    def unnormalize(self, weights):
        """jzf_quantize.py:549-564: add the mean back and refresh the per-layer mean / std that define
        the next round's alpha."""
        for layer_cnt, k in enumerate(weights.walking_order):
            weights._weights[k] += self.past_layer_mean_list[layer_cnt]
            self.past_layer_mean_list[layer_cnt] = np.mean(weights._weights[k])
            self.past_layer_std_list[layer_cnt] = np.std(weights._weights[k])
        return weights
