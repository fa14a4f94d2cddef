#!/usr/bin/env python
"""Golden fixtures for the SURVEY §8 "next" rows (f1 wire packing, f2 top-k sparsify, f3 layer
statistics), produced by EXECUTING THE REFERENCE'S OWN SOURCE.

Run (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_golden_f.py

`framework/jzf_weights.py` and `framework/homo/procedure/jzf_aggregator.py` cannot be imported (they
import compress_pickle / modules that do not exist upstream), so the functions on the path
are lifted out of those files at run time with `ast` and compiled from the reference's text — nothing
is copied into this repository:

    _to_bytes, _from_bytes          framework/jzf_weights.py:45-137
    Client.sparsify                 framework/homo/procedure/jzf_aggregator.py:578-623
    QuantizingClient.unnormalize    imported normally from secureprotol/jzf_quantize.py:549-564

`_to_bytes` / `_from_bytes` were written for numpy 1.17 (requirements.txt:74), where a Python int
shifted by an np.int64 falls back to Python arithmetic; numpy 2 raises OverflowError there
(SURVEY P11).  The functions are therefore run with `np.lcm` returning an int subclass that keeps
Python-int semantics under arithmetic and offers the `.item()` the code calls (`_NpInt` below) — the
reference's statements execute unchanged.

Output: tests/golden/flashe_golden_f.npz.
"""
import ast
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(0, REF)


class _NpInt(int):
    """int with numpy-1.17-like surface: arithmetic stays closed, .item() exists."""

    def item(self):
        return int(self)


def _closed(name):
    def op(self, other):
        r = getattr(int, name)(int(self), int(other))
        return _NpInt(r) if isinstance(r, int) and not isinstance(r, bool) else r
    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__floordiv__", "__rfloordiv__",
           "__mod__", "__rmod__"):
    setattr(_NpInt, _n, _closed(_n))


class _NpProxy(object):
    """`np` as seen by the lifted functions: numpy, except lcm returns an _NpInt."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def lcm(a, b):
        return _NpInt(int(np.lcm(int(a), int(b))))


def lift(path, names, cls=None, extra=None):
    """Compile the named top-level functions (or methods of `cls`) of a reference file from its text."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in picked) == sorted(names), (path, names)
    mod = ast.Module(body=picked, type_ignores=[])
    ns = {"np": _NpProxy()}
    ns.update(extra or {})
    exec(compile(mod, os.path.join(REF, path), "exec"), ns)
    return [ns[n] for n in names]


class _Log(object):
    def info(self, *a, **k):
        pass

    debug = warning = info


OUT = {}
MANIFEST = {"cases": []}


def put(name, arr):
    assert name not in OUT, name
    OUT[name] = arr


def int_to_bytes(s, nbytes):
    return np.frombuffer(int(s).to_bytes(nbytes, "big"), dtype=np.uint8).copy()


def case_wire(to_bytes, from_bytes):
    rng = np.random.RandomState(77)
    cases = [(20, 1), (20, 2), (20, 3), (20, 7), (20, 1000), (20, 4099), (32, 5), (32, 2050), (24, 999), (8, 33),
             (13, 257), (26, 500), (27, 1234), (33, 100), (48, 301), (64, 77), (120, 50), (120, 3), (128, 9), (65, 40)]
    for i, (bits, L) in enumerate(cases):
        vals = [int.from_bytes(rng.bytes(16), "little") & ((1 << bits) - 1) for _ in range(L)]
        # edge values first / last so the boundary fields are exercised
        vals[0] = (1 << bits) - 1
        vals[-1] = (1 << bits) - 1 if L > 1 else vals[-1]
        name = "wire_%02d" % i
        try:
            s, l = to_bytes(np.array(vals, dtype=object), bits)
        except TypeError:
            # reference defect: with L <= lcm(bits, 8) / bits no full batch exists, `s` stays None
            # and int.from_bytes(None) raises (jzf_weights.py:54, 79).  Recorded, no data.
            assert L <= int(np.lcm(bits, 8)) // bits
            MANIFEST["cases"].append({"name": name, "bits": bits, "L": L, "ref_raises": True})
            continue
        assert l == L
        back = from_bytes(s, L, bits)
        back.reverse()                                    # decompress, jzf_weights.py:224
        assert [int(v) for v in back] == vals, (bits, L)
        nbytes = (bits * L + 7) // 8
        m64 = (1 << 64) - 1
        put(name + "_lo", np.array([v & m64 for v in vals], dtype=np.uint64))
        put(name + "_hi", np.array([v >> 64 for v in vals], dtype=np.uint64))
        put(name + "_bytes", int_to_bytes(s, nbytes))
        MANIFEST["cases"].append({"name": name, "bits": bits, "L": L})


class _W(object):
    def __init__(self, layers):
        self._weights = dict(layers)
        self.walking_order = sorted(self._weights.keys(), key=str)


def case_sparsify(sparsify):
    rng = np.random.RandomState(4242)
    shapes = {"a_conv": (5, 5, 8), "b_dense": (300, 17), "c_bias": (10,), "d_one": (1,), "e_big": (6000,)}
    for ci, sparsity in enumerate([0.01, 0.1, 0.5]):
        me = types.SimpleNamespace(remain_weights=None, _sparsity=sparsity, shape_dict_used_for_sparsification=None)
        keys = sorted(shapes.keys())
        for rnd in range(3):                              # three rounds: residual accumulates
            layers = {k: (rng.standard_normal(shapes[k]) * 0.1).astype(np.float32) for k in keys}
            w = _W({k: v.copy() for k, v in layers.items()})
            enc, le, bits, base = sparsify(me, w)
            name = "sp_%d_r%d" % (ci, rnd)
            put(name + "_x", np.concatenate([layers[k].reshape(-1) for k in keys]))
            put(name + "_values", np.concatenate([np.asarray(w._weights[k], dtype=np.float32).reshape(-1) for k in keys]))
            put(name + "_remain", np.concatenate([np.asarray(me.remain_weights[k], dtype=np.float32).reshape(-1) for k in keys]))
            put(name + "_locbytes", int_to_bytes(enc, (bits * le + 7) // 8))
            MANIFEST["cases"].append({"name": name, "sparsity": sparsity, "round": rnd, "le": int(le), "bits": int(bits),
                                      "base": int(base), "sizes": [int(np.prod(shapes[k])) for k in keys]})


def case_stats():
    import federatedml.secureprotol.jzf_quantize as ref_quant
    rng = np.random.RandomState(99)
    sizes = [1, 7, 4096, 4097, 9001, 3]
    qc = ref_quant.QuantizingClient(int_bits=32, from_arbiter=None, to_arbiter=None, batch=False, element_bits=16,
                                    padding=True, secure=True)
    layers = {"l%02d" % i: rng.standard_normal(n) * (0.05 * (i + 1)) + 0.01 * i for i, n in enumerate(sizes)}
    w = _W({k: v.copy() for k, v in layers.items()})
    qc.past_layer_mean_list = [0.125 * i - 0.2 for i in range(len(sizes))]
    qc.past_layer_std_list = [1.0] * len(sizes)
    shift = list(qc.past_layer_mean_list)
    qc.unnormalize(w)
    keys = w.walking_order
    put("stats_w", np.concatenate([layers[k] for k in keys]))
    put("stats_shift", np.array(shift, dtype=np.float64))
    put("stats_sizes", np.array(sizes, dtype=np.int64))
    put("stats_w_out", np.concatenate([w._weights[k] for k in keys]))
    put("stats_mean", np.array([float(v) for v in qc.past_layer_mean_list], dtype=np.float64))
    put("stats_std", np.array([float(v) for v in qc.past_layer_std_list], dtype=np.float64))
    # the un-batched mode holds OBJECT arrays of Python floats at this point (unquantize of an object array,
    # jzf_quantize.py:102-107): numpy then sums them left to right instead of pairwise
    qo = ref_quant.QuantizingClient(int_bits=32, from_arbiter=None, to_arbiter=None, batch=False, element_bits=16,
                                    padding=True, secure=True)
    wo = _W({k: v.copy().astype(object) for k, v in layers.items()})
    qo.past_layer_mean_list = list(shift)
    qo.past_layer_std_list = [1.0] * len(sizes)
    qo.unnormalize(wo)
    put("stats_obj_w_out", np.concatenate([np.asarray(wo._weights[k], dtype=np.float64) for k in keys]))
    put("stats_obj_mean", np.array([float(v) for v in qo.past_layer_mean_list], dtype=np.float64))
    put("stats_obj_std", np.array([float(v) for v in qo.past_layer_std_list], dtype=np.float64))


def case_model(flatten_weights, unflatten_weights):
    """A whole multi-layer model through the reference's QuantizingClient (quantize -> [n-client sum] ->
    unquantize -> unnormalize) and Client.flatten_weights / unflatten_weights, un-batched and batched.
    Inputs and every intermediate the reference produced are stored per layer, in walking order."""
    import federatedml.secureprotol.jzf_quantize as ref_quant
    rng = np.random.RandomState(2024)
    shapes = {"conv1": (3, 3, 4, 8), "dense1": (50, 17), "dense1_b": (17,), "out": (17, 3), "zzz": (1,)}
    n = 5
    for batch in (False, True):
        tag = "model_b" if batch else "model_u"
        qc = ref_quant.QuantizingClient(int_bits=120 if batch else 20, from_arbiter=None, to_arbiter=None, batch=batch,
                                        element_bits=16, padding=True, secure=True)
        qc.num_clients = n
        layers = {k: (rng.standard_normal(shapes[k]) * 0.05).astype(np.float32) for k in shapes}
        layers["zzz"] = np.zeros(1, dtype=np.float32)
        w = _W({k: v.copy() for k, v in layers.items()})
        keys = list(w.walking_order)
        qc.set_layer_size_list(w)
        qc.past_layer_std_list = [0.05, 0.07, 0.0, 0.031, 1.0]          # one zero std: alpha falls back to 0.1
        qc.past_layer_mean_list = [0.01, -0.02, 0.0, 0.003, 0.0]
        mean_in = list(qc.past_layer_mean_list)
        np.random.seed(777 + int(batch))
        qc.normalize(w)
        qc.quantize(w)
        put(tag + "_x", np.concatenate([layers[k].reshape(-1) for k in keys]))
        put(tag + "_std_in", np.array(qc.past_layer_std_list, dtype=np.float64))
        put(tag + "_mean_in", np.array(mean_in, dtype=np.float64))
        qflat = [int(v) for k in keys for v in np.asarray(w._weights[k], dtype=object).reshape(-1)]
        m64 = (1 << 64) - 1
        put(tag + "_q_lo", np.array([v & m64 for v in qflat], dtype=np.uint64))
        put(tag + "_q_hi", np.array([v >> 64 for v in qflat], dtype=np.uint64))
        put(tag + "_q_sizes", np.array([int(np.asarray(w._weights[k]).size) for k in keys], dtype=np.int64))
        # flatten (client side, jzf_aggregator.py:723) ... the n-client sum stands in for the server ...
        me = types.SimpleNamespace(shape_dict=None)
        flat = flatten_weights(me, w)
        only = flat.walking_order[0]
        put(tag + "_flat_len", np.array([len(flat._weights[only])], dtype=np.int64))
        assert [int(v) for v in flat._weights[only]] == qflat
        flat._weights[only] = flat._weights[only] * n                    # every client sent the same vector
        # ... unflatten + unquantize + unnormalize (client side, jzf_aggregator.py:896-904).  shape_dict holds the
        # shapes flatten saw: the word counts per layer in batched mode, no entry for the sentinel 'zzz'
        back = unflatten_weights(me, flat)
        qc.unquantize(back)
        qc.unnormalize(back)
        okeys = list(back.walking_order)
        put(tag + "_out", np.concatenate([np.asarray(back._weights[k], dtype=np.float64).reshape(-1) for k in okeys]))
        put(tag + "_out_is_object", np.array([int(np.asarray(back._weights[k]).dtype == object) for k in okeys], dtype=np.int64))
        put(tag + "_mean_out", np.array([float(v) for v in qc.past_layer_mean_list], dtype=np.float64))
        put(tag + "_std_out", np.array([float(v) for v in qc.past_layer_std_list], dtype=np.float64))
        MANIFEST["cases"].append({"name": tag, "batch": batch, "n_clients": n, "keys": keys, "out_keys": okeys,
                                  "shapes": [list(shapes[k]) for k in keys], "int_bits": 120 if batch else 20})


def main():
    to_bytes, from_bytes = lift("federatedml/framework/jzf_weights.py", ["_to_bytes", "_from_bytes"])
    case_wire(to_bytes, from_bytes)
    (sparsify,) = lift("federatedml/framework/homo/procedure/jzf_aggregator.py", ["sparsify"], cls="Client",
                       extra={"LOGGER": _Log(), "_to_bytes": to_bytes, "_from_bytes": from_bytes})
    case_sparsify(sparsify)
    case_stats()
    fw, uw = lift("federatedml/framework/homo/procedure/jzf_aggregator.py", ["flatten_weights", "unflatten_weights"], cls="Client")
    case_model(fw, uw)
    OUT["manifest"] = np.frombuffer(json.dumps(MANIFEST).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "flashe_golden_f.npz")
    np.savez_compressed(path, **OUT)
    print("wrote %s: %d arrays, %d bytes" % (path, len(OUT), os.path.getsize(path)))


if __name__ == "__main__":
    main()
