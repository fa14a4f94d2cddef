"""Drives the reference's OWN code (the unmodified files staged under oracle/_ref/ by oracle/make_ref.py)
through one FLASHE round — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__ and bench.py (`--impl reference`, the `cpu_baseline` leg) import this module.
The round is driven the way the reference's own stand-alone benchmark drives it
(encrypt_test/final_big_table.ipynb:222-233, 408-411) and the way client / arbiter code does
(framework/homo/procedure/jzf_aggregator.py:722-739 quantize + encrypt, :421-430 element-wise sum, :883-899
decrypt + unquantize):

    q   = jzf_quantize._static_quantize_padding_asymmetric(x_c, alpha, 16)          per client
    ct  = FlasheCipher(int_bits).encrypt(q)          (idx = c, iter_index = t, Pool(N_JOBS) inside)
    agg = reduce((x + y) % 2^int_bits) over object arrays
    p   = FlasheCipher.decrypt(agg) after set_idx_list(survivors, "decrypt")
    y   = jzf_quantize._static_unquantize_padding_asymmetric(p, alpha, 16, n)

The reference takes N_JOBS from multiprocessing.cpu_count() (jzf_flashe.py:7); `n_jobs` overrides that
module global (the ciphertext format depends on it).  `inline_pool=True` swaps multiprocessing.Pool for an
in-process pool with the same starmap semantics (for use inside a process that already holds a CUDA
context, where forking is best avoided); timing runs use the real Pool, as the reference does.
"""
import importlib
import os
import sys
import time
from functools import reduce

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "federatedml", "secureprotol", "jzf_flashe.py"))


def load():
    """(jzf_flashe module, jzf_quantize module) imported from oracle/_ref under the pycryptodome shim."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
        for p in (REF_DIR, os.path.join(HERE, "ref_shim")):
            if p not in sys.path:
                sys.path.insert(0, p)
        fl = importlib.import_module("federatedml.secureprotol.jzf_flashe")
        qu = importlib.import_module("federatedml.secureprotol.jzf_quantize")
        if not os.path.abspath(fl.__file__).startswith(REF_DIR):
            raise RuntimeError("federatedml was imported from %s, not from oracle/_ref" % fl.__file__)
        _mods = (fl, qu, fl.Pool)
    return _mods[0], _mods[1]


class _InlinePool(object):
    def __init__(self, n):
        self.n = n

    def starmap(self, fn, inputs):
        return [fn(*a) for a in inputs]

    def close(self):
        pass

    def join(self):
        pass


def configure(n_jobs=None, inline_pool=False):
    fl, _ = load()
    fl.N_JOBS = int(n_jobs) if n_jobs else os.cpu_count()
    fl.Pool = _InlinePool if inline_pool else _mods[2]
    return fl.N_JOBS


def make_cipher(key, int_bits, idx, it, mask="double"):
    fl, _ = load()
    c = fl.FlasheCipher(int_bits, mask=mask)
    c.generate_prp_seed(bytes(key))
    c.idx = idx
    c.set_iter_index(it)
    return c


def run_round(key, int_bits, it, xs, alpha, element_bits=16, n_jobs=None, seeds=None, timings=None,
              inline_pool=False, survivors=None):
    """One round over the clients' float32 gradients `xs`.  seeds[c] (optional) seeds numpy's global
    generator before client c's quantize, so that the caller can reproduce the rounding noise
    (np.random.seed(s); np.random.random(L)).  Returns (qs, cts, agg, dec, out): lists / object arrays
    exactly as the reference produces them."""
    fl, qu = load()
    used_jobs = configure(n_jobs, inline_pool)
    n = len(xs)
    t = {"encode": 0.0, "encrypt": 0.0, "aggregate": 0.0, "decrypt": 0.0, "decode": 0.0, "n_jobs": used_jobs}
    qs, cts = [], []
    for c, x in enumerate(xs):
        if seeds is not None:
            np.random.seed(seeds[c])
        t0 = time.perf_counter()
        q = qu._static_quantize_padding_asymmetric(x, float(alpha), element_bits)
        t1 = time.perf_counter()
        ct = make_cipher(key, int_bits, c, it).encrypt(q)
        t2 = time.perf_counter()
        t["encode"] += t1 - t0
        t["encrypt"] += t2 - t1
        qs.append(q); cts.append(ct)
    alive = list(range(n)) if survivors is None else list(survivors)
    mod = 1 << int_bits
    t0 = time.perf_counter()
    agg = reduce(lambda a, b: (a + b) % mod, [cts[c] for c in alive])
    t1 = time.perf_counter()
    cipher = make_cipher(key, int_bits, 0, it)
    cipher.set_num_clients(n)
    cipher.set_idx_list(raw_idx_list=alive, mode="decrypt")
    dec = cipher.decrypt(agg)
    t2 = time.perf_counter()
    out = qu._static_unquantize_padding_asymmetric(dec, float(alpha), element_bits, len(alive))
    t3 = time.perf_counter()
    t["aggregate"] += t1 - t0
    t["decrypt"] += t2 - t1
    t["decode"] += t3 - t2
    if timings is not None:
        timings.update(t)
    return qs, cts, agg, dec, out
