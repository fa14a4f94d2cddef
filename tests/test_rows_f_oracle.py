"""CPU: the numpy restatements of the SURVEY §8 "next" rows (oracle/oracle.py) against the fixtures
produced by executing the reference's own source (tests/golden/make_golden_f.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(HERE, "golden", "flashe_golden_f.npz"))
    return d, json.loads(bytes(d["manifest"]).decode())


def words_of(d, name, bits):
    lo, hi = d[name + "_lo"], d[name + "_hi"]
    if bits <= 32:
        return lo.astype(np.uint32)
    return lo if bits <= 64 else np.stack([lo, hi], axis=1)


def test_wire_pack_unpack_match_reference(gold):
    d, man = gold
    seen = 0
    for c in man["cases"]:
        if not c["name"].startswith("wire_") or c.get("ref_raises"):
            continue
        bits, L = c["bits"], c["L"]
        w = words_of(d, c["name"], bits)
        packed = O.wire_pack(w, bits)
        assert np.array_equal(packed, d[c["name"] + "_bytes"]), c
        back = O.wire_unpack(packed, L, bits)
        want = w.astype(np.uint64) if bits <= 64 else w
        assert np.array_equal(back, want), c
        seen += 1
    assert seen >= 15


def test_wire_short_inputs_where_reference_raises(gold):
    # L <= lcm(bits, 8)/bits: the reference raises (s stays None); the format is still well defined
    _, man = gold
    short = [c for c in man["cases"] if c.get("ref_raises")]
    assert short
    for c in short:
        vals = np.arange(1, c["L"] + 1, dtype=np.uint32)
        s = 0
        for v in vals:
            s = (s << c["bits"]) | int(v)
        nb = (c["bits"] * c["L"] + 7) // 8
        assert bytes(O.wire_pack(vals, c["bits"])) == s.to_bytes(nb, "big")


def sparsify_cases(man):
    return [c for c in man["cases"] if c["name"].startswith("sp_")]


def test_sparsify_matches_reference(gold):
    d, man = gold
    remain = {}
    for c in sparsify_cases(man):
        ci = c["name"].split("_")[1]
        if c["round"] == 0:
            remain[ci] = None
        sizes = c["sizes"]
        x = d[c["name"] + "_x"]
        offs = np.cumsum([0] + sizes)
        layers = [x[offs[i]:offs[i + 1]] for i in range(len(sizes))]
        vals, rem, loc, base = O.sparsify(layers, remain[ci], c["sparsity"])
        assert base == c["base"] and loc.shape[0] == c["le"] and base.bit_length() == c["bits"]
        assert np.array_equal(np.concatenate(vals).view(np.uint32), d[c["name"] + "_values"].view(np.uint32)), c
        assert np.array_equal(np.concatenate(rem).view(np.uint32), d[c["name"] + "_remain"].view(np.uint32)), c
        # the reference encodes the locations with _to_bytes(locations, base.bit_length())
        assert np.array_equal(O.wire_pack(loc.astype(np.uint64), c["bits"]), d[c["name"] + "_locbytes"]), c
        remain[ci] = rem


def test_unnormalize_stats_match_reference(gold):
    d, _ = gold
    sizes = d["stats_sizes"]
    ends = np.cumsum(sizes)
    w_out, stats = O.unnormalize_stats(d["stats_w"], ends, d["stats_shift"])
    assert np.array_equal(w_out.view(np.uint64), d["stats_w_out"].view(np.uint64))
    assert np.array_equal(stats[:, 0].view(np.uint64), d["stats_mean"].view(np.uint64))
    assert np.array_equal(stats[:, 1].view(np.uint64), d["stats_std"].view(np.uint64))
    # object arrays (Python floats): the reference's un-batched mode; numpy sums those left to right
    w_out, stats = O.unnormalize_stats(d["stats_w"], ends, d["stats_shift"], order="sequential")
    assert np.array_equal(w_out.view(np.uint64), d["stats_obj_w_out"].view(np.uint64))
    assert np.array_equal(stats[:, 0].view(np.uint64), d["stats_obj_mean"].view(np.uint64))
    assert np.array_equal(stats[:, 1].view(np.uint64), d["stats_obj_std"].view(np.uint64))
    assert not np.array_equal(d["stats_obj_std"].view(np.uint64), d["stats_std"].view(np.uint64))   # the order matters
