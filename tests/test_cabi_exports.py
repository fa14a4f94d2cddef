"""CPU-side checks of the drop-in boundary: the library builds/loads and exports every symbol that
include/flashe_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "flashe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flashe_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from flashe_b200 import build
    return build.build_library()


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("flashe_ctx_create", "flashe_encrypt", "flashe_decrypt", "flashe_masks", "flashe_encode_encrypt",
                 "flashe_aggregate", "flashe_decrypt_decode", "flashe_sparse_expand", "flashe_batch_pack"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_signatures_cover_header(lib_path):
    from flashe_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_functions()
    lib = _cabi.load()
    assert lib.flashe_abi_version() == 2
    assert lib.flashe_word_bytes(20) == 4 and lib.flashe_word_bytes(64) == 8 and lib.flashe_word_bytes(120) == 16
    assert lib.flashe_word_bytes(0) < 0 and b"int_bits" in lib.flashe_last_error()


def test_library_is_sm100a_only(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_product_never_imports_oracle():
    # a product path that routes through the oracle would void every parity claim
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flashe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libflashe_oracle" not in src, f


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flashe_b200 import DeviceContext
    with pytest.raises(RuntimeError):
        DeviceContext(bytes(32), 20)
    from flashe_b200.secureprotol import FlasheCipher
    c = FlasheCipher(20)
    with pytest.raises(RuntimeError):
        c.generate_prp_seed(bytes(32))
