"""ctypes view of include/flashe_b200.h (the C ABI of libflashe_b200.so).

Fails loudly: if the library has not been built, or does not load, importing a symbol raises
RuntimeError — nothing in this package falls back to a CPU implementation."""
import ctypes as C
import os

from . import build as _build

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
SCHEME_SINGLE, SCHEME_DOUBLE = 0, 1
AGG_ELEMENTWISE, AGG_PACKED = 0, 1
SUM_PAIRWISE, SUM_SEQUENTIAL = 0, 1
NOISE_53, NOISE_32 = 0, 1
MAX_STREAMS = 128
ABI_VERSION = 2


class Span(C.Structure):
    _fields_ = [("total_len", C.c_uint64), ("begin", C.c_uint64), ("count", C.c_uint64),
                ("n_jobs", C.c_uint32), ("reserved", C.c_uint32)]


class Codec(C.Structure):
    _fields_ = [("element_bits", C.c_int32), ("n_clients", C.c_int32), ("nseg", C.c_int32),
                ("batch_lane_bits", C.c_int32), ("seg_end", C.POINTER(C.c_uint64)), ("alpha", C.POINTER(C.c_double))]


class Noise(C.Structure):
    _fields_ = [("u", C.c_void_p), ("rng_seed", C.c_uint64), ("rng_stream", C.c_uint64), ("resolution", C.c_int32), ("reserved", C.c_int32)]


class FlasheError(RuntimeError):
    def __init__(self, code, msg):
        super(FlasheError, self).__init__("libflashe_b200 error %d: %s" % (code, msg))
        self.code = code


_vp, _i32p, _u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
_u32, _u64, _int = C.c_uint32, C.c_uint64, C.c_int
_spanp, _codecp, _noisep = C.POINTER(Span), C.POINTER(Codec), C.POINTER(Noise)

# name -> (restype, argtypes); one entry per function declared in include/flashe_b200.h
SIGNATURES = {
    "flashe_abi_version": (_int, []),
    "flashe_last_error": (C.c_char_p, []),
    "flashe_word_bytes": (_int, [_int]),
    "flashe_ctx_create": (_int, [_u8p, C.c_size_t, _int, _int, C.POINTER(_vp)]),
    "flashe_ctx_destroy": (_int, [_vp]),
    "flashe_ctx_int_bits": (_int, [_vp]),
    "flashe_ctx_device": (_int, [_vp]),
    "flashe_prp_block": (_int, [_vp, _u8p, _u8p, _vp]),
    "flashe_masks": (_int, [_vp, _u32, _i32p, _i32p, _int, _spanp, _vp, _vp]),
    "flashe_precompute": (_int, [_vp, _u32, _int, _i32p, _i32p, _int, _spanp, _vp, _u64, _vp]),
    "flashe_apply_masks": (_int, [_vp, _u32, _i32p, _i32p, _int, _spanp, _vp, _vp, _vp]),
    "flashe_encrypt": (_int, [_vp, _u32, C.c_int32, _int, _spanp, _vp, _vp, _vp]),
    "flashe_decrypt": (_int, [_vp, _u32, _i32p, _int, _i32p, _int, _spanp, _vp, _vp, _vp]),
    "flashe_add_premasked": (_int, [_vp, _vp, _vp, _int, _u64, _vp, _vp]),
    "flashe_encode": (_int, [_vp, _spanp, _vp, _codecp, _noisep, _vp, _vp]),
    "flashe_encode_encrypt": (_int, [_vp, _u32, C.c_int32, _int, _spanp, _vp, _codecp, _noisep, _vp, _vp, _vp]),
    "flashe_encode_encrypt_batch": (_int, [_vp, _u32, C.c_int32, _int, _int, _spanp, _vp, _u64, _codecp, _noisep,
                                           _u64, _vp, _u64, _int, _vp]),
    "flashe_encode_add_premasked": (_int, [_vp, _spanp, _vp, _codecp, _noisep, _vp, _vp, _vp]),
    "flashe_encode_add_premasked_batch": (_int, [_vp, _spanp, _int, _vp, _u64, _codecp, _noisep, _u64, _vp, _u64, _vp, _u64, _vp]),
    "flashe_aggregate": (_int, [_vp, _vp, _u64, _int, _u64, _int, _u32, _vp, _vp, _vp]),
    "flashe_aggregate_carry_fixup": (_int, [_vp, _vp, _u64, _u32, _vp]),
    "flashe_decode": (_int, [_vp, _spanp, _vp, _codecp, _vp, _vp]),
    "flashe_decrypt_decode": (_int, [_vp, _u32, _i32p, _int, _i32p, _int, _spanp, _vp, _codecp, _vp, _vp, _vp]),
    "flashe_rng_uniform": (_int, [_vp, _u64, _u64, _int, _u64, _u64, _vp, _vp]),
    "flashe_batch_pack": (_int, [_vp, _vp, _u64, _int, _int, _vp, _vp]),
    "flashe_batch_unpack": (_int, [_vp, _vp, _u64, _int, _int, _vp, _vp]),
    "flashe_batch_layout": (_int, [_int, _int, _int, C.POINTER(_u64), _int, C.POINTER(_u64)]),
    "flashe_batch_pack_layers": (_int, [_vp, _vp, C.POINTER(_u64), _int, _int, _int, _vp, _vp]),
    "flashe_batch_unpack_layers": (_int, [_vp, _vp, C.POINTER(_u64), _int, _int, _int, _vp, _vp]),
    "flashe_sparse_expand": (_int, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "flashe_sparse_sum": (_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_u64), _vp, _int, _u64, _vp, _vp]),
    "flashe_sparse_apply_masks": (_int, [_vp, _u32, _i32p, _i32p, _int, _spanp, _vp, _vp, _u64, _vp]),
    "flashe_sparse_apply_masks_batch": (_int, [_vp, _u32, _i32p, _int, _int, C.POINTER(_u64), _u32, C.POINTER(_vp), _vp, _u64, _vp]),
    "flashe_sparse_overlap": (_int, [_vp, C.POINTER(_vp), C.POINTER(_u64), _int, _u64, C.POINTER(_u64), _vp]),
    "flashe_wire_nbytes": (_int, [_int, _u64, C.POINTER(_u64)]),
    "flashe_wire_pack": (_int, [_vp, _vp, _int, _u64, _int, _vp, _vp]),
    "flashe_wire_unpack": (_int, [_vp, _vp, _u64, _int, _int, _vp, _vp]),
    "flashe_topk_sparsify": (_int, [_vp, _vp, _vp, _u64, C.POINTER(_u64), C.POINTER(_u64), _int, _vp, _vp, _vp, _vp]),
    "flashe_segment_stats": (_int, [_vp, _vp, _vp, _u64, C.POINTER(_u64), C.POINTER(C.c_double), _int, _int, _vp, _vp]),
    "flashe_peer_alloc": (_int, [_vp, _u64, C.POINTER(_vp), _u8p]),
    "flashe_peer_open": (_int, [_vp, _u8p, C.POINTER(_vp)]),
    "flashe_peer_close": (_int, [_vp, _vp]),
    "flashe_peer_free": (_int, [_vp, _vp]),
    "flashe_launch_count": (_u64, []),
}

_lib = None


def library_path():
    # FLASHE_B200_LIB: alternative build of the same library (kernel tuning experiments)
    return os.environ.get("FLASHE_B200_LIB") or _build.LIB


def load():
    """Load libflashe_b200.so (building it first when nvcc is available and the source is newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "libflashe_b200.so is missing (%s). Build it with `python -m flashe_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise RuntimeError("libflashe_b200.so failed to load: %s" % e)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise RuntimeError("libflashe_b200.so does not export %s (stale build?)" % name)
        fn.restype = res
        fn.argtypes = args
    if lib.flashe_abi_version() != ABI_VERSION:
        raise RuntimeError("libflashe_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise FlasheError(rc, load().flashe_last_error().decode("utf-8", "replace"))
    return rc
