"""Builds libflashe_b200.so in-tree (flashe_b200/_lib/) with nvcc for sm_100a.

The library is the product: there is no CPU fallback.  `python -m flashe_b200.build` or
`__graft_entry__.build()` compile it; nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", f) for f in ("flashe_kernels.cu", "flashe_stream_masks.cu", "flashe_stream_apply.cu", "flashe_stream_encode.cu",
                                                    "flashe_stream_encode_shared.cu", "flashe_stream_encode_n32.cu", "flashe_stream_decode.cu", "flashe_stream_scatter.cu",
                                                    "flashe_elementwise.cu", "flashe_wire.cu", "flashe_stats.cu")]
DEPS = SRCS + [os.path.join(HERE, "csrc", f) for f in ("flashe_internal.h", "flashe_device.cuh", "flashe_codec_host.cuh", "flashe_stream.cuh",
                                                           "flashe_stream_decl.h")]
HDR = os.path.join(os.path.dirname(HERE), "include", "flashe_b200.h")
LIB_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIB_DIR, "libflashe_b200.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--shared", "-Xcompiler", "-fPIC", "-cudart", "static", "--threads", "0"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libflashe_b200.so cannot be built")
    return p


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS + [HDR])


def build_library(force=False, verbose=False, variant=None, extra_flags=()):
    """variant / extra_flags: a tuning build of the same library under _lib/libflashe_b200_<variant>.so (select it at
    run time with FLASHE_B200_LIB=<path>; scripts/gpu_ab*.sh interleave such builds on one box)."""
    out = LIB if not variant else os.path.join(LIB_DIR, "libflashe_b200_%s.so" % variant)
    if not variant and not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRCS
    subprocess.check_call(cmd)
    return out


MICRO_SRC = os.path.join(HERE, "csrc", "micro", "flashe_micro.cu")
MICRO_LIB = os.path.join(LIB_DIR, "libflashe_micro.so")


def build_micro(force=False, verbose=False):
    """libflashe_micro.so: the micro-benchmarks of scripts/microbench.py (LDS lookup ceiling, bit-sliced AES-256).
    Not part of the product library."""
    deps = [MICRO_SRC, os.path.join(HERE, "csrc", "micro", "sbox_bitslice.inc")]
    if not force and os.path.exists(MICRO_LIB) and all(os.path.getmtime(p) <= os.path.getmtime(MICRO_LIB) for p in deps):
        return MICRO_LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", MICRO_LIB, MICRO_SRC]
    subprocess.check_call(cmd)
    return MICRO_LIB


if __name__ == "__main__":
    if "--micro" in sys.argv:
        print(build_micro(force="--force" in sys.argv, verbose="-v" in sys.argv))
    elif "--variant" in sys.argv:      # python -m flashe_b200.build --variant NAME [-DFLAG=1 ...]
        i = sys.argv.index("--variant")
        print(build_library(variant=sys.argv[i + 1], extra_flags=[a for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
