"""Builds libflashe_b200.so in-tree (flashe_b200/_lib/) with nvcc for sm_100a.

The library is the product: there is no CPU fallback.  `python -m flashe_b200.build` or
`__graft_entry__.build()` compile it; nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = [os.path.join(HERE, "csrc", f) for f in ("flashe_kernels.cu", "flashe_stream_masks.cu", "flashe_stream_apply.cu", "flashe_stream_encode.cu",
                                                    "flashe_stream_encode_shared.cu", "flashe_stream_decode.cu", "flashe_stream_scatter.cu",
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


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SRCS
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
