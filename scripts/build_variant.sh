#!/bin/bash
# usage: scripts/build_variant.sh <name> [extra nvcc flags...]  -> flashe_b200/_lib/libflashe_b200_<name>.so
# (a complete library: both translation units; select it at run time with FLASHE_B200_LIB=<path>)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC -cudart static -Xptxas -v "$@" \
  -o $ROOT/flashe_b200/_lib/libflashe_b200_$name.so $ROOT/flashe_b200/csrc/flashe_kernels.cu $ROOT/flashe_b200/csrc/flashe_wire.cu \
  > /tmp/build_$name.log 2>&1 || { grep -E "error" /tmp/build_$name.log; exit 1; }
echo "== $name $@"
grep -A2 "k_streamILi1ELi4ELi2ELb[01]ELb1" /tmp/build_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
