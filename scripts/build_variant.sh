#!/bin/bash
# usage: scripts/build_variant.sh <name> [extra nvcc flags...]  -> flashe_b200/_lib/libflashe_b200_<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC -cudart static -Xptxas -v "$@" \
  -o $ROOT/flashe_b200/_lib/libflashe_b200_$name.so $ROOT/flashe_b200/csrc/flashe_kernels.cu > /tmp/build_$name.log 2>&1 || { grep -E "error" /tmp/build_$name.log; exit 1; }
echo "== $name $@"
grep -A2 "k_streamILi1ELi4ELi2ELb[01]" /tmp/build_$name.log | grep -E "Used|spill" | tr '\n' ' '; echo
cuobjdump -sass $ROOT/flashe_b200/_lib/libflashe_b200_$name.so | awk '/Function : /{n=$3} /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/{c[n]++} END{print "instrs enc unshared/shared:", c["_Z8k_streamILi1ELi4ELi2ELb0EEv8KeySched9StreamTab4Geom5IoDev8CodecDev8NoiseDev"], c["_Z8k_streamILi1ELi4ELi2ELb1EEv8KeySched9StreamTab4Geom5IoDev8CodecDev8NoiseDev"]}'
