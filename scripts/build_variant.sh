#!/bin/bash
# usage: scripts/build_variant.sh <name> [-DFLAG=1 ...]  -> flashe_b200/_lib/libflashe_b200_<name>.so
# (a complete library; select it at run time with FLASHE_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
python -m flashe_b200.build --variant "$@" 2>&1 | grep -v RuntimeWarning
