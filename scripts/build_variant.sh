#!/bin/bash
# Build a kernel-experiment variant of both libraries in-tree:  scripts/build_variant.sh NAME "-DFOO=1 -DBAR=2"
# -> parsmurf_b200/_build_NAME/ ; select it at run time with PSM_BUILD_DIR=$PWD/parsmurf_b200/_build_NAME
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/parsmurf_b200/_build_$1
mkdir -p "$OUT"
make -s -j8 -C "$ROOT/parsmurf_b200/csrc" OUT="$OUT" EXTRA="$2"
make -s -C "$ROOT/parsmurf_b200/host" OUT="$OUT"
echo "built $OUT"
