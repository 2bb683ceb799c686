#!/bin/bash
# Build a kernel variant of libfwi_b200.so under variants/ (git-ignored, travels to the GPU box):
#   scripts/build_variant.sh name "-DFWI_L2PF=0 ..."
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants"
make -s -C "$ROOT/fwiflow/jl_b200/csrc" OUT="$ROOT/variants/libfwi_$1.so" EXTRA="$2" 2>&1 | grep -E "error|ptxas info.*spill" || true
ls -la "$ROOT/variants/libfwi_$1.so"
