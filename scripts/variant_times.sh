#!/bin/bash
# On the GPU box: time the hot kernels of every variants/libfwi_*.so (and the in-tree build) at C2 and C3 shapes.
#   scripts/variant_times.sh [c2 shots] [c3 shots]
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
for lib in fwiflow/jl_b200/libfwi_b200.so variants/libfwi_*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  FWI_B200_LIB="$ROOT/$lib" timeout 120 python scripts/kernel_times.py c2 ${1:-30} 200 2>&1 | grep -v "^fwd " | tail -6
  if [ "${2:-0}" != "0" ]; then FWI_B200_LIB="$ROOT/$lib" timeout 180 python scripts/kernel_times.py c3 $2 50 2>&1 | tail -4; fi
done
