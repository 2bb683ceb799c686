#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Builds the UNMODIFIED reference CUDA library (the 11 translation units of
# deps/CustomOps/FWI/Src/Makefile:7-10) for sm_100 straight from
# /root/reference, plus oracle/ref_shim.cu, into oracle/_ref/libCUFD_ref.so.
# Outputs go ONLY to oracle/_ref/ (git-ignored, travels to the GPU box).
# The reference's own build system (CMake + ADCME + TensorFlow) is not run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${FWI_REFERENCE_ROOT:-/root/reference}/deps/CustomOps/FWI/Src"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
if [ -f "$OUT/libCUFD_ref.so" ] && [ "$OUT/libCUFD_ref.so" -nt "$HERE/ref_shim.cu" ]; then
  echo "build_ref: up to date"; exit 0
fi
NVCC="${NVCC:-nvcc}"
FLAGS=(-O3 -std=c++14 -gencode arch=compute_100,code=sm_100 -x cu -DNDEBUG
       -I "$REF" -I "$REF/rapidjson" -Xcompiler -fPIC -w)
SRCS=(Parameter.cpp libCUFD.cu el_stress.cu el_velocity.cu el_stress_adj.cu
      el_velocity_adj.cu Model.cu Cpml.cu utilities.cu Src_Rec.cu Boundary.cu)
pids=()
for s in "${SRCS[@]}"; do
  "$NVCC" "${FLAGS[@]}" -c "$REF/$s" -o "$OUT/${s%.*}.o" &
  pids+=($!)
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_shim.cu" -o "$OUT/ref_shim.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC \
  "$OUT"/*.o -lcufft -o "$OUT/libCUFD_ref.so"
rm -f "$OUT"/*.o
echo "build_ref: built $OUT/libCUFD_ref.so"
