#!/bin/bash
# Builds libqasr.so (sm_100a) next to the Python host package.  Used by __graft_entry__.build().
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../qasr/libqasr.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
       -I"$HERE/../../include" -I"$HERE" --expt-relaxed-constexpr ${QASR_NVCC_EXTRA:-})
mkdir -p "$HERE/build"
pids=()
for f in api dwconv dwconv_rt dwconv_tc dwconv_tape gemm_tc conv_ntc elementwise melspec dynamic; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || \
     [ "$HERE/qasr_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/tc_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/dw_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/../../include/qasr.h" -nt "$HERE/build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "$HERE"/build/{api,dwconv,dwconv_rt,dwconv_tc,dwconv_tape,gemm_tc,conv_ntc,elementwise,melspec,dynamic}.o -lcudart
echo "built $OUT"
