#!/bin/sh
# Builds a VARIANT of libalgames_b200.so for experiments: recompiles the given translation units with extra nvcc flags into
# profiles/_variants/<name>/ and links them with the untouched objects of the regular build (algames.jl_b200/build).
# usage: profiles/build_variant.sh <name> "<extra nvcc flags>" unit [unit ...]      e.g.  timing "-DAGB_PHASE_TIMING" agb_kernels_p3
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; flags="$2"; shift 2
out="$root/profiles/_variants/$name"; mkdir -p "$out"
pids=""
for u in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets $flags \
    -c "$root/algames.jl_b200/csrc/$u.cu" -o "$out/$u.o" &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
objs=""
for u in agb_capi agb_band agb_kernels_p1 agb_kernels_p2 agb_kernels_p3 agb_kernels_p3b agb_kernels_p3m agb_kernels_p3t agb_kernels_p4; do
  if [ -f "$out/$u.o" ]; then objs="$objs $out/$u.o"; else objs="$objs $root/algames.jl_b200/build/$u.o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o "$out/libalgames_b200.so" $objs
echo "$out/libalgames_b200.so"
