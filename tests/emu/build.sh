#!/bin/sh
# TEST-ONLY: builds the CPU functional model of the CUDA library (same sources, g++, fiber CTA emulator).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
src="$root/algames.jl_b200/csrc"
mkdir -p "$here/obj"
pids=""
for f in agb_capi agb_band agb_kernels_p1 agb_kernels_p2 agb_kernels_p3 agb_kernels_p3b agb_kernels_p3m agb_kernels_p3t agb_kernels_p4; do
  g++ -O2 -std=c++17 -fPIC -DAGB_EMULATE -I "$here" -x c++ -c "$src/$f.cu" -o "$here/obj/$f.o" &
  pids="$pids $!"
done
g++ -O2 -std=c++17 -fPIC -DAGB_EMULATE -I "$here" -c "$here/emu.cpp" -o "$here/obj/emu.o" &
pids="$pids $!"
for p in $pids; do wait $p; done
g++ -shared -o "$here/libagb_emu.so" "$here"/obj/*.o
