#!/bin/sh
# TEST-ONLY: builds the CPU functional model of the CUDA library (same sources, g++, fiber CTA emulator).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
g++ -O2 -std=c++17 -fPIC -shared -DAGB_EMULATE -I "$here" -x c++ "$root/algames.jl_b200/csrc/agb_capi.cu" -x c++ "$here/emu.cpp" \
    -o "$here/libagb_emu.so"
