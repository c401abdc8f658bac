#!/bin/sh
# Builds tests/emu/libmmn_emu.so: the kernel sources compiled for the HOST against cuda_emu.h.
# Test infrastructure only (see cuda_emu.h); the product never loads this library.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
g++ -x c++ -std=c++17 -O2 -g -DMMN_EMU -fPIC -shared -Wall -Wno-unknown-pragmas -Wno-unused-function \
    -I"$HERE" -I"$ROOT/include" -I"$ROOT/multimodn_b200/csrc" \
    -o "$HERE/libmmn_emu.so" "$ROOT/multimodn_b200/csrc/mmn_api.cu" "$ROOT/multimodn_b200/csrc/mmn_fma.cu" \
    "$ROOT/multimodn_b200/csrc/mmn_tc.cu" "$ROOT/multimodn_b200/csrc/mmn_tc2.cu" "$ROOT/multimodn_b200/csrc/mmn_nb.cu"
