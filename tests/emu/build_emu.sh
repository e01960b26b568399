#!/bin/sh
# TEST INFRASTRUCTURE ONLY: compile the kernel sources for the HOST (-DSPB_EMU) so the CPU-only test
# tier can check kernel logic against the oracle.  The product never loads this library.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
g++ -O2 -g -std=c++17 -fPIC -shared -DSPB_EMU -x c++ "$ROOT/superpang_b200/csrc/spb.cu" -o "$HERE/libspb_emu.so" -Wall -Wno-unused-function -Wno-unused-variable -Wno-unknown-pragmas
