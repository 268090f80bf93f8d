#!/bin/bash
# The kernels' own source (csrc/spmv_generic_kernels.cuh) on the host emulator under AddressSanitizer + UBSan: every access a kernel
# makes to off / col / val / x / y in tests/test_generic_emulation.py is bounds-checked against the numpy buffers it was given.
# Round 2: 86 passed, no sanitizer report (includes the COO / Sliced-ELL kernels whose first hardware run is pending).
set -e
cd "$(dirname "$0")/.."
B=tests/host_emulation/_build; mkdir -p $B
g++ -O1 -g -std=c++17 -pthread -shared -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -Itests/host_emulation \
    -Icudalibrarysamples_b200/csrc tests/host_emulation/emulate_generic.cpp -o $B/libgeneric_emu.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 \
    python -m pytest tests/test_generic_emulation.py -x -q -p no:cacheprovider
rm -f $B/libgeneric_emu.so        # the next ordinary test run rebuilds the plain library
