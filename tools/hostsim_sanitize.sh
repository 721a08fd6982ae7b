#!/bin/bash
# AddressSanitizer + UBSan over the HOST build of the kernels' per-thread code (tests/hostsim: the same .cuh files compiled with -DDRP_HOSTSIM):
# LBVH phases, optimal collapse, wide traversal (fast path and the fix-up kernel's strided stack), refit, instanced assembly, shading, tone map.
# Complements compute-sanitizer on the GPU (tools/sanitize.sh); runs in the GPU-less build container.  Log: profiles/r2/hostsim_asan_ubsan_r2.log
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so); UBSAN=$(gcc -print-file-name=libubsan.so)
cp tests/hostsim/libhostsim.so /tmp/libhostsim_orig.so 2>/dev/null || true
g++ -O1 -g -std=c++17 -fPIC -fopenmp -mavx2 -mfma -ffp-contract=off -DDRP_HOSTSIM -fsanitize=undefined,address -fno-sanitize-recover=undefined \
    -x c++ -shared -o tests/hostsim/libhostsim.so tests/hostsim/hostsim.cpp
nm -D tests/hostsim/libhostsim.so | grep -c "__asan\|__ubsan" | sed 's/^/instrumented symbols: /'
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_hostsim.py -q 2>&1 | tail -5
rm -f tests/hostsim/libhostsim.so; make -C tests/hostsim > /dev/null   # back to the plain build
# the same for the CPU oracle (the checker itself): oracle/*.c under ASan + UBSan over its golden-vector tests
cp oracle/liborc.so /tmp/liborc_orig.so 2>/dev/null || true
gcc -O1 -g -fPIC -fopenmp -mavx2 -mfma -ffp-contract=off -fno-fast-math -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -shared -o oracle/liborc.so oracle/orc_raycast.c oracle/orc_shade.c oracle/orc_tonemap.c -lm
LD_PRELOAD="$ASAN $UBSAN" ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 python -m pytest tests/test_oracle_golden.py -q 2>&1 | tail -3
rm -f oracle/liborc.so; make -C oracle > /dev/null   # back to the plain build
