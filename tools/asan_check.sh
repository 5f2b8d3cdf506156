#!/bin/bash
# Host-side memory check: builds the library and the test-only emulator with AddressSanitizer out of tree (/tmp/mrhyde_b200_asan) and runs
# the CPU tests that exercise the host code (plan construction, expression compiler, stage replay, builders, gloo multirank) against it.
# The device code is compiled too but never runs here (no GPU in this container).  Last runs: address 135 passed, undefined 133 passed
# (without the gloo file), no report.     usage: tools/asan_check.sh [address|undefined]
set -e
SAN=${1:-address}
RT=$([ "$SAN" = address ] && echo libasan.so || echo libubsan.so)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/mrhyde_b200_$SAN
rm -rf $W && mkdir -p $W/a/b $W/a/include
cp -r $ROOT/mrhyde_b200/csrc $W/a/b/csrc && rm -rf $W/a/b/csrc/build
cp $ROOT/include/mrhyde_b200.h $W/a/include/
cd $W/a/b/csrc
sed -i 's/-Xcompiler -fPIC,-Wall,-Wno-unused-function/-Xcompiler -fPIC,-Wall,-Wno-unused-function,-fsanitize='$SAN',-fno-omit-frame-pointer/; s/^CXXFLAGS := -O2/CXXFLAGS := -O1 -g -fsanitize='$SAN' -fno-omit-frame-pointer/; s#^OUT := .*#OUT := '$W'/libmrhyde_b200.so#; s#^EMU := .*#EMU := '$W'/libmrhyde_b200_emulate.so#; s#-shared -o \$@ \$(OBJ) -lcudart#-shared -o $@ $(OBJ) -Xcompiler -fsanitize='$SAN' -lcudart#; s#g++ -shared -o \$@ build/general_emulate.o#g++ -shared -fsanitize='$SAN' -o $@ build/general_emulate.o#' Makefile
make -s -j8
cd $ROOT
# libstdc++ must be loaded before the sanitizer runtime intercepts __cxa_throw in a python process
LD_PRELOAD="$(gcc -print-file-name=$RT) $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0:log_path=$W/report UBSAN_OPTIONS=print_stacktrace=1:log_path=$W/report \
  MRHYDE_B200_LIB=$W/libmrhyde_b200.so python -m pytest tests/test_abi_cpu.py tests/test_general_emulation.py tests/test_builders.py tests/test_multirank_gloo.py -q -p no:cacheprovider
ls $W/report.* 2>/dev/null && { echo "sanitizer reports above"; exit 1; } || echo "no sanitizer report"
