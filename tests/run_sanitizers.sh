#!/bin/bash
# Runs parity cases through ASan+UBSan and TSan builds of the host-emulation library (the device function bodies executed on
# the CPU plus all host code).  Not part of pytest: the sanitizer runtimes have to be preloaded into the interpreter.
#   bash tests/run_sanitizers.sh          (about two minutes)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
make -s -C "$ROOT/autocycler_b200/csrc" sanitize
DRIVER="$ROOT/tests/sanitizer_cases.py"
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 AC_EXPAND_MIN_DUE=1 AC_CHECK_CANDIDATES=1 \
  python "$DRIVER" "$ROOT/tests/emu/libautocycler_emu_asan.so" 12
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0" AC_EXPAND_MIN_DUE=1 AC_HOST_THREADS=6 \
  python "$DRIVER" "$ROOT/tests/emu/libautocycler_emu_tsan.so" 4
echo "sanitizers clean"
