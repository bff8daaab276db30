#!/bin/bash
# ThreadSanitizer run of decomposed cases on the CPU emulator (ranks = threads): data races of the peer-store
# halo protocol (a ghost cell written while its reader may still run, or read before its writer's word).
cd "$(dirname "$0")/.."
python -c "import sys; sys.path.insert(0, 'tests'); import emu_lib; print(emu_lib.build(sanitize='thread'))" || exit 1
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) OMP_NUM_THREADS=1 TSAN_OPTIONS="report_bugs=1 halt_on_error=0 history_size=7 second_deadlock_stack=0" \
  python scripts/emu_tsan.py "$@"
