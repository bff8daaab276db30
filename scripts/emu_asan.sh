#!/bin/bash
# AddressSanitizer run of the CUDA sources on the CPU emulator (no GPU needed): heap overflows of any "device"
# allocation, use-after-free and double frees in kernels and in the ABI's host code.
cd "$(dirname "$0")/.."
python -c "import sys; sys.path.insert(0, 'tests'); import emu_lib; print(emu_lib.build(sanitize=True))" || exit 1
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
  python scripts/emu_asan.py
