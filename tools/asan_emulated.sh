#!/bin/bash
# Memory check of the kernel source without a GPU (compute-sanitizer needs one): the host-emulated builds of tests/host_emul/
# compiled with AddressSanitizer, the emulated CPU tests run under it.  Any out-of-bounds read / write of a kernel on the
# tests' inputs aborts with the kernel's source line.  ~2.5 min.
set -e
cd "$(dirname "$0")/.."
ASAN=$(g++ -print-file-name=libasan.so)
BTC_EMUL_SANITIZE=address LD_PRELOAD=$(readlink -f "$ASAN") \
  ASAN_OPTIONS=detect_leaks=0:verify_asan_link_order=0:halt_on_error=1 \
  python -m pytest tests/test_emulated_kernels_cpu.py tests/test_roi_pool_cpu.py -x -q -p no:cacheprovider "$@"
