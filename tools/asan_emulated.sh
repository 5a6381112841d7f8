#!/bin/bash
# Memory check of the kernel source without a GPU (compute-sanitizer needs one): the host-emulated builds of tests/host_emul/
# compiled with AddressSanitizer, the emulated CPU tests run under it.  Any out-of-bounds read / write of a kernel on the
# tests' inputs aborts with the kernel's source line.
#   tools/asan_emulated.sh            the default emulated tests, OS-thread engine                     (~3 min)
#   tools/asan_emulated.sh --full     + the full-size cases (20k-point scenes on the KITTI grid, both backbones' pyramids,
#                                     the reference topologies through the shim), fiber engine         (~20 min)
set -e
cd "$(dirname "$0")/.."
ASAN=$(readlink -f "$(g++ -print-file-name=libasan.so)")
OPTS=detect_leaks=0:verify_asan_link_order=0:halt_on_error=1:detect_stack_use_after_return=0
if [ "$1" = "--full" ]; then
  shift
  BTC_EMUL_FULL=1 BTC_EMUL_FIBERS=1 BTC_EMUL_SANITIZE=address LD_PRELOAD=$ASAN ASAN_OPTIONS=$OPTS \
    python -m pytest tests/test_emulated_kernels_cpu.py -x -q -p no:cacheprovider --timeout=3600 "$@"
else
  BTC_EMUL_SANITIZE=address LD_PRELOAD=$ASAN ASAN_OPTIONS=$OPTS \
    python -m pytest tests/test_emulated_kernels_cpu.py tests/test_roi_pool_cpu.py -x -q -p no:cacheprovider "$@"
fi
