#!/bin/bash
# Race check of the kernel source without a GPU (compute-sanitizer racecheck needs one): the host-emulated builds of
# tests/host_emul/ compiled with ThreadSanitizer — every CUDA thread of a block is an OS thread, __syncthreads / warp
# collectives / atomics are the only synchronisation TSan sees, so a missing barrier inside a block shows up as a data
# race with the kernel's source line.  (Blocks run one after the other: races BETWEEN blocks are not visible.)  ~12 min.
# Reports land in /tmp/tsan_emul.*; profiles/r2_emulated_sanitizers.md has the triage of the round-2 run.
cd "$(dirname "$0")/.."
rm -f /tmp/tsan_emul.*
BTC_EMUL_SANITIZE=thread LD_PRELOAD=$(readlink -f "$(g++ -print-file-name=libtsan.so)") \
  TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=0 log_path=/tmp/tsan_emul" \
  python -m pytest tests/test_emulated_kernels_cpu.py tests/test_roi_pool_cpu.py -q -p no:cacheprovider "$@"
grep -h "SUMMARY" /tmp/tsan_emul.* 2>/dev/null | sed 's#.*/_build/##' | sort | uniq -c | sort -rn
