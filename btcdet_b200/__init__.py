"""btcdet_b200 — B200-native (sm_100a) hot path of BtcDet behind a C ABI.

Layout: csrc/ (CUDA kernels + extern "C" entry points, built in-tree into libbtcdet_b200.so),
_lib.py (ctypes binding, no fallback), ops.py (torch-facing operators / autograd), engine.py
(sync-free planned forward under CUDA graphs), synthetic.py (seeded KITTI-range scenes).
The drop-in `spconv` package at the repo root is the reference-facing surface.
"""
__version__ = "0.1.0"
