"""Drop-in for the reference's compiled extension `btcdet.ops.iou3d_nms.iou3d_nms_cuda`.

Same five entry points and calling convention as `btcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17` — the caller
pre-allocates the outputs, the functions return an int — so `btcdet/ops/iou3d_nms/iou3d_nms_utils.py` (boxes_iou_bev,
boxes_iou3d_gpu, nms_gpu, nms_normal_gpu) runs unchanged on top of it:

    import sys, btcdet_b200.iou3d_nms_cuda as m
    sys.modules["btcdet.ops.iou3d_nms.iou3d_nms_cuda"] = m      # before btcdet.ops.iou3d_nms is imported

The kernels are this library's (`csrc/iou3d_nms.cu`: clip-in-the-other-box's-frame overlap, device-side greedy scan).
`ops_nms` below is the sync-free form (keep list and count stay on the device).
"""
import ctypes

import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_boxes(*ts):
    for t in ts:
        if not t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or t.dim() != 2 or t.shape[1] != 7:
            raise RuntimeError("iou3d_nms: boxes must be contiguous float32 CUDA tensors of shape (N, 7)")


def _bev(boxes_a, boxes_b, out, mode):
    _check_boxes(boxes_a, boxes_b)
    assert out.is_cuda and out.is_contiguous() and out.dtype == torch.float32
    assert tuple(out.shape) == (boxes_a.shape[0], boxes_b.shape[0])
    check(_lib.load().btc_boxes_bev(_ptr(boxes_a), boxes_a.shape[0], _ptr(boxes_b), boxes_b.shape[0], mode, _ptr(out),
                                    _stream()), "btc_boxes_bev")
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """(N,7), (M,7) -> ans_overlap (N,M): area of the rotated BEV intersection (iou3d_nms.cpp:58-78)."""
    return _bev(boxes_a, boxes_b, ans_overlap, 1)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """(N,7), (M,7) -> ans_iou (N,M): rotated BEV IoU (iou3d_nms.cpp:80-100)."""
    return _bev(boxes_a, boxes_b, ans_iou, 0)


def ops_nms(boxes, thresh, normal=False):
    """Sync-free NMS over boxes sorted by descending score: returns (keep int64 [N] on the device, count int32 [1] on the
    device); the first `count` entries of `keep` are the kept row indices in ascending order."""
    _check_boxes(boxes)
    lib = _lib.load()
    n = boxes.shape[0]
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    ws_bytes = int(lib.btc_nms_workspace_bytes(n))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=boxes.device)
    check(lib.btc_nms(_ptr(boxes), n, ctypes.c_float(float(thresh)), int(bool(normal)), _ptr(keep), _ptr(num), _ptr(ws),
                      ws_bytes, _stream()), "btc_nms")
    return keep, num


def _nms_into(boxes, keep, thresh, normal):
    keep_dev, num = ops_nms(boxes, thresh, normal)
    n_out = int(num.item())                  # the reference returns the count to Python as well (iou3d_nms.cpp:139)
    keep[:n_out].copy_(keep_dev[:n_out])     # `keep` is the caller's CPU LongTensor (iou3d_nms_utils.py:97)
    return n_out


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes (N,7) sorted by score, keep: LongTensor (N) filled with the kept indices; returns their number."""
    return _nms_into(boxes, keep, nms_overlap_thresh, False)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """As nms_gpu with the axis-aligned IoU (headings ignored; iou3d_nms_kernel.cu:224-234)."""
    return _nms_into(boxes, keep, nms_overlap_thresh, True)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """The reference computes this one on the host (iou3d_cpu.cpp:232-252); this library has no CPU path, so the CPU
    tensors take a round trip through the GPU kernel."""
    assert not (boxes_a.is_cuda or boxes_b.is_cuda or ans_iou.is_cuda), "Only support CPU tensors"
    out = torch.empty(ans_iou.shape, dtype=torch.float32, device="cuda")
    _bev(boxes_a.float().contiguous().cuda(), boxes_b.float().contiguous().cuda(), out, 0)
    ans_iou.copy_(out.cpu())
    return 1
