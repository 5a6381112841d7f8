// TEST INFRASTRUCTURE ONLY — never linked into or imported by the product (btcdet_b200/, spconv/).
//
// Compiles the REFERENCE'S OWN CPU implementation of the rotated BEV overlap / IoU
// (btcdet/ops/iou3d_nms/src/iou3d_cpu.cpp: box_overlap :128-220, iou_bev :222-230) from where it lies under the
// reference checkout — the source is #included by path (BTC_REF_IOU3D_CPU, given by oracle/Makefile), nothing of it is
// copied into this repository — and exposes it through a plain C ABI so that the tests can use the reference itself as
// the oracle for SURVEY §8(f) N2 (`oracle/_ref/libiou3d_ref.so`, git-ignored, travels to the GPU box).
#include BTC_REF_IOU3D_CPU

extern "C" {

float ref_box_overlap(const float* box_a, const float* box_b) { return box_overlap(box_a, box_b); }

float ref_iou_bev(const float* box_a, const float* box_b) { return iou_bev(box_a, box_b); }

// boxes [n, 7] / [m, 7] (x, y, z, dx, dy, dz, heading) -> out [n, m]; what = 0: IoU, 1: overlap area
void ref_boxes_bev(const float* a, int n, const float* b, int m, int what, float* out) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
            out[(long)i * m + j] = what ? box_overlap(a + i * 7, b + j * 7) : iou_bev(a + i * 7, b + j * 7);
}

}  // extern "C"
