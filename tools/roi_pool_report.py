"""Timings of the RoI grid pooling ops on one B200, at the reference yaml's training shape
(B = 2 scenes x 128 RoIs x 27 grid points; 20 k points per scene): this library's kernels next to the reference's own
CUDA kernels (oracle/_ref/libpointnet2_ref.so) and the reference's torch code, same inputs, same GPU, CUDA events.

    python tools/roi_pool_report.py [out.json]        # -> gpurun_out/roi_pool_report.json
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _time(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return {"median_us": round(1e3 * t[len(t) // 2], 1), "min_us": round(1e3 * t[0], 1)}


def main(out_path):
    import ref_loader
    import test_roi_pool_gpu as T
    from btcdet_b200 import pointnet2_stack_cuda as ext, roi_pool, synthetic as S
    from oracle import roi_pool as R
    rep = {"device": torch.cuda.get_device_name(0), "shape": "B=2 x 128 RoIs x 27 grid points, 20k points per scene"}
    case = S.roi_head_case(batch=2, n_points=20000, n_rois=128, n_occ=3000)
    xyz, cnt = T._scene_points(case)
    q, qcnt = T._queries(case)
    xyz_t, cnt_t, q_t, qcnt_t = [torch.from_numpy(a).cuda() for a in (xyz, cnt, q, qcnt)]
    M = q.shape[0]
    outs = [torch.zeros((M, ns), dtype=torch.int32, device="cuda") for ns in T.NSAMPLES]
    rep["ball_query"] = {"queries": M, "points": int(xyz.shape[0]), "radii": T.RADII, "nsample": T.NSAMPLES,
                         "ours_all_radii_one_launch": _time(lambda: ext.ball_query_multi(T.RADII, T.NSAMPLES, q_t, qcnt_t, xyz_t, cnt_t, outs))}
    if R.RefPointnet2.available():
        ref = R.RefPointnet2()
        lib = ref.lib

        def ref_all():
            for r, ns, o in zip(T.RADII, T.NSAMPLES, outs):
                lib.ref_ball_query_stack(2, M, ctypes.c_float(r), ns, q_t.data_ptr(), qcnt_t.data_ptr(),
                                         xyz_t.data_ptr(), cnt_t.data_ptr(), o.data_ptr())
        # the reference launches on the legacy default stream; torch's current stream is that stream here
        rep["ball_query"]["reference_kernels_four_launches"] = _time(ref_all, reps=5, warm=1)
        ext.ball_query_multi(T.RADII, T.NSAMPLES, q_t, qcnt_t, xyz_t, cnt_t, outs)
        rep["ball_query"]["identical_to_reference_kernels"] = bool(all(
            torch.equal(ref.ball_query(r, ns, xyz_t, cnt_t, q_t, qcnt_t), o) for r, ns, o in zip(T.RADII, T.NSAMPLES, outs)))
    # reverse trilinear gather + compaction
    sp = T._sparse(case)
    pts, zyx = T._grid_targets(case)
    per_scene = pts.shape[1]
    rep["trilinear"] = {"targets": int(zyx.shape[0]), "source_rows": int(sp.features.shape[0]), "channels": 128,
                        "ours_fused": _time(lambda: roi_pool.trilinear_gather_rows(sp, zyx, per_scene, [2, 4, 12])),
                        "torch_expression_of_the_reference": _time(
                            lambda: R.interpolate_rows(sp.features, sp.indices, 2, case["x_shape"], zyx, per_scene, [2, 4, 12]),
                            reps=5, warm=1)}
    c, r = roi_pool.trilinear_gather_rows(sp, zyx, per_scene, [2, 4, 12])
    wc, wf, _ = R.interpolate_rows(sp.features, sp.indices, 2, case["x_shape"], zyx, per_scene, [2, 4, 12])
    rep["trilinear"]["rows"] = int(r.shape[0])
    rep["trilinear"]["identical"] = bool(torch.equal(r, wf) and torch.equal(c.long(), wc))
    # algorithmic bytes of the fused gather: targets in (12 B) + 8 index-volume probes + the emitted rows out + the
    # source rows read once
    alg = zyx.shape[0] * (12 + 32) + r.shape[0] * (128 * 4 + 16) + sp.features.shape[0] * 128 * 4
    rep["trilinear"]["algorithmic_bytes"] = int(alg)
    rep["trilinear"]["algorithmic_GBps"] = round(alg / (rep["trilinear"]["ours_fused"]["median_us"] * 1e-6) / 1e9, 1)
    # whole roi_conv_pool of the reference's ConvHead, unchanged vs with the fused ops
    if ref_loader.available():
        mods = ref_loader.load_roi_head_modules(device="cuda")
        torch.manual_seed(0)
        head = ref_loader.build_conv_head(mods, S.DET_VOXEL_SIZE, S.KITTI_RANGE).cuda().eval()
        with torch.no_grad():
            want, _ = head.roi_conv_pool(T._batch_dict(case))
            t_ref = _time(lambda: head.roi_conv_pool(T._batch_dict(case)), reps=5, warm=1)
            roi_pool.patch_conv_head(head)
            got, _ = head.roi_conv_pool(T._batch_dict(case))
            t_fused = _time(lambda: head.roi_conv_pool(T._batch_dict(case)), reps=5, warm=1)
        rep["roi_conv_pool"] = {"reference_python_on_drop_ins": t_ref, "with_fused_ops": t_fused,
                                "identical": bool(torch.equal(got, want)), "out_shape": list(want.shape),
                                "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(rep, fh, indent=1)
    print(json.dumps(rep))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "roi_pool_report.json"))
