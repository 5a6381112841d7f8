"""`RoIHeadTemplate.proposal_layer` (btcdet/models/roi_heads/roi_head_template.py:46-101; SURVEY §8(f) N2: "proposal_layer
per-scene loop") without host synchronisation.

The reference loops over the scenes of a batch and, per scene, runs `class_agnostic_nms`
(btcdet/models/model_utils/model_nms_utils.py:6-28) -> `iou3d_nms_utils.nms_gpu` (iou3d_nms_utils.py:85-99): a top-k, a
sort, the N x N/64 suppression mask on the device, a copy of that mask to the host, the greedy scan on the host, and a
variable-length gather — one blocking device-to-host copy per scene, result shapes that depend on it.  Here every scene
issues the same torch selection ops (so ties resolve identically) and `btc_nms` (mask + greedy scan on the device, keep
list and count stay there); the variable-length tail becomes a mask, so the whole layer is a fixed launch sequence with
no host read (CUDA-graph capturable) and returns the identical `rois` / `roi_scores` / `roi_labels`.
"""
import torch

from . import iou3d_nms_cuda as _ext


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


def proposal_layer(batch_dict, nms_config):
    """batch_dict: batch_size, batch_box_preds (B, N, 7+C), batch_cls_preds (B, N, num_class | 1) — the dense form the
    anchor head produces (the stacked `batch_index` form is not handled).  nms_config: NMS_TYPE (nms_gpu |
    nms_normal_gpu), NMS_THRESH, NMS_PRE_MAXSIZE, NMS_POST_MAXSIZE, MULTI_CLASSES_NMS (False).
    Adds rois (B, NMS_POST_MAXSIZE, 7+C), roi_scores, roi_labels (1-based), has_class_labels, like the reference."""
    if batch_dict.get('batch_index', None) is not None:
        raise NotImplementedError("proposal_layer: stacked predictions (batch_index) are not supported")
    if _get(nms_config, 'MULTI_CLASSES_NMS'):
        raise NotImplementedError          # as in the reference (roi_head_template.py:83-84)
    nms_type = _get(nms_config, 'NMS_TYPE')
    assert nms_type in ('nms_gpu', 'nms_normal_gpu'), nms_type
    thresh = float(_get(nms_config, 'NMS_THRESH'))
    pre, post = int(_get(nms_config, 'NMS_PRE_MAXSIZE')), int(_get(nms_config, 'NMS_POST_MAXSIZE'))
    batch_size = batch_dict['batch_size']
    batch_box_preds, batch_cls_preds = batch_dict['batch_box_preds'], batch_dict['batch_cls_preds']
    assert batch_cls_preds.dim() == 3
    rois = batch_box_preds.new_zeros((batch_size, post, batch_box_preds.shape[-1]))
    roi_scores = batch_box_preds.new_zeros((batch_size, post))
    roi_labels = batch_box_preds.new_zeros((batch_size, post), dtype=torch.long)
    slot = torch.arange(post, device=batch_box_preds.device)
    for index in range(batch_size):
        box_preds, cls_preds = batch_box_preds[index], batch_cls_preds[index]
        cur_roi_scores, cur_roi_labels = torch.max(cls_preds, dim=1)
        n = int(cur_roi_scores.shape[0])
        if n == 0:
            continue
        # the reference's own selection ops, in its order: top-k (model_nms_utils.py:15), then the descending sort inside
        # nms_gpu / nms_normal_gpu (iou3d_nms_utils.py:93 / 110; NMS_PRE_MAXSIZE reaches that call only as an ignored kwarg)
        box_scores_nms, indices = torch.topk(cur_roi_scores, k=min(pre, n))
        order = box_scores_nms.sort(0, descending=True)[1]
        boxes = box_preds[indices][:, 0:7][order].contiguous().float()
        keep, num = _ext.ops_nms(boxes, thresh, normal=(nms_type == 'nms_normal_gpu'))   # device keep list + count
        k = int(boxes.shape[0])
        take = min(post, k)
        valid = slot[:take] < num.to(torch.long)                                       # kept rows among the first `post`
        src = indices[order[keep[:take].clamp(0, k - 1)]]                              # rows beyond the count: masked below
        rois[index, :take] = torch.where(valid.unsqueeze(-1), box_preds[src], rois[index, :take])
        roi_scores[index, :take] = torch.where(valid, cur_roi_scores[src], roi_scores[index, :take])
        roi_labels[index, :take] = torch.where(valid, cur_roi_labels[src], roi_labels[index, :take])
    batch_dict['rois'] = rois
    batch_dict['roi_scores'] = roi_scores
    batch_dict['roi_labels'] = roi_labels + 1
    batch_dict['has_class_labels'] = True if batch_cls_preds.shape[-1] > 1 else False
    batch_dict.pop('batch_index', None)
    return batch_dict
