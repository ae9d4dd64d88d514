"""Per-class score threshold -> 3D NMS -> detections-per-image limit behind the reference's name
(lib/core/test.py:806-878, called by core/test.py:114 and prm/peak_response_mapping_3d.py:124).

The NMS of every class runs in ONE batched launch of the CUDA kernels (nms3d.cu); thresholding and the final limit are
index bookkeeping on a few hundred rows.  Settings come from core.config.cfg inside the reference tree, else from the
keyword arguments (defaults = lib/core/config.py:199, 228, 233, 451)."""
import numpy as np

from . import boxes_3d


def _settings(kw):
    try:
        from core.config import cfg
        # the reference switches to soft_nms_3d / box_voting under these flags (core/test.py:839-861); neither is built
        # (both off in every shipped config, and the soft-NMS call site names a function that does not exist): fail loudly
        # instead of silently returning hard-NMS results
        if getattr(getattr(cfg.TEST, "SOFT_NMS", None), "ENABLED", False):
            raise NotImplementedError("cfg.TEST.SOFT_NMS.ENABLED is set: soft_nms_3d is not implemented by b200seg")
        if getattr(getattr(cfg.TEST, "BBOX_VOTE", None), "ENABLED", False):
            raise NotImplementedError("cfg.TEST.BBOX_VOTE.ENABLED is set: box voting is not implemented by b200seg")
        d = dict(num_classes=cfg.MODEL.NUM_CLASSES, rpn_only=cfg.MODEL.RPN_ONLY, score_thresh=cfg.TEST.SCORE_THRESH, nms=cfg.TEST.NMS,
                 detections_per_im=cfg.TEST.DETECTIONS_PER_IM)
    except ImportError:
        d = dict(num_classes=None, rpn_only=False, score_thresh=0.05, nms=0.3, detections_per_im=100)
    d.update({k: v for k, v in kw.items() if v is not None})
    return d


def box_results_with_nms_and_limit(scores, boxes, scores_keep_idx=None, num_classes=None, rpn_only=None, score_thresh=None, nms=None,
                                   detections_per_im=None):
    """scores [R, num_classes], boxes [R, 6*num_classes] -> (scores, boxes, cls_boxes, cls_keep_idx) as the reference.
    Soft-NMS and box voting are disabled in every shipped config (TEST.SOFT_NMS.ENABLED / BBOX_VOTE.ENABLED False) and are
    not offered.  Where the reference's limit step indexes `cls_keep_idx[j][keep, :]` (a TypeError for the default empty
    list, an IndexError for 1-D indices) the rows are selected along the first axis instead."""
    import torch
    s = _settings(dict(num_classes=num_classes, rpn_only=rpn_only, score_thresh=score_thresh, nms=nms, detections_per_im=detections_per_im))
    scores, boxes = np.asarray(scores), np.asarray(boxes)
    ncls = int(s["num_classes"] if s["num_classes"] is not None else scores.shape[1])
    cls_boxes = [[] for _ in range(ncls)]
    cls_keep_idx = [[] for _ in range(ncls)]
    dets, idxs = [], []
    for j in range(1, ncls):
        if s["rpn_only"]:
            inds = np.where(scores > s["score_thresh"])[1]
            scores_j, boxes_j = scores[0, inds], boxes[inds, :]
        else:
            inds = np.where(scores[:, j] > s["score_thresh"])[0]
            scores_j, boxes_j = scores[inds, j], boxes[inds, j * 6:(j + 1) * 6]
        dets.append(np.hstack((boxes_j, scores_j[:, np.newaxis])).astype(np.float32, copy=False))
        idxs.append(None if scores_keep_idx is None else np.asarray(scores_keep_idx)[inds].copy())
    if ncls > 1:
        sizes = [len(d) for d in dets]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        if off[-1] > 0:                                       # one batched NMS launch for all classes
            all_d = torch.from_numpy(np.ascontiguousarray(np.vstack(dets))).cuda()
            keep, cnt, _ = boxes_3d.nms_3d_batched(all_d, torch.from_numpy(off).cuda(), int(max(sizes)), float(np.float32(s["nms"])))
            keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
        for j in range(1, ncls):
            k = keep[off[j - 1]:off[j - 1] + cnt[j - 1]] if sizes[j - 1] else np.empty(0, np.int64)
            cls_boxes[j] = dets[j - 1][k, :]
            if idxs[j - 1] is not None:
                cls_keep_idx[j] = idxs[j - 1][k]
    if s["detections_per_im"] > 0 and ncls > 1:
        image_scores = np.hstack([cls_boxes[j][:, -1] for j in range(1, ncls)])
        if len(image_scores) > s["detections_per_im"]:
            image_thresh = np.sort(image_scores)[-s["detections_per_im"]]
            for j in range(1, ncls):
                k = np.where(cls_boxes[j][:, -1] >= image_thresh)[0]
                cls_boxes[j] = cls_boxes[j][k, :]
                if not isinstance(cls_keep_idx[j], list):
                    cls_keep_idx[j] = cls_keep_idx[j][k]
    im_results = np.vstack([cls_boxes[j] for j in range(1, ncls)]) if ncls > 1 else np.empty((0, 7), np.float32)
    return im_results[:, -1], im_results[:, :-1], cls_boxes, cls_keep_idx
