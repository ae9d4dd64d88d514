"""Instance-mask overlap matrices behind the reference names (tools/evaluation/mask_iou.py:49-109).

mask_overlaps_labels(pred, gt, pred_ids, gt_ids)  device op on two uint16 label volumes (what the evaluation scripts
                                                  really have, eval_instance_segmentation_soma.py:186-197)
mask_iou_fast / mask_ios_fast / mask_iog_fast      the reference's stack signature ([N,S,H,W] and [K,S,H,W] boolean
                                                  masks); the stacks are folded back into label volumes (instances of a
                                                  stack must be disjoint, as they are when cut out of a label volume)."""
import numpy as np

from . import _lib


def mask_overlaps_labels(pred, gt, pred_ids, gt_ids, want=("iou", "ios", "iog")):
    """pred, gt: uint16 label volumes of equal shape (numpy or cuda tensors); pred_ids / gt_ids: label ids, one row /
    column each, in the caller's order.  Returns dict(iou, ios, iog [Np, Ng] fp32 cuda tensors, inter, area_pred, area_gt int64)."""
    import torch
    L = _lib.lib()

    def dev(x):
        if hasattr(x, "data_ptr"):
            return x.to(device="cuda", dtype=torch.uint16).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint16)).cuda()
    p, g = dev(pred), dev(gt)
    if p.shape != g.shape:
        raise IndexError                                      # mask_iou.py:33-34
    pred_ids = np.asarray(pred_ids, dtype=np.int64).ravel()
    gt_ids = np.asarray(gt_ids, dtype=np.int64).ravel()
    Np, Ng = pred_ids.size, gt_ids.size
    lut_p = np.full(65536, -1, np.int32); lut_p[pred_ids] = np.arange(Np, dtype=np.int32)
    lut_g = np.full(65536, -1, np.int32); lut_g[gt_ids] = np.arange(Ng, dtype=np.int32)
    d = p.device
    lp, lg = torch.from_numpy(lut_p).to(d), torch.from_numpy(lut_g).to(d)
    out = {k: (torch.empty((Np, Ng), dtype=torch.float32, device=d) if k in want else None) for k in ("iou", "ios", "iog")}
    inter = torch.empty((Np + 1, Ng + 1), dtype=torch.int64, device=d)
    ap = torch.empty(Np + 1, dtype=torch.int64, device=d)
    ag = torch.empty(Ng + 1, dtype=torch.int64, device=d)
    ws_bytes = L.b200seg_mask_overlaps_workspace_bytes(Np, Ng)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d)
    _lib.check(L.b200seg_mask_overlaps_dev(_lib.ptr(p), _lib.ptr(g), p.numel(), _lib.ptr(lp), _lib.ptr(lg), Np, Ng,
                                           _lib.ptr(out["iou"]), _lib.ptr(out["ios"]), _lib.ptr(out["iog"]), _lib.ptr(inter),
                                           _lib.ptr(ap), _lib.ptr(ag), _lib.ptr(ws), ws_bytes, _lib.current_stream()), "mask_overlaps")
    out.update(inter=inter[1:, 1:], area_pred=ap[1:], area_gt=ag[1:])
    return out


def _stack_to_labels(stack):
    s = np.asarray(stack).astype(bool)
    if s.ndim != 4:
        raise ValueError("expected a [N,S,H,W] stack of boolean masks")
    if s.shape[0] > 65535:
        raise ValueError("more than 65535 masks do not fit uint16 labels")
    if s.shape[0] and s.sum(axis=0).max() > 1:
        raise NotImplementedError("overlapping masks inside one stack: the CUDA path works on label volumes (disjoint instances), "
                                  "which is what the reference's callers build their stacks from")
    lab = np.zeros(s.shape[1:], np.uint16)
    for i in range(s.shape[0]):
        lab[s[i]] = i + 1
    return lab


def _stacks(mask_a, mask_b, key):
    a, b = np.asarray(mask_a), np.asarray(mask_b)
    if a.shape[1:] != b.shape[1:]:
        raise IndexError
    r = mask_overlaps_labels(_stack_to_labels(a), _stack_to_labels(b), np.arange(1, a.shape[0] + 1), np.arange(1, b.shape[0] + 1),
                             want=(key,))
    return r[key].cpu().numpy()


def mask_iou_fast(mask_a, mask_b):
    return _stacks(mask_a, mask_b, "iou")


def mask_ios_fast(mask_a, mask_b):
    return _stacks(mask_a, mask_b, "ios")


def mask_iog_fast(mask_a, mask_b):
    return _stacks(mask_a, mask_b, "iog")


mask_iou = mask_iou_fast          # mask_iou.py:10-45, the slow twin of mask_iou_fast


# ------------------------------------------------------------------------------------------------------------------
# Per-volume evaluation records of the two evaluation scripts.  The volume-sized work (ids present in a label volume,
# the instance overlap matrix, the voxel counts) runs on the GPU; the greedy matching over a few hundred rows is the
# reference's own sequential loop and stays on the host.  The records of many volumes are what the ranks all-reduce
# (dist.py) before precision / recall / AP are formed.
def _dev_u16(x):
    import torch
    if hasattr(x, "data_ptr"):
        return x.to(device="cuda", dtype=torch.uint16).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint16)).cuda()


def label_ids(labels):
    """np.unique(labels) without the background id 0 (eval_instance_segmentation_soma.py:177-181), ascending int64."""
    import torch
    t = _dev_u16(labels)
    present = torch.empty(65536, dtype=torch.uint8, device=t.device)
    _lib.check(_lib.lib().b200seg_label_presence_dev(_lib.ptr(t), t.numel(), _lib.ptr(present), _lib.current_stream()), "label_presence")
    ids = torch.nonzero(present).flatten().cpu().numpy().astype(np.int64)
    return ids[ids != 0]


def match_by_iou(iou, iou_thresh):
    """eval_instance_segmentation_soma.py:198-216: predictions in score order, each takes its arg-max ground truth if the IoU
    reaches the threshold and the ground truth is still free -> list of 0 / 1."""
    iou = np.asarray(iou)
    if iou.shape[1] == 0:
        return [0] * iou.shape[0]
    gt_index = iou.argmax(axis=1)
    gt_index[iou.max(axis=1) < iou_thresh] = -1
    taken = np.zeros(iou.shape[1], dtype=bool)
    match = []
    for g in gt_index:
        match.append(int(g >= 0 and not taken[g]))
        if g >= 0:
            taken[g] = True
    return match


def eval_volume_soma(pred_mask, gt_mask, pred_score, iou_thresh=0.3):
    """One image of eval_instance_segmentation_soma.py:156-216: pred_mask / gt_mask uint16 label volumes, pred_score
    [[id, score], ...].  Returns dict(score, match, n_pos): the rows this image appends to the global lists."""
    pred_score = np.asarray(pred_score, dtype=np.float64).reshape(-1, 2)
    pred_score = pred_score[pred_score[:, 1].argsort()[::-1], :]
    p, g = _dev_u16(pred_mask), _dev_u16(gt_mask)
    gt_ids = label_ids(g)
    out = dict(score=pred_score[:, 1].tolist(), match=[], n_pos=int(len(gt_ids)))
    if len(pred_score) == 0:
        return out
    if len(gt_ids) == 0:
        out["match"] = [0] * len(label_ids(p))
        return out
    iou = mask_overlaps_labels(p, g, pred_score[:, 0].astype(np.int64), gt_ids, want=("iou",))["iou"].cpu().numpy()
    out["match"] = match_by_iou(iou, iou_thresh)
    return out


def precision_recall(score, match, n_pos):
    """eval_instance_segmentation_soma.py:238-254 on the concatenated records of all images."""
    score, match = np.asarray(score), np.asarray(match, dtype=np.int8)
    match = match[score.argsort()[::-1]]
    tp, fp = np.cumsum(match == 1), np.cumsum(match == 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = tp / (fp + tp)
    rec = tp / n_pos if n_pos > 0 else None
    return prec, rec


def voc_ap(rec, prec):
    """eval_instance_segmentation_soma.py:18-50 with use_07_metric=False: area under the monotone precision envelope."""
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], np.nan_to_num(prec), [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def match_boxes_tp_fp(dets_bbox, gt_bbox, ovthresh=0.4):
    """evaluation_nuclei_f1score_seg.py:92-131: detections in file order, float64 box IoU with the +1 convention, each takes
    its arg-max ground truth if IoU > ovthresh and that ground truth is still free.  Returns (tp, fp float arrays, matched bool)."""
    dets_bbox, gt_bbox = np.asarray(dets_bbox, dtype=float).reshape(-1, 6), np.asarray(gt_bbox, dtype=float).reshape(-1, 6)
    n = dets_bbox.shape[0]
    tp, fp, matched = np.zeros(n), np.zeros(n), np.zeros(n, dtype=bool)
    if gt_bbox.shape[0] == 0:
        return tp, fp, matched
    detected = np.zeros(gt_bbox.shape[0], dtype=bool)
    gvol = (gt_bbox[:, 3] - gt_bbox[:, 0] + 1.) * (gt_bbox[:, 4] - gt_bbox[:, 1] + 1.) * (gt_bbox[:, 5] - gt_bbox[:, 2] + 1.)
    for ib, b in enumerate(dets_bbox):
        lo = np.maximum(gt_bbox[:, :3], b[:3])
        hi = np.minimum(gt_bbox[:, 3:], b[3:])
        ext = np.maximum(hi - lo + 1., 0.)
        inters = ext[:, 0] * ext[:, 1] * ext[:, 2]
        uni = (b[3] - b[0] + 1.) * (b[4] - b[1] + 1.) * (b[5] - b[2] + 1.) + gvol - inters
        overlaps = inters / uni
        j = int(np.argmax(overlaps))
        if overlaps[j] > ovthresh and not detected[j]:
            tp[ib] = 1.
            detected[j] = True
            matched[ib] = True
        else:
            fp[ib] = 1.
    return tp, fp, matched


def eval_volume_nuclei(pred_mask, gt_mask, dets_bbox, gt_bbox, ovthresh=0.4):
    """One image of evaluation_nuclei_f1score_seg.py:70-131.  Returns dict(tp, fp, tp_pixel, gt_pixel, pre_pixel); like the
    script, tp_pixel stays 0 when the image has no ground-truth box.  Detection boxes are clipped to the volume (the script's
    negative slice bounds would wrap)."""
    import torch
    p, g = _dev_u16(pred_mask), _dev_u16(gt_mask)
    if p.shape != g.shape or p.dim() != 3:
        raise ValueError("eval_volume_nuclei: pred_mask and gt_mask must be equally shaped [S,H,W] volumes")
    S, H, W = (int(v) for v in p.shape)
    dets_bbox = np.asarray(dets_bbox, dtype=float).reshape(-1, 6)
    tp, fp, matched = match_boxes_tp_fp(dets_bbox, gt_bbox, ovthresh)
    boxes = np.ascontiguousarray(dets_bbox[matched].astype(int).astype(np.int32))
    L = _lib.lib()
    counts = torch.zeros(3, dtype=torch.int64, device=p.device)
    ws_bytes = L.b200seg_eval_voxel_counts_workspace_bytes(p.numel())
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
    d_boxes = torch.from_numpy(boxes).to(p.device) if len(boxes) else None
    _lib.check(L.b200seg_eval_voxel_counts_dev(_lib.ptr(p), _lib.ptr(g), S, H, W, _lib.ptr(d_boxes), int(len(boxes)), _lib.ptr(counts),
                                               _lib.ptr(ws), ws_bytes, _lib.current_stream()), "eval_voxel_counts")
    c = counts.cpu().numpy()
    return dict(tp=tp, fp=fp, tp_pixel=int(c[0]), gt_pixel=int(c[1]), pre_pixel=int(c[2]))
