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
