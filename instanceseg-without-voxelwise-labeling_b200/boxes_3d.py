"""3D box operators behind the reference names (lib/utils/boxes_3d.py:55,364-374 and
lib/utils/cython_nms_3d.pyx / cython_bbox_3d.pyx).

numpy in -> numpy out (host entry points of the C ABI do the H2D/D2H), torch CUDA tensors in ->
torch CUDA tensors out (device entry points, current stream, no synchronisation)."""
import ctypes as C

import numpy as np

from . import _lib


def _is_tensor(x):
    return hasattr(x, "data_ptr")


def _nms_numpy(dets, thresh, by_volume):
    if dets.dtype != np.float32:                      # same error as the Cython buffer check
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % dets.dtype)
    if dets.ndim != 2 or dets.shape[1] != 7:
        raise ValueError("dets must be [N,7] (x1,y1,z1,x2,y2,z2,score)")
    dets = np.ascontiguousarray(dets)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    cnt = C.c_int(0)
    _lib.check(_lib.lib().b200seg_nms3d_host(_lib.ptr(dets), n, float(np.float32(thresh)), int(by_volume),
                                             _lib.ptr(keep), C.byref(cnt)), "nms3d_host")
    return keep[:cnt.value].copy()


def nms_3d_batched(dets, offsets, n_max, thresh, by_volume=False, want_rank_order=False):
    """Device batched NMS.  dets [total,7] f32 cuda, offsets [batch+1] int32 cuda.
    Returns (keep int64 [total], keep_count int32 [batch], rank_order int32 [total] | None)."""
    import torch
    L = _lib.lib()
    batch = offsets.numel() - 1
    total = dets.shape[0]
    dev = dets.device
    keep = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    cnt = torch.empty(max(batch, 1), dtype=torch.int32, device=dev)
    rank = torch.empty(max(total, 1), dtype=torch.int32, device=dev) if want_rank_order else None
    ws_bytes = L.b200seg_nms3d_workspace_bytes(batch, n_max)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.b200seg_nms3d_dev(_lib.ptr(dets), _lib.ptr(offsets), batch, n_max, float(np.float32(thresh)),
                                   int(by_volume), _lib.ptr(keep), _lib.ptr(cnt), _lib.ptr(rank), _lib.ptr(ws),
                                   ws_bytes, _lib.current_stream()), "nms3d_dev")
    return keep, cnt, rank


def _nms_tensor(dets, thresh, by_volume):
    import torch
    assert dets.is_cuda and dets.dtype == torch.float32 and dets.dim() == 2 and dets.shape[1] == 7
    dets = dets.contiguous()
    n = dets.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dets.device)
    off = torch.tensor([0, n], dtype=torch.int32, device=dets.device)
    keep, cnt, _ = nms_3d_batched(dets, off, n, thresh, by_volume)
    return keep[:int(cnt[0].item())]


def nms_3d(dets, thresh):
    """Classic greedy 3D NMS (boxes_3d.py:364-368 -> cython_nms_3d.pyx:39-96).
    Returns kept indices, ascending; [] for an empty numpy input, like the reference."""
    if _is_tensor(dets):
        return _nms_tensor(dets, thresh, False)
    if dets.shape[0] == 0:
        return []
    return _nms_numpy(dets, thresh, False)


def nms_3d_volume(dets, thresh):
    """Greedy 3D NMS visiting boxes by descending volume (boxes_3d.py:370-374 -> pyx:102-159)."""
    if _is_tensor(dets):
        return _nms_tensor(dets, thresh, True)
    if dets.shape[0] == 0:
        return []
    return _nms_numpy(dets, thresh, True)


def bbox_overlaps_3d(boxes, query_boxes):
    """IoU matrix [N,K] fp32 (cython_bbox_3d.pyx:32-80; boxes_3d.py:55)."""
    L = _lib.lib()
    if _is_tensor(boxes):
        import torch
        assert boxes.is_cuda and query_boxes.is_cuda and boxes.dtype == torch.float32
        boxes = boxes.contiguous()
        query_boxes = query_boxes.contiguous()
        N, K = boxes.shape[0], query_boxes.shape[0]
        out = torch.empty((N, K), dtype=torch.float32, device=boxes.device)
        _lib.check(L.b200seg_iou3d_dev(_lib.ptr(boxes), N, _lib.ptr(query_boxes), K, _lib.ptr(out),
                                       _lib.current_stream()), "iou3d_dev")
        return out
    for a in (boxes, query_boxes):
        if a.dtype != np.float32:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got '%s'" % a.dtype)
    boxes = np.ascontiguousarray(boxes)
    query_boxes = np.ascontiguousarray(query_boxes)
    N, K = boxes.shape[0], query_boxes.shape[0]
    out = np.zeros((N, K), dtype=np.float32)
    _lib.check(L.b200seg_iou3d_host(_lib.ptr(boxes), N, _lib.ptr(query_boxes), K, _lib.ptr(out)), "iou3d_host")
    return out
