"""Per-volume post-processing of tools/binarization_nuclei.py:57-154 on the GPU.

The script's work splits at the point where it reads the per-instance PRM tifs from disk:
  1. `nuclei_select(dets, width, nms_thresh)`  -- :72-87: edge filter (host logic on the detection rows), 3D NMS ordered by
     box volume (nms3d.cu through boxes_3d.nms_3d_volume), score > 0.4 -> indices of the instances to visit, in visit order;
  2. `nuclei_boxes(dets, tile_off, norm_side, slices)` -- :96-106: tile-relative int() truncation, clamp to the tile, back
     to volume coordinates;
  3. `binarize_nuclei_host(volume, boxes, prm_crops)` -- :98-149 for all visited instances in one C call
     (b200seg_binarize_nuclei_host: crop + normalise, 2D-Otsu, largest component, hole filling, closing, label paste,
     survivor test), numpy in / numpy out; `binarize_nuclei` is the same on cuda tensors.
`id_det_rows` builds the [mask_id, x1, y1, z1, x2, y2, z2, score] table the script saves (:148-151)."""
import numpy as np

from . import _lib
from .binarization import crop_offsets


def nuclei_edge_filter(dets, width):
    """:72-77 ("remove broken boxes at edges"), reproduced as written -- including `width` in the y condition."""
    d = np.asarray(dets)
    c1 = (d[:, 0] > 10) & (d[:, 3] < width - 10) & ((d[:, 3] - d[:, 0] + 1) < 32)
    c2 = (d[:, 1] > 10) & (d[:, 4] < width - 10) & ((d[:, 4] - d[:, 1] + 1) < 32)
    return ~(c1 | c2)


def nuclei_select(dets, width, nms_thresh=0.15, score_thresh=0.4):
    """Indices into `dets` of the instances the script visits, in its visit order (:72-87)."""
    from .boxes_3d import nms_3d_volume
    dets = np.asarray(dets)
    idx = np.nonzero(nuclei_edge_filter(dets, width))[0]
    d = np.ascontiguousarray(dets[idx], dtype=np.float32)
    keep = np.asarray(nms_3d_volume(d, nms_thresh), dtype=np.int64)
    idx, d = idx[keep], d[keep]
    return idx[d[:, -1] > score_thresh]


def nuclei_boxes(dets, tile_off, norm_side, slices):
    """:96-106: det - tile offset, astype(int), clamp to [0, norm_side-1] (x, y) / [0, slices-1] (z); returned in volume
    coordinates (tile offset added back, :147) as int32 [n,6].  tile_off [n,3] = (w, h, s) of each instance's tile."""
    d = np.asarray(dets)
    t = np.asarray(tile_off, dtype=np.int64).reshape(-1, 3)
    rel = (d[:, :6] - np.concatenate([t, t], axis=1)).astype(int)
    hi = np.array([norm_side - 1, norm_side - 1, slices - 1], dtype=np.int64)
    rel[:, 0:3] = np.maximum(0, rel[:, 0:3])
    rel[:, 3:6] = np.minimum(hi, rel[:, 3:6])
    return (rel + np.concatenate([t, t], axis=1)).astype(np.int32)


def _pack(prm_crops):
    if isinstance(prm_crops, np.ndarray) and prm_crops.ndim == 1:
        return np.ascontiguousarray(prm_crops, dtype=np.uint8)
    return np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint8).ravel() for p in prm_crops]) if len(prm_crops) else
                                np.zeros(0, np.uint8))


def binarize_nuclei_host(volume, boxes, prm_crops, seg_out=None, want_masks=False):
    """volume uint8 / uint16 [S,H,W] (prefiltered, :43-44), boxes int32 [n,6] volume coordinates, prm_crops: list of
    box-shaped uint8 arrays or one packed uint8 array.  Returns dict(seg uint16, status, b_max, survive [, masks])."""
    volume = np.ascontiguousarray(volume)
    if volume.dtype not in (np.uint8, np.uint16):
        raise TypeError("binarize_nuclei_host: volume must be uint8 or uint16")
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 6)
    n = boxes.shape[0]
    prm = _pack(prm_crops)
    off = crop_offsets(boxes)
    if prm.size != off[-1]:
        raise ValueError("binarize_nuclei_host: PRM crops hold %d voxels, the boxes %d" % (prm.size, off[-1]))
    S, H, W = volume.shape
    seg = seg_out if seg_out is not None else np.empty((S, H, W), np.uint16)
    nn = max(n, 1)
    b_max, status, survive = np.zeros(nn, np.int32), np.zeros(nn, np.int32), np.zeros(nn, np.uint8)
    masks = np.zeros(max(int(off[-1]), 1), np.uint8) if want_masks else None
    _lib.check(_lib.lib().b200seg_binarize_nuclei_host(_lib.ptr(volume), volume.dtype.itemsize, S, H, W, _lib.ptr(boxes), _lib.ptr(prm),
                                                       _lib.ptr(off), n, _lib.ptr(seg), _lib.ptr(masks), _lib.ptr(b_max), _lib.ptr(status),
                                                       _lib.ptr(survive)), "b200seg_binarize_nuclei_host")
    if (status[:n] == 4).any():
        # the generic 2D-Otsu kernel holds at most 2048 gray levels in shared memory; the reference would binarize such a crop
        # (uint16 data whose range exceeds 2048 levels inside one box): refuse loudly instead of dropping the instance
        raise _lib.B200SegError("binarize_nuclei_host: instance(s) %s span more than 2048 gray levels (unsupported by the 2D-Otsu kernel)"
                                % np.flatnonzero(status[:n] == 4).tolist())
    out = dict(seg=seg, status=status[:n], b_max=b_max[:n], survive=survive[:n].astype(bool), crop_off=off)
    if want_masks:
        out["masks"] = masks[:int(off[-1])]
    return out


def binarize_nuclei(volume, boxes, prm, crop_off):
    """Device form: volume uint8 / uint16 cuda [S,H,W], boxes int32 cuda [n,6], prm uint8 cuda (packed), crop_off int64 cuda
    [n+1].  Returns (seg uint16 cuda, masks uint8 cuda, b_max, status int32 cuda, survive uint8 cuda); no synchronisation -- the
    caller must check `status`: 4 = more than 2048 gray levels in the box (unsupported, nothing pasted)."""
    import torch
    S, H, W = volume.shape
    n = int(boxes.shape[0])
    total = int(prm.numel())
    dev = volume.device
    L = _lib.lib()
    seg = torch.empty((S, H, W), dtype=torch.uint16, device=dev)
    masks = torch.empty(max(total, 1) + 16, dtype=torch.uint8, device=dev)
    b_max = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    status = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    survive = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
    ws_bytes = L.b200seg_binarize_nuclei_workspace_bytes(n, total, S, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.b200seg_binarize_nuclei_dev(_lib.ptr(volume), volume.element_size(), S, H, W, _lib.ptr(boxes), _lib.ptr(prm), _lib.ptr(crop_off),
                                             n, total, _lib.ptr(seg), _lib.ptr(masks), _lib.ptr(b_max), _lib.ptr(status), _lib.ptr(survive),
                                             _lib.ptr(ws), ws_bytes, _lib.current_stream()), "b200seg_binarize_nuclei_dev")
    return seg, masks[:total], b_max[:n], status[:n], survive[:n]


def id_det_rows(boxes, scores, survive):
    """:148-151: one row [mask_id, x1, y1, z1, x2, y2, z2, score] per instance whose label survives in the volume.  float64 like
    the array the script saves (np.concatenate promotes its float32 accumulator on the first row; the score is the float32
    detection score widened)."""
    boxes, survive = np.asarray(boxes, dtype=np.float64).reshape(-1, 6), np.asarray(survive, dtype=bool)
    scores = np.asarray(scores, dtype=np.float32).astype(np.float64)
    ids = np.arange(1, len(boxes) + 1, dtype=np.float64)
    if not survive.any():
        return np.zeros((0, 8), dtype=np.float32)              # the script's untouched accumulator
    return np.concatenate([ids[:, None], boxes, scores[:, None]], axis=1)[survive]
