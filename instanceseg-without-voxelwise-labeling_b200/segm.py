"""Mask R-CNN mask paste-back behind the reference's name (lib/core/test.py:886-945, called by core/test.py:118).

segm_results(cls_boxes, masks, ref_boxes, im_s, im_h, im_w) -> cls_segms, a list per class of one uint8 volume per
detection, exactly as the reference returns it.  The per-detection work -- zero-pad the M^3 block, resize it to the
expanded box (skimage.transform.resize -> scipy.ndimage gaussian_filter + zoom arithmetic), threshold, clip to the
volume -- runs in ONE launch of segm_resize_paste_kernel (csrc/segm_paste.cu) for all detections; only the packed crops
cross PCIe, and the host drops them into freshly zeroed volumes (the reference's output format is n full volumes).
`segm_results_device` keeps everything on the GPU (packed crops, or the [n,S,H,W] stack via segm_expand_kernel).

Settings come from core.config.cfg inside the reference tree, else from the keyword arguments
(defaults = lib/core/config.py: MODEL.NUM_CLASSES, MRCNN.RESOLUTION 14, MRCNN.CLS_SPECIFIC_MASK True,
MRCNN.THRESH_BINARIZE 0.5)."""
import numpy as np

from . import _lib

_tables = {}


def _settings(kw):
    try:
        from core.config import cfg
        d = dict(num_classes=cfg.MODEL.NUM_CLASSES, resolution=cfg.MRCNN.RESOLUTION, cls_specific_mask=cfg.MRCNN.CLS_SPECIFIC_MASK,
                 thresh_binarize=cfg.MRCNN.THRESH_BINARIZE)
    except ImportError:
        d = dict(num_classes=None, resolution=None, cls_specific_mask=True, thresh_binarize=0.5)
    d.update({k: v for k, v in kw.items() if v is not None})
    return d


def expand_boxes(boxes, scale):
    """lib/utils/boxes_3d.py:271-292: scale every box about its centre; arithmetic in the dtype of `boxes`, float64 result."""
    b = np.asarray(boxes)
    half = (b[:, 3:6] - b[:, 0:3]) * .5
    ctr = (b[:, 3:6] + b[:, 0:3]) * .5
    half = half * scale
    out = np.zeros(b.shape)
    out[:, 0:3] = ctr - half
    out[:, 3:6] = ctr + half
    return out


def gauss_table(M):
    """The anti-aliasing taps scipy.ndimage.gaussian_filter builds when an axis of M+2 samples is resized to o < M+2
    (skimage: sigma = (in/out - 1)/2; scipy _gaussian_kernel1d with truncate 4.0), row o = taps x = -R..0; see b200seg.h."""
    if M not in _tables:
        M2 = M + 2
        tab = np.zeros((M2, 2 * M2), np.float64)
        for o in range(1, M2):
            sigma = float(max(0, (np.divide(M2, o) - 1) / 2))
            R = int(4.0 * sigma + 0.5)
            if R < 1:
                continue
            x = np.arange(-R, R + 1)
            phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
            phi = phi / phi.sum()
            tab[o, :R + 1] = phi[:R + 1]
        assert tab.size == _lib.lib().b200seg_segm_gauss_table_size(M)
        _tables[M] = tab
    return _tables[M]


def clipped_boxes(ref_boxes_i32, im_s, im_h, im_w):
    """core/test.py:923-928 -> int64 [n,6] (x0,y0,z0,x1,y1,z1), upper bounds exclusive, and the crop byte offsets [n+1]."""
    b = np.asarray(ref_boxes_i32, np.int64).reshape(-1, 6)
    lo = np.maximum(b[:, 0:3], 0)
    hi = np.minimum(b[:, 3:6] + 1, np.array([im_w, im_h, im_s], np.int64))
    ext = np.maximum(hi - lo, 0)
    vol = np.where((ext > 0).all(axis=1), ext.prod(axis=1), 0)
    off = np.zeros(len(b) + 1, np.int64)
    np.cumsum(vol, out=off[1:])
    return np.concatenate([lo, hi], axis=1), off


def _plan(cls_boxes, masks_shape, ref_boxes, s):
    ncls = int(s["num_classes"] if s["num_classes"] is not None else len(cls_boxes))
    n_masks, C, M = int(masks_shape[0]), int(masks_shape[1]), int(masks_shape[2])
    if s["resolution"] is not None and int(s["resolution"]) != M:
        raise ValueError("segm_results: masks are %d^3 but MRCNN.RESOLUTION is %d" % (M, int(s["resolution"])))
    counts = [0] + [int(np.shape(cls_boxes[j])[0]) if len(cls_boxes[j]) else 0 for j in range(1, ncls)]
    if sum(counts) != n_masks:                               # the reference's closing assert (core/test.py:944)
        raise AssertionError("segm_results: %d detections in cls_boxes but %d masks" % (sum(counts), n_masks))
    cls_of = np.concatenate([np.full(c, j, np.int64) for j, c in enumerate(counts)]) if n_masks else np.zeros(0, np.int64)
    mask_index = (np.arange(n_masks, dtype=np.int64) * C + (cls_of if s["cls_specific_mask"] else 0)).astype(np.int32)
    scale = (M + 2.0) / M
    boxes = expand_boxes(np.asarray(ref_boxes).reshape(-1, 6)[:n_masks], scale).astype(np.int32)      # :895-898
    return ncls, counts, M, mask_index, np.ascontiguousarray(boxes)


def segm_results_device(cls_boxes, masks, ref_boxes, im_s, im_h, im_w, expand=False, num_classes=None, resolution=None,
                        cls_specific_mask=None, thresh_binarize=None):
    """GPU-resident form: returns dict(crops uint8 cuda tensor (packed), crop_off int64 numpy [n+1], boxes int64 numpy [n,6]
    clipped (upper bounds exclusive), counts per class, and with expand=True `volumes` uint8 cuda [n,S,H,W])."""
    import torch
    s = _settings(dict(num_classes=num_classes, resolution=resolution, cls_specific_mask=cls_specific_mask, thresh_binarize=thresh_binarize))
    if not torch.is_tensor(masks):
        masks = torch.from_numpy(np.ascontiguousarray(masks, dtype=np.float32))
    if masks.dim() != 5 or masks.shape[2] != masks.shape[3] or masks.shape[3] != masks.shape[4]:
        raise ValueError("segm_results: masks must be [n, C, M, M, M]")
    ncls, counts, M, mask_index, boxes = _plan(cls_boxes, masks.shape, ref_boxes, s)
    n = len(mask_index)
    clip, off = clipped_boxes(boxes, im_s, im_h, im_w)
    L = _lib.lib()
    dev = torch.device("cuda")
    crops = torch.empty(int(off[-1]) + 16, dtype=torch.uint8, device=dev)
    out = dict(crops=crops[:int(off[-1])], crop_off=off, boxes=clip, counts=counts, num_classes=ncls)
    if n == 0:
        if expand:
            out["volumes"] = torch.zeros((0, im_s, im_h, im_w), dtype=torch.uint8, device=dev)
        return out
    d_masks = masks.to(device=dev, dtype=torch.float32).contiguous()
    d_idx = torch.from_numpy(mask_index).to(dev)
    d_boxes = torch.from_numpy(boxes).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_tab = torch.from_numpy(gauss_table(M)).to(dev)
    st = _lib.current_stream()
    _lib.check(L.b200seg_segm_paste_dev(_lib.ptr(d_masks), _lib.ptr(d_idx), _lib.ptr(d_boxes), n, M, _lib.ptr(d_tab),
                                        float(np.float32(s["thresh_binarize"])), int(im_s), int(im_h), int(im_w),
                                        _lib.ptr(crops), _lib.ptr(d_off), st), "b200seg_segm_paste_dev")
    if expand:
        vols = torch.empty((n, im_s, im_h, im_w), dtype=torch.uint8, device=dev)
        _lib.check(L.b200seg_segm_expand_dev(_lib.ptr(crops), _lib.ptr(d_off), _lib.ptr(d_boxes), n, int(im_s), int(im_h), int(im_w),
                                             _lib.ptr(vols), st), "b200seg_segm_expand_dev")
        out["volumes"] = vols
    return out


def segm_results(cls_boxes, masks, ref_boxes, im_s, im_h, im_w, num_classes=None, resolution=None, cls_specific_mask=None,
                 thresh_binarize=None):
    """The reference's signature and return value (numpy in, list of lists of uint8 volumes out), through the numpy seam
    b200seg_segm_paste_host.  Boxes that miss the volume give an all-zero volume."""
    s = _settings(dict(num_classes=num_classes, resolution=resolution, cls_specific_mask=cls_specific_mask, thresh_binarize=thresh_binarize))
    masks = np.ascontiguousarray(masks.detach().cpu().numpy() if hasattr(masks, "detach") else masks, dtype=np.float32)
    if masks.ndim != 5 or masks.shape[2] != masks.shape[3] or masks.shape[3] != masks.shape[4]:
        raise ValueError("segm_results: masks must be [n, C, M, M, M]")
    ncls, counts, M, mask_index, boxes = _plan(cls_boxes, masks.shape, ref_boxes, s)
    n = len(mask_index)
    clip, off = clipped_boxes(boxes, im_s, im_h, im_w)
    crops = np.empty(int(off[-1]), np.uint8)
    if n:
        L = _lib.lib()
        tab = gauss_table(M)
        _lib.check(L.b200seg_segm_paste_host(_lib.ptr(masks), masks.shape[0] * masks.shape[1], _lib.ptr(mask_index), _lib.ptr(boxes), n, M,
                                             _lib.ptr(tab), float(np.float32(s["thresh_binarize"])), int(im_s), int(im_h), int(im_w),
                                             _lib.ptr(crops), _lib.ptr(off)), "b200seg_segm_paste_host")
    cls_segms = [[] for _ in range(ncls)]
    d = 0
    for j in range(1, ncls):
        segms = []
        for _ in range(counts[j]):
            im_mask = np.zeros((im_s, im_h, im_w), dtype=np.uint8)
            x0, y0, z0, x1, y1, z1 = clip[d]
            if off[d + 1] > off[d]:
                im_mask[z0:z1, y0:y1, x0:x1] = crops[off[d]:off[d + 1]].reshape(z1 - z0, y1 - y0, x1 - x0)
            segms.append(im_mask)
            d += 1
        cls_segms[j] = segms
    return cls_segms
