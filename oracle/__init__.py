"""CPU oracle for the b200seg hot path.  TEST INFRASTRUCTURE ONLY.

`oracle/oracle.c` restates the reference algorithms (each function cites the reference
file:line it follows); this module builds it with gcc and wraps it with numpy signatures that
mirror the reference call sites.  `oracle/_ref/` holds the reference's own Cython/CUDA code built
by `oracle/build_ref.py` (validation of the restatement, and `cpu_baseline.kind == "reference"`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this package -- as the checker or the reported CPU baseline, never as the product.
"""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build(force=False):
    """gcc-compile oracle.c -> liboracle.so (baseline x86-64, no FMA contraction)."""
    src = os.path.join(HERE, "oracle.c")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= os.path.getmtime(src)):
        return LIB_PATH
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99",
                           src, "-o", LIB_PATH, "-lm"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        c = ctypes
        L.oracle_nms_3d.restype = c.c_long
        L.oracle_nms_3d.argtypes = [c.c_void_p, c.c_long, c.c_float, c.c_int, c.c_void_p]
        L.oracle_argsort_desc.restype = None
        L.oracle_argsort_desc.argtypes = [c.c_void_p, c.c_long, c.c_void_p]
        L.oracle_bbox_overlaps_3d.restype = None
        L.oracle_bbox_overlaps_3d.argtypes = [c.c_void_p, c.c_long, c.c_void_p, c.c_long, c.c_void_p]
        for name in ("oracle_roialign3d_fwd", "oracle_roialign3d_bwd"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_int,
                          c.c_int, c.c_int, c.c_int, c.c_float, c.c_int, c.c_void_p]
        L.oracle_peak_stimulation.restype = c.c_long
        L.oracle_peak_stimulation.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int,
                                              c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_long,
                                              c.c_void_p, c.c_void_p]
        L.oracle_otsu_2d_fast.restype = c.c_int
        L.oracle_otsu_2d_fast.argtypes = [c.c_void_p, c.c_void_p, c.c_long, c.c_void_p,
                                          c.POINTER(c.c_int), c.c_void_p, c.POINTER(c.c_int)]
        L.oracle_paste_labels.restype = None
        L.oracle_paste_labels.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p,
                                          c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ------------------------------------------------------------------------------- boxes
def nms_3d(dets, thresh, by_volume=False):
    """cython_nms_3d.pyx:39-96 / :102-159 -> int64 kept indices, ascending."""
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n = dets.shape[0]
    keep = np.empty(max(n, 1), dtype=np.int64)
    m = lib().oracle_nms_3d(_p(dets), n, np.float32(thresh), int(by_volume), _p(keep))
    return keep[:m].copy()


def nms_3d_volume(dets, thresh):
    return nms_3d(dets, thresh, by_volume=True)


def argsort_desc(keys):
    keys = np.ascontiguousarray(keys, dtype=np.float32)
    out = np.empty(keys.shape[0], dtype=np.int64)
    lib().oracle_argsort_desc(_p(keys), keys.shape[0], _p(out))
    return out


def bbox_overlaps_3d(boxes, query_boxes):
    """cython_bbox_3d.pyx:32-80 -> [N,K] fp32."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    query_boxes = np.ascontiguousarray(query_boxes, dtype=np.float32)
    out = np.empty((boxes.shape[0], query_boxes.shape[0]), dtype=np.float32)
    lib().oracle_bbox_overlaps_3d(_p(boxes), boxes.shape[0], _p(query_boxes), query_boxes.shape[0], _p(out))
    return out


# ------------------------------------------------------------------------------- RoIAlign3D
def roialign3d_fwd(features, rois, P, scale, sr):
    """roi_align_kernel_3d.cu:81-151; returns [R,C,P,P,P] holding the reference's (H,W,S) bin order."""
    features = np.ascontiguousarray(features, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    B, C, S, H, W = features.shape
    Ps, Ph, Pw = (P, P, P) if np.isscalar(P) else P
    out = np.empty((rois.shape[0], C, Ps, Ph, Pw), dtype=np.float32)
    lib().oracle_roialign3d_fwd(_p(features), B, C, S, H, W, _p(rois), rois.shape[0], Ps, Ph, Pw,
                                np.float32(scale), int(sr), _p(out))
    return out


def roialign3d_bwd(grad_out, rois, feat_shape, scale, sr):
    """roi_align_kernel_3d.cu:238-338 with sequential (deterministic) adds."""
    grad_out = np.ascontiguousarray(grad_out, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    B, C, S, H, W = feat_shape
    _, _, Ps, Ph, Pw = grad_out.shape
    gi = np.empty(feat_shape, dtype=np.float32)
    lib().oracle_roialign3d_bwd(_p(grad_out), B, C, S, H, W, _p(rois), rois.shape[0], Ps, Ph, Pw,
                                np.float32(scale), int(sr), _p(gi))
    return gi


# ------------------------------------------------------------------------------- peaks
def peak_stimulation_3d(inp, win_size=3, filter_mode="median", thresholds=None, return_aggregation=True):
    """peak_stimulation_3d.py:9-41.  filter_mode: None | 'median' | 'given' (thresholds [B,A])."""
    inp = np.ascontiguousarray(inp, dtype=np.float32)
    B, A, S, H, W = inp.shape
    mode = {None: 0, "none": 0, "median": 1, "given": 2}[filter_mode]
    thr_in = None
    if mode == 2:
        thr_in = np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, dtype=np.float32).reshape(-1), (B * A,)))
    cap = max(int(inp.size), 1)
    cap = min(cap, 1 << 22)
    peaks = np.empty((cap, 5), dtype=np.int64)
    agg = np.empty((B, A), dtype=np.float32)
    thr = np.empty((B, A), dtype=np.float32)
    n = lib().oracle_peak_stimulation(_p(inp), B, A, S, H, W, int(win_size), mode, _p(thr_in),
                                      _p(peaks), cap, _p(agg), _p(thr))
    if n > cap:
        peaks = np.empty((n, 5), dtype=np.int64)
        lib().oracle_peak_stimulation(_p(inp), B, A, S, H, W, int(win_size), mode, _p(thr_in),
                                      _p(peaks), n, _p(agg), _p(thr))
    peaks = peaks[:n].copy()
    if return_aggregation:
        return peaks, agg, thr
    return peaks


# ------------------------------------------------------------------------------- Otsu
class OtsuNoThreshold(UnboundLocalError):
    """The reference raises NameError/UnboundLocalError on k_max when no b wins (otsu.py:277)."""


def otsu_py_2d_fast(image, prm, want_hist=False):
    """tools/otsu.py:199-284 -> (uint8 mask {0,255}, k_max=-1, b_max[, hist.T])."""
    shape = image.shape
    img = np.ascontiguousarray(image, dtype=np.uint16).ravel()
    pr = np.ascontiguousarray(prm, dtype=np.uint16).ravel()
    mask = np.empty(img.size, dtype=np.uint8)
    b = ctypes.c_int(0)
    G = ctypes.c_int(0)
    hist = None
    if want_hist:
        g = int(img.max()) - int(img.min()) + 1
        hist = np.empty((g, g), dtype=np.float64)
    st = lib().oracle_otsu_2d_fast(_p(img), _p(pr), img.size, _p(mask), ctypes.byref(b), _p(hist), ctypes.byref(G))
    if st == -1:
        raise ValueError("gray range too large for the oracle")
    if st == 1:
        raise OtsuNoThreshold("local variable 'k_max' referenced before assignment")
    out = (mask.reshape(shape), -1, int(b.value))
    return out + (hist,) if want_hist else out


def soma_normalise(box_img, box_prm):
    """binarization_soma.py:85-91: image -> uint16 in [30,330], prm -> uint16 round(prm/max*300+30)."""
    gray_max = np.max(box_img)
    f = box_img.astype(float)
    with np.errstate(all="ignore"):
        f = np.clip(f / gray_max * 300, 0, 300) + 30
        img = f.astype(np.uint16)
        p = box_prm.astype(float)
        p = np.round(p / np.max(p) * 300 + 30).astype(np.uint16)
    return img, p


# ------------------------------------------------------------------------------- paste
def paste_labels(seg, boxes, ids, masks):
    """binarization_soma.py:100-104.  seg uint16 [S,H,W] updated in place; boxes int [n,6]
    (x1,y1,z1,x2,y2,z2 inclusive); masks = list of uint8/bool crops [sz,sy,sx].  Returns survive[n]."""
    assert seg.dtype == np.uint16 and seg.flags.c_contiguous
    n = len(masks)
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(n, 6)
    ids = np.ascontiguousarray(ids, dtype=np.uint16)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, m in enumerate(masks):
        b = boxes[i]
        assert m.shape == (b[5] - b[2] + 1, b[4] - b[1] + 1, b[3] - b[0] + 1), (m.shape, b)
        off[i + 1] = off[i] + m.size
    flat = np.empty(max(int(off[-1]), 1), dtype=np.uint8)
    for i, m in enumerate(masks):
        flat[off[i]:off[i + 1]] = (np.asarray(m) != 0).ravel()
    S, H, W = seg.shape
    surv = np.zeros(max(n, 1), dtype=np.uint8)
    lib().oracle_paste_labels(_p(seg), S, H, W, n, _p(boxes), _p(ids), _p(flat), _p(off), _p(surv))
    return surv[:n].astype(bool)


# ------------------------------------------------------------------------------- largest connected component
def largest_cc(mask):
    """binarization_soma.py:97-99: `labels = label(box_bi)` (skimage.measure.label, default = full connectivity)
    then `labels == argsort(bincount(labels.flat)[1:])[-1] + 1`.  PARITY UNPINNED against skimage (not installed
    in this image, unversioned in the reference README): restated with scipy.ndimage.label and a 3x3x3 structuring
    element -- both libraries number the components in raster order of their first voxel.  Tie rule (numpy's
    default argsort is not stable): equal sizes -> the higher label, i.e. what a stable argsort()[-1] returns.
    Returns the boolean mask of the largest component; raises IndexError on an all-background mask like the
    reference does."""
    from scipy import ndimage
    fg = np.asarray(mask) != 0
    labels, n = ndimage.label(fg, structure=np.ones((3,) * fg.ndim, dtype=bool))
    sizes = np.bincount(labels.ravel())[1:]
    if sizes.size == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    best = np.argsort(sizes, kind="stable")[-1] + 1
    return labels == best


# ------------------------------------------------------------------------------- RLE codec
def binary_mask_to_rle(binary_mask):
    """lib/utils/mask_3d.py:15-46 / cython_mask_3d.pyx:19-50: Fortran-order ravel, alternating run lengths starting
    with the zeros (0 when the first element is set); blank mask -> [size]."""
    m = np.asfortranarray(np.asarray(binary_mask) != 0)
    flat = np.ravel(m, order="F").astype(np.int8)
    n = flat.size
    if not flat.any():
        return {"counts": [int(n)], "size": list(m.shape)}
    change = np.flatnonzero(np.diff(flat)) + 1               # first index of every run but the first
    bounds = np.concatenate([[0], change, [n]])
    counts = np.diff(bounds).tolist()
    if flat[0]:
        counts.insert(0, 0)
    return {"counts": [int(c) for c in counts], "size": list(m.shape)}


def rle_to_binary_mask(rle):
    """lib/utils/mask_3d.py:48-71: runs alternate 0,1,...; a single count is a blank mask; Fortran-order reshape."""
    counts, size = list(rle["counts"]), list(rle["size"])
    assert sum(counts) == int(np.prod(size))
    flat = np.zeros(int(np.prod(size)), dtype=np.uint8)
    if len(counts) > 1:
        pos, val = 0, 0
        for c in counts:
            flat[pos:pos + c] = val
            pos += c
            val ^= 1
    return flat.reshape(size, order="F")


# ------------------------------------------------------------------------------- mask overlaps (evaluation)
def mask_overlaps(mask_a, mask_b):
    """tools/evaluation/mask_iou.py:49-109: for stacks [N,...] and [K,...] of boolean masks, iou = I/U, ios = I/A,
    iog = I/B with float (fp64) accumulators stored into float32 arrays.  Returns (iou, ios, iog)."""
    a = np.asarray(mask_a).astype(bool).reshape(len(mask_a), -1)
    b = np.asarray(mask_b).astype(bool).reshape(len(mask_b), -1)
    inter = a.astype(np.int64) @ b.astype(np.int64).T
    A = a.sum(axis=1).astype(np.float64)[:, None]
    B = b.sum(axis=1).astype(np.float64)[None, :]
    I = inter.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (I / (A + B - I)).astype(np.float32), (I / A).astype(np.float32), (I / B).astype(np.float32)


# ------------------------------------------------------------------------------- whole-volume prefilters
def _reflect_pad_last(x, r):
    """scipy 'reflect' (d c b a | a b c d | d c b a) along the last axis, any radius."""
    n = x.shape[-1]
    idx = np.arange(-r, n + r)
    if n == 1:
        idx[:] = 0
    else:
        idx = np.mod(idx, 2 * n)
        idx = np.where(idx < n, idx, 2 * n - 1 - idx)
    return x[..., idx]


def gaussian_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d, order 0 (the arithmetic behind binarization_nuclei.py:43)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return phi / phi.sum(), radius


def gaussian_filter(img, sigma=1, truncate=4.0):
    """scipy.ndimage.gaussian_filter(img, sigma) on an integer volume (tools/binarization_nuclei.py:43), restated:
    axis 0, 1, 2 in turn; per output  acc = w[R]*x0; for j = R..1: acc += (x[-j] + x[+j]) * w[R-j]  in fp64 (separate
    multiply and add), the result cast (truncated) to img.dtype after EVERY pass (scipy filters into the output array)."""
    w, R = gaussian_kernel1d(sigma, truncate)
    out = np.asarray(img)
    if R < 1:
        return out.copy()
    for axis in range(out.ndim):
        x = np.moveaxis(out, axis, -1).astype(np.float64)
        n = x.shape[-1]
        xp = _reflect_pad_last(x, R)
        acc = w[R] * xp[..., R:R + n]
        for jj in range(-R, 0):
            acc = acc + (xp[..., R + jj:R + jj + n] + xp[..., R - jj:R - jj + n]) * w[jj + R]
        out = np.moveaxis(acc, -1, axis).astype(img.dtype)
    return np.ascontiguousarray(out)


def median_filter3(img):
    """scipy.ndimage.median_filter(img, size=3) (tools/binarization_nuclei.py:44): rank 13 of the 3x3x3 neighbourhood,
    'reflect' borders."""
    from numpy.lib.stride_tricks import sliding_window_view
    a = np.asarray(img)
    ap = np.pad(a, 1, mode="symmetric")
    win = sliding_window_view(ap, (3, 3, 3)).reshape(a.shape + (27,))
    return np.ascontiguousarray(np.partition(win, 13, axis=-1)[..., 13])


def zscore_norm(im):
    """tools/infer_simple.py:180-183: (im - mean(im[im>0])) / std(im[im>0]); returns (float64 array, mean, std)."""
    im = np.asarray(im)
    mask = im > 0
    mean_val = np.mean(im[mask])
    std_val = np.std(im[mask])
    return (im - mean_val) / std_val, float(mean_val), float(std_val)


def prm_to_uint8(prm):
    """tools/infer_simple.py:233-238, per channel, in place on a float32 copy: -= min, /= max, *= 255., astype(uint8)."""
    prm = np.array(prm, dtype=np.float32, copy=True)
    out = np.empty(prm.shape, np.uint8)
    for ch in range(prm.shape[0]):
        fm = prm[ch, :]
        fm -= np.min(fm)
        fm /= np.max(fm)
        fm *= 255.
        out[ch] = fm.astype(np.uint8)
    return out


# ------------------------------------------------------------------------------- RPN proposal generation
BBOX_XFORM_CLIP = np.float32(np.log(1000. / 16.))            # lib/core/config.py:947, used at float32 (see below)


def shifted_anchors(anchors, S, H, W, feat_stride):
    """generate_proposals_3d.py:66-88: cell anchors [A,6] + (x,y,z,x,y,z) shifts, rows ordered (S,H,W,A)."""
    sx = np.arange(0, W) * feat_stride
    sy = np.arange(0, H) * feat_stride
    sz = np.arange(0, S) * feat_stride
    zz, yy, xx = np.meshgrid(sz, sy, sx, indexing="ij")
    shifts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    return (np.asarray(anchors, np.float64)[None, :, :] + shifts[:, None, :]).reshape(-1, 6)


def bbox_transform_3d(boxes, deltas):
    """lib/utils/boxes_3d.py:168-225 with unit weights, every operation in float32 (the arithmetic of the reference's
    numpy 1.x era: under numpy >= 2 the float64 scalar cfg.BBOX_XFORM_CLIP would silently promote dw/dh/ds to float64)."""
    b = np.asarray(boxes).astype(np.float32)
    d = np.asarray(deltas, np.float32)
    one, half = np.float32(1.0), np.float32(0.5)
    w = b[:, 3] - b[:, 0] + one
    h = b[:, 4] - b[:, 1] + one
    s = b[:, 5] - b[:, 2] + one
    cx, cy, cz = b[:, 0] + half * w, b[:, 1] + half * h, b[:, 2] + half * s
    dw, dh, ds = (np.minimum(d[:, k], BBOX_XFORM_CLIP) for k in (3, 4, 5))
    pcx, pcy, pcz = d[:, 0] * w + cx, d[:, 1] * h + cy, d[:, 2] * s + cz
    pw, ph, ps = np.exp(dw) * w, np.exp(dh) * h, np.exp(ds) * s
    out = np.zeros(d.shape, np.float32)
    out[:, 0], out[:, 1], out[:, 2] = pcx - half * pw, pcy - half * ph, pcz - half * ps
    out[:, 3], out[:, 4], out[:, 5] = pcx + half * pw - one, pcy + half * ph - one, pcz + half * ps - one
    return out


def generate_proposals(scores, deltas, im_info, anchors, feat_stride, pre_nms_topN, post_nms_topN, nms_thresh, min_size=0):
    """GenerateProposalsOp_3d.proposals_for_one_image (generate_proposals_3d.py:104-177) + _filter_boxes_3d (:180-192).
    scores [A,S,H,W], deltas [6A,S,H,W] float32, im_info [slices, height, width, scale].  Ties between equal scores go to
    the smaller (S,H,W,A) index (the reference's argpartition / argsort are unstable there).
    Returns (proposals [n,6] f32, scores [n] f32, index into the flattened (S,H,W,A) score map [n] int64)."""
    scores, deltas = np.asarray(scores, np.float32), np.asarray(deltas, np.float32)
    im_info = np.asarray(im_info, np.float32)
    A, S, H, W = scores.shape
    sc = scores.transpose(1, 2, 3, 0).reshape(-1)
    dl = deltas.reshape(A, 6, S, H, W).transpose(2, 3, 4, 0, 1).reshape(-1, 6)
    order = np.argsort(-sc, kind="stable")
    if 0 < pre_nms_topN < sc.size:
        order = order[:pre_nms_topN]
    prop = bbox_transform_3d(shifted_anchors(anchors, S, H, W, feat_stride)[order], dl[order])
    lim = (im_info[2] - np.float32(1), im_info[1] - np.float32(1), im_info[0] - np.float32(1))
    for c in range(6):                                                   # clip_tiled_boxes_3d, boxes_3d.py:144-164
        prop[:, c] = np.maximum(np.minimum(prop[:, c], lim[c % 3]), np.float32(0))
    ss = prop[:, 3] - prop[:, 0] + np.float32(1)                         # the reference tests the x extent only (:186-192)
    half_ss = ss / np.float32(2)
    keep = np.where((ss >= np.float32(min_size) * im_info[3]) & (prop[:, 0] + half_ss < im_info[2]) &
                    (prop[:, 1] + half_ss < im_info[1]) & (prop[:, 2] + half_ss < im_info[0]))[0]
    prop, sck, order = prop[keep], sc[order][keep], order[keep]
    if nms_thresh > 0:
        k = nms_3d(np.hstack([prop, sck[:, None]]).astype(np.float32), nms_thresh)
        if post_nms_topN > 0:
            k = k[:post_nms_topN]
        prop, sck, order = prop[k], sck[k], order[k]
    return prop, sck, order.astype(np.int64)


def box_results_with_nms_and_limit(scores, boxes, scores_keep_idx, num_classes, score_thresh, nms, detections_per_im):
    """lib/core/test.py:806-878 (RPN_ONLY False, soft-NMS and box voting off as in every shipped config).  The limit step
    selects rows of cls_keep_idx along axis 0 (the reference's `[keep, :]` only works for 2-D index arrays)."""
    cls_boxes = [[] for _ in range(num_classes)]
    cls_keep_idx = [[] for _ in range(num_classes)]
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] > score_thresh)[0]
        dets_j = np.hstack((boxes[inds, j * 6:(j + 1) * 6], scores[inds, j][:, None])).astype(np.float32, copy=False)
        keep = nms_3d(dets_j, nms)
        cls_boxes[j] = dets_j[keep, :]
        if scores_keep_idx is not None:
            cls_keep_idx[j] = scores_keep_idx[inds][keep]
    if detections_per_im > 0:
        image_scores = np.hstack([cls_boxes[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > detections_per_im:
            image_thresh = np.sort(image_scores)[-detections_per_im]
            for j in range(1, num_classes):
                keep = np.where(cls_boxes[j][:, -1] >= image_thresh)[0]
                cls_boxes[j] = cls_boxes[j][keep, :]
                if scores_keep_idx is not None:
                    cls_keep_idx[j] = cls_keep_idx[j][keep]
    im_results = np.vstack([cls_boxes[j] for j in range(1, num_classes)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes, cls_keep_idx


# ------------------------------------------------------------------------------- Mask R-CNN mask paste-back
def _mirror_idx(idx, n):
    """scipy 'mirror' (d c b | a b c d | c b a) index map for any integer offset."""
    idx = np.asarray(idx)
    if n == 1:
        return np.zeros_like(idx)
    m = np.mod(idx, 2 * n - 2)
    return np.where(m < n, m, 2 * n - 2 - m)


def gaussian_filter_mirror_f32(a, sigmas, truncate=4.0):
    """scipy.ndimage.gaussian_filter(a float32, sigmas, mode='mirror') restated: axes in turn, skipped where sigma <= 1e-15,
    symmetric correlate1d in fp64 (see gaussian_filter above), float32 stored after every pass."""
    out = np.asarray(a, np.float32)
    for axis in range(out.ndim):
        sg = float(sigmas[axis])
        if sg <= 1e-15:
            continue
        w, R = gaussian_kernel1d(sg, truncate)
        x = np.moveaxis(out, axis, -1).astype(np.float64)
        n = x.shape[-1]
        xp = x[..., _mirror_idx(np.arange(-R, n + R), n)]
        acc = w[R] * xp[..., R:R + n]
        for jj in range(-R, 0):
            acc = acc + (xp[..., R + jj:R + jj + n] + xp[..., R - jj:R - jj + n]) * w[jj + R]
        out = np.moveaxis(acc, -1, axis).astype(np.float32)
    return np.ascontiguousarray(out)


def _zoom_axis_taps(n_in, n_out):
    """scipy NI_ZoomShift with grid_mode: cc = (o + .5) * (n_in / n_out) - .5, map_coordinate('mirror'), order-1 taps
    floor(cc), floor(cc)+1 (mirrored) with weights w0 = 1 - frac, w1 = 1 - w0."""
    zoom = np.float64(n_in) / np.float64(n_out)
    cc = np.arange(n_out, dtype=np.float64)
    cc = cc + 0.5
    cc = cc * zoom
    cc = cc - 0.5
    s2 = 2 * n_in - 2
    c = cc.copy()
    if n_in <= 1:
        c[:] = 0
    else:
        neg = c < 0
        cn = c[neg]
        cn = s2 * np.trunc(-cn / s2) + cn
        c[neg] = np.where(cn <= 1 - n_in, cn + s2, -cn)
        big = c > n_in - 1
        cb = c[big]
        cb = cb - s2 * np.trunc(cb / s2)
        c[big] = np.where(cb >= n_in, s2 - cb, cb)
    f = np.floor(c)
    w0 = 1.0 - (c - f)
    w1 = 1.0 - w0
    i0 = f.astype(np.int64)
    return _mirror_idx(i0, n_in), _mirror_idx(i0 + 1, n_in), w0, w1


def zoom_linear_mirror_f32(a, out_shape):
    """scipy.ndimage.zoom(a float32, order=1, mode='mirror', grid_mode=True) to out_shape, restated: the eight taps in
    z-major order, each  ((v * wz) * wy) * wx  in fp64, summed in that order, rounded to float32 once."""
    tabs = [_zoom_axis_taps(a.shape[d], out_shape[d]) for d in range(3)]
    a64 = np.asarray(a, np.float32).astype(np.float64)
    t = np.zeros(out_shape, np.float64)
    for tz in range(2):
        for ty in range(2):
            for tx in range(2):
                coeff = a64[np.ix_(tabs[0][tz], tabs[1][ty], tabs[2][tx])]
                coeff = coeff * tabs[0][2 + tz][:, None, None]
                coeff = coeff * tabs[1][2 + ty][None, :, None]
                coeff = coeff * tabs[2][2 + tx][None, None, :]
                t = t + coeff
    return t.astype(np.float32)


def resize_reflect_antialias(image, out_shape):
    """skimage.transform.resize(image float32, out_shape, mode='reflect', anti_aliasing=True) (order 1, clip=True) as the
    reference calls it (lib/core/test.py:919).  scikit-image is an un-vendored, unversioned dependency of the reference
    (README.md:15) and is not installed here: PARITY UNPINNED against skimage itself.  What IS pinned: skimage's resize is
    a thin wrapper over scipy.ndimage -- gaussian_filter(sigma=max(0,(in/out-1)/2), mode='mirror') then
    zoom(order=1, mode='mirror', grid_mode=True) (map_coordinates at the same half-pixel-centre coordinates in the 0.14-0.18
    releases), then np.clip to the input's [min, max] -- and this restatement is bit-identical to those scipy calls
    (tests/test_oracle_golden.py, tests/golden/segm.npz)."""
    image = np.asarray(image, np.float32)
    factors = np.divide(image.shape, out_shape)
    filtered = gaussian_filter_mirror_f32(image, np.maximum(0, (factors - 1) / 2))
    out = zoom_linear_mirror_f32(filtered, tuple(int(v) for v in out_shape))
    return np.clip(out, image.min(), image.max())


def expand_boxes(boxes, scale):
    """lib/utils/boxes_3d.py:271-292."""
    boxes = np.asarray(boxes)
    out = np.zeros(boxes.shape)
    for lo, hi in ((0, 3), (1, 4), (2, 5)):
        half = (boxes[:, hi] - boxes[:, lo]) * .5
        ctr = (boxes[:, hi] + boxes[:, lo]) * .5
        half = half * scale
        out[:, lo] = ctr - half
        out[:, hi] = ctr + half
    return out


def segm_results(cls_boxes, masks, ref_boxes, im_s, im_h, im_w, num_classes, cls_specific_mask=True, thresh_binarize=0.5,
                 resize=None):
    """lib/core/test.py:886-945.  `resize` defaults to the restatement above (tests also pass the scipy-backed one).
    Boxes that miss the volume give an all-zero volume (the reference's negative slice bounds misbehave there)."""
    resize = resize or resize_reflect_antialias
    masks = np.asarray(masks, np.float32)
    M = masks.shape[2]
    cls_segms = [[] for _ in range(num_classes)]
    mask_ind = 0
    scale = (M + 2.0) / M
    ref_boxes = expand_boxes(ref_boxes, scale).astype(np.int32)
    padded = np.zeros((M + 2, M + 2, M + 2), np.float32)
    for j in range(1, num_classes):
        segms = []
        for _ in range(len(cls_boxes[j])):
            padded[1:-1, 1:-1, 1:-1] = masks[mask_ind, j if cls_specific_mask else 0]
            b = ref_boxes[mask_ind]
            w, h, s = max(b[3] - b[0] + 1, 1), max(b[4] - b[1] + 1, 1), max(b[5] - b[2] + 1, 1)
            mask = (resize(padded, (s, h, w)) > np.float32(thresh_binarize)).astype(np.uint8)
            im_mask = np.zeros((im_s, im_h, im_w), np.uint8)
            x0, x1 = max(b[0], 0), min(b[3] + 1, im_w)
            y0, y1 = max(b[1], 0), min(b[4] + 1, im_h)
            z0, z1 = max(b[2], 0), min(b[5] + 1, im_s)
            if x1 > x0 and y1 > y0 and z1 > z0:
                im_mask[z0:z1, y0:y1, x0:x1] = mask[z0 - b[2]:z1 - b[2], y0 - b[1]:y1 - b[1], x0 - b[0]:x1 - b[0]]
            segms.append(im_mask)
            mask_ind += 1
        cls_segms[j] = segms
    assert mask_ind == masks.shape[0]
    return cls_segms


def scipy_resize_reflect_antialias(image, out_shape):
    """The scipy.ndimage calls skimage.transform.resize makes for (mode='reflect', anti_aliasing=True, order=1): the pin of
    resize_reflect_antialias.  scipy is the reference's own dependency (unversioned; 1.18.1 in this image)."""
    from scipy import ndimage as ndi
    image = np.asarray(image, np.float32)
    factors = np.divide(image.shape, out_shape)
    filtered = ndi.gaussian_filter(image, np.maximum(0, (factors - 1) / 2), cval=0, mode="mirror")
    out = ndi.zoom(filtered, [1 / f for f in factors], order=1, mode="mirror", cval=0, grid_mode=True)
    assert out.shape == tuple(out_shape)
    return np.clip(out, image.min(), image.max())


# ------------------------------------------------------------------------------- nuclei per-instance chain
def nuclei_normalise(box_img, box_prm):
    """tools/binarization_nuclei.py:111-121 (numpy 1.x scalar semantics for gray_range: no wrap-around).
    Where the script divides 0 by 0 (gray_max == 0 under the stretch, or a constant PRM crop) numpy yields NaN, warns, and
    the following `.astype(np.uint16)` turns NaN into 0 on x86-64 / aarch64; the script carries on with those zeros, and so
    does this restatement (round 1 reported status 7 there instead)."""
    g_lo, g_hi = box_img.min(), box_img.max()
    if int(g_hi) - int(g_lo) + 1 < 400:
        if g_hi == 0:
            box_img = np.zeros(box_img.shape, np.uint16) + g_lo          # uint16(NaN) == 0, + gray_min (== 0 here)
        else:
            box_img = (box_img.astype(float) / g_hi * 400).astype(np.uint16) + g_lo
    g_lo, g_hi = box_img.min(), box_img.max()
    p = box_prm.astype(float)
    if p.max() == p.min():
        return box_img.astype(np.uint16), np.zeros(box_prm.shape, np.uint16)     # round(NaN * span + gray_min) -> uint16 -> 0
    span = (int(g_hi) - int(g_lo)) & 0xFFFF
    box_prm = np.round((p - p.min()) / (p.max() - p.min()) * span + g_lo).astype(np.uint16)
    return box_img.astype(np.uint16), box_prm


def _largest_first(mask):
    """largest 26-connected component, ties -> the first label (np.argmax over ascending labels, :126-130);
    scipy.ndimage.label numbers components in raster order of their first voxel."""
    from scipy import ndimage as ndi
    lab, k = ndi.label(mask, structure=np.ones((3, 3, 3), bool))
    if k == 0:
        raise ValueError("no component")                     # np.argmax([]) in the reference
    return lab == (np.argmax(np.bincount(lab.ravel())[1:]) + 1)


def binarize_nuclei(volume, boxes, prm_crops):
    """tools/binarization_nuclei.py:92-149 for already selected instances (boxes int [n,6] in volume coordinates, clamped).
    cc3d.connected_components / skimage.morphology.binary_closing are unversioned, un-vendored dependencies of the reference
    and are not installed here: PARITY UNPINNED against them; restated with scipy.ndimage (label with the 3x3x3 structure;
    binary_closing(selem=None) = binary_dilation then binary_erosion(border_value=True) with the 6-neighbour cross, which
    is how skimage implements it).  Returns (seg uint16, status list, survive list, masks list)."""
    from scipy import ndimage as ndi
    seg = np.zeros(volume.shape, np.uint16)
    cross = ndi.generate_binary_structure(3, 1)
    status, survive, masks = [], [], []
    for i, b in enumerate(np.asarray(boxes)):
        mask_id = i + 1
        x1, y1, z1, x2, y2, z2 = (int(v) for v in b)
        box_img = volume[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1]
        box_prm = np.asarray(prm_crops[i]).reshape(box_img.shape)
        st, m = 0, np.zeros(box_img.shape, bool)
        try:
            i16, p16 = nuclei_normalise(box_img, box_prm)
            try:
                bi, _, _ = otsu_py_2d_fast(i16, p16)
            except UnboundLocalError:
                raise RuntimeError(1)
            try:
                cc = _largest_first(bi > 0)
                outside = _largest_first(~cc)
            except ValueError:
                raise RuntimeError(5)
            m = ndi.binary_erosion(ndi.binary_dilation(~outside, structure=cross), structure=cross, border_value=True)
        except RuntimeError as e:
            st = int(e.args[0])
        region = seg[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1]
        free = region == 0
        region[free] = (m.astype(np.uint16) * mask_id)[free]
        status.append(st)
        masks.append(m)
        survive.append(bool((seg == mask_id).any()))
    return seg, status, survive, masks


# ------------------------------------------------------------------------------- evaluation records
def eval_volume_soma(pred_mask, gt_mask, pred_score, iou_thresh=0.3):
    """tools/evaluation/eval_instance_segmentation_soma.py:166-216 for one image -> (score list, match list, n_pos)."""
    pred_score = np.asarray(pred_score, dtype=np.float64).reshape(-1, 2)
    pred_score = pred_score[pred_score[:, 1].argsort()[::-1], :]
    score = list(pred_score[:, 1])
    pred_ids = [v for v in np.unique(pred_mask).tolist() if v != 0]
    gt_ids = [v for v in np.unique(gt_mask).tolist() if v != 0]
    if len(pred_score) == 0:
        return score, [], len(gt_ids)
    if len(gt_ids) == 0:
        return score, [0] * len(pred_ids), 0
    pm = np.stack([pred_mask == i for i in pred_score[:, 0]])
    gm = np.stack([gt_mask == i for i in gt_ids])
    iou = mask_overlaps(pm, gm)[0]
    gt_index = iou.argmax(axis=1)
    gt_index[iou.max(axis=1) < iou_thresh] = -1
    selec = np.zeros(len(gt_ids), dtype=bool)
    match = []
    for gi in gt_index:
        if gi >= 0:
            match.append(0 if selec[gi] else 1)
            selec[gi] = True
        else:
            match.append(0)
    return score, match, len(gt_ids)


def eval_volume_nuclei(pred_mask, gt_mask, dets_bbox, gt_bbox, ovthresh=0.4):
    """tools/evaluation/evaluation_nuclei_f1score_seg.py:70-131 for one image -> (tp, fp, tp_pixel, gt_pixel, pre_pixel)."""
    gt_bbox = np.asarray(gt_bbox, dtype=float).reshape(-1, 6)
    dets_bbox = np.asarray(dets_bbox, dtype=float).reshape(-1, 6)
    detected = np.zeros(gt_bbox.shape[0], dtype=bool)
    tp, fp = np.zeros(dets_bbox.shape[0]), np.zeros(dets_bbox.shape[0])
    gt_b, pred_b = gt_mask > 0, pred_mask > 0
    keep = np.zeros(pred_mask.shape, dtype=bool)
    tp_pixel = 0
    if gt_bbox.shape[0] > 0:
        for ib, bbox in enumerate(dets_bbox):
            iw = np.maximum(np.minimum(gt_bbox[:, 3], bbox[3]) - np.maximum(gt_bbox[:, 0], bbox[0]) + 1., 0.)
            ih = np.maximum(np.minimum(gt_bbox[:, 4], bbox[4]) - np.maximum(gt_bbox[:, 1], bbox[1]) + 1., 0.)
            iss = np.maximum(np.minimum(gt_bbox[:, 5], bbox[5]) - np.maximum(gt_bbox[:, 2], bbox[2]) + 1., 0.)
            inters = iw * ih * iss
            uni = ((bbox[3] - bbox[0] + 1.) * (bbox[4] - bbox[1] + 1.) * (bbox[5] - bbox[2] + 1.) +
                   (gt_bbox[:, 3] - gt_bbox[:, 0] + 1.) * (gt_bbox[:, 4] - gt_bbox[:, 1] + 1.) * (gt_bbox[:, 5] - gt_bbox[:, 2] + 1.) - inters)
            overlaps = inters / uni
            jmax = np.argmax(overlaps)
            if overlaps[jmax] > ovthresh and not detected[jmax]:
                tp[ib] = 1.
                detected[jmax] = True
                x1, y1, z1, x2, y2, z2 = (max(int(v), 0) for v in bbox.astype(int))
                keep[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1] = pred_b[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1]
            else:
                fp[ib] = 1.
        tp_pixel = int(np.sum(keep & gt_b))
    return tp, fp, tp_pixel, int(gt_b.sum()), int(pred_b.sum())


# ------------------------------------------------------------------------------- reference builds
def ref_module(name):
    """Import a reference Cython module built into oracle/_ref (None when it is not there)."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        return importlib.import_module(name)
    except ImportError:
        return None


def ref_roialign_lib():
    p = os.path.join(REF_DIR, "libref_roialign3d.so")
    return ctypes.CDLL(p) if os.path.exists(p) else None
