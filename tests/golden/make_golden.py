"""Generate the golden vectors under tests/golden/ by running THE REFERENCE'S OWN CODE.

Run in the authoring container only (needs /root/reference and oracle/_ref built by
oracle/build_ref.py):   python tests/golden/make_golden.py

  nms_iou.npz   cython_nms_3d.nms_3d / nms_3d_volume / cython_bbox_3d.bbox_overlaps_3d
                (reference .pyx built unmodified except the 2-token numpy-2 patch)
  otsu.npz      tools/otsu.py:otsu_py_2d_fast imported with matplotlib/skimage stubbed and
                np.float = float (the module imports but never uses them)
  peaks.npz     lib/prm/peak_stimulation_3d.py:peak_stimulation_3d on CPU torch, with the median
                filter of lib/prm/peak_response_mapping_3d.py:45-49
  rle.npz       lib/utils/mask_3d.py literal of :75-79 (the only known-answer in the reference)
  segm.npz      lib/core/test.py:segm_results (its own source, skimage's resize replaced by the scipy.ndimage calls it makes)
  nuclei.npz    tools/binarization_nuclei.py lines 110-139 cut out of the script and executed per crop (cc3d / skimage closing
                bound to scipy.ndimage)
  eval.npz      tools/evaluation/eval_instance_segmentation_soma.py (calc_instance_segmentation_voc_prec_rec, voc_ap, unmodified,
                file IO bound to in-memory volumes) and evaluation_nuclei_f1score_seg.py (per-image body, lines 82-133)
  nuclei_script.npz  tools/binarization_nuclei.py lines 73-148 (selection + the whole per-instance loop) executed on one small volume
  soma_script.npz    tools/binarization_soma.py lines 57-104 (NMS, visit order, the whole per-instance loop, score table) executed on one small volume
  mask_iou.npz  tools/evaluation/mask_iou.py (mask_iou, mask_iou_fast, mask_ios_fast, mask_iog_fast) run as plain
                Python with numba stubbed, on stacks cut out of two small label volumes
The fixtures are small (about 2 MB in total) and are what `-m "not gpu"` tests pin the oracle against and
what the `-m gpu` tests pin the CUDA path against on the GPU box (no /root/reference there).
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle  # noqa: E402
from b200seg import synth  # noqa: E402


def load_ref_otsu():
    for m in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io", "skimage.exposure"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["skimage"].exposure = sys.modules["skimage.exposure"]
    np.float = float
    spec = importlib.util.spec_from_file_location("ref_otsu", os.path.join(REF, "tools", "otsu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_peaks():
    spec = importlib.util.spec_from_file_location("ref_peak", os.path.join(REF, "lib", "prm", "peak_stimulation_3d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_nms_iou():
    nms = oracle.ref_module("cython_nms_3d")
    bb = oracle.ref_module("cython_bbox_3d")
    assert nms is not None and bb is not None, "run oracle/build_ref.py first"
    rng = np.random.default_rng(1001)
    out = {}
    cases = [(0, 1, 100, 0.23, False), (1, 2, 100, 0.5, False), (2, 50, 120, 0.23, False), (3, 50, 60, 0.23, True),
             (4, 200, 150, 0.15, False), (5, 200, 150, 0.15, True), (6, 1000, 300, 0.23, False),
             (7, 1000, 250, 0.7, True), (8, 333, 80, 0.3, False)]
    for cid, n, ext, thr, integer in cases:
        d = synth.random_dets(rng, n, extent=(ext, ext, ext // 2 + 8), integer=integer)
        out["nms%d_dets" % cid] = d
        out["nms%d_thr" % cid] = np.float32(thr)
        out["nms%d_keep" % cid] = nms.nms_3d(d, np.float32(thr)).astype(np.int64)
        # volume-ordered variant only where volumes are distinct (ties are platform dependent)
        vol = (d[:, 3] - d[:, 0] + 1) * (d[:, 4] - d[:, 1] + 1) * (d[:, 5] - d[:, 2] + 1)
        if len(np.unique(vol)) == n:
            out["nms%d_keepvol" % cid] = nms.nms_3d_volume(d, np.float32(thr)).astype(np.int64)
    for cid, (N, K, integer) in enumerate([(1, 1, False), (40, 7, False), (300, 50, True), (257, 33, False)]):
        b = synth.random_dets(rng, N, extent=(120, 120, 60), integer=integer)[:, :6].copy()
        q = synth.random_dets(rng, K, extent=(120, 120, 60), integer=integer)[:, :6].copy()
        q[: min(N, K) // 2] = b[: min(N, K) // 2]            # identical boxes: IoU 0.9999999, not 1.0
        out["iou%d_boxes" % cid] = b
        out["iou%d_query" % cid] = q
        out["iou%d_out" % cid] = bb.bbox_overlaps_3d(b, q)
    np.savez_compressed(os.path.join(HERE, "nms_iou.npz"), **out)
    print("nms_iou.npz", len(out), "arrays")


def make_otsu():
    ro = load_ref_otsu()
    rng = np.random.default_rng(1003)
    out = {}
    n = 0
    for cid in range(14):
        shape = tuple(int(v) for v in rng.integers(6, 30, 3))
        vol, boxes, blobs = synth.blob_volume(rng, shape, 1, sigma_xy=(shape[1] / 6, shape[1] / 3),
                                              sigma_z=(shape[0] / 6, shape[0] / 3))
        prm = synth.prm_crop(blobs[0], (0, 0, 0, shape[2] - 1, shape[1] - 1, shape[0] - 1))
        if prm.max() == 0:
            continue
        if cid % 2 == 0:
            img16, prm16 = oracle.soma_normalise(vol, prm)       # soma call-site normalisation (uint16, ~301 levels)
        else:
            img16, prm16 = vol.astype(np.uint16), prm.astype(np.uint16)   # raw uint8 levels
        if cid == 5:
            prm16[:] = 77                                          # constant PRM axis (edge expansion +-0.5)
        mask, k, b = ro.otsu_py_2d_fast(img16, prm16)
        g = int(img16.max()) - int(img16.min()) + 1
        hist = np.histogram2d(img16.ravel(), prm16.ravel(), bins=g)[0].T
        out["otsu%d_img" % n] = img16
        out["otsu%d_prm" % n] = prm16
        out["otsu%d_b" % n] = np.int64(b)
        out["otsu%d_k" % n] = np.int64(k)
        out["otsu%d_mask" % n] = np.packbits(mask.ravel() > 0)
        out["otsu%d_hist" % n] = hist.astype(np.uint32)
        n += 1
    out["count"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "otsu.npz"), **out)
    print("otsu.npz", n, "crops")


def make_peaks():
    import torch
    rp = load_ref_peaks()

    def median_filter(input):                     # lib/prm/peak_response_mapping_3d.py:45-49
        b, c, s, h, w = input.size()
        thr, _ = torch.median(input.view(b, c, s * h * w), dim=2)
        return thr.contiguous().view(b, c, 1, 1, 1)

    rng = np.random.default_rng(1005)
    out = {}
    cases = [((1, 2, 8, 16, 16), 3, "smooth"), ((2, 3, 5, 9, 11), 3, "ties"), ((1, 1, 6, 10, 33), 5, "noise"),
             ((1, 2, 4, 7, 40), 3, "plateau"), ((1, 1, 9, 12, 12), 7, "noise"), ((1, 14, 8, 20, 20), 3, "smooth")]
    for cid, (shape, win, kind) in enumerate(cases):
        if kind == "smooth":
            x = np.concatenate([synth.response_map(rng, shape[2:], n_peaks=6, channels=shape[1]) for _ in range(shape[0])], 0)
        elif kind == "ties":
            x = (np.round(rng.normal(size=shape) * 2) / 2).astype(np.float32)
        elif kind == "plateau":
            x = np.zeros(shape, np.float32); x[..., 1:3, 2:4, 5:9] = 1.0; x[..., 0, 0, 0] = 1.0
        else:
            x = rng.normal(size=shape).astype(np.float32)
        out["pk%d_in" % cid] = x
        out["pk%d_win" % cid] = np.int64(win)
        t = torch.from_numpy(x)
        pl, agg = rp.peak_stimulation_3d(t, win_size=win, peak_filter=median_filter)
        out["pk%d_peaks_med" % cid] = pl.numpy().astype(np.int64)
        out["pk%d_agg_med" % cid] = agg.numpy()
        pl0, agg0 = rp.peak_stimulation_3d(t, win_size=win, peak_filter=None)
        out["pk%d_peaks_none" % cid] = pl0.numpy().astype(np.int64)
        out["pk%d_agg_none" % cid] = agg0.numpy()
    out["count"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "peaks.npz"), **out)
    print("peaks.npz", len(cases), "cases")


def make_mask_iou():
    """tools/evaluation/mask_iou.py run as plain Python (numba stubbed: nb.jit becomes the identity)."""
    nb = types.ModuleType("numba")
    nb.jit = lambda *a, **k: (lambda f: f)
    sys.modules["numba"] = nb
    spec = importlib.util.spec_from_file_location("ref_mask_iou", os.path.join(REF, "tools", "evaluation", "mask_iou.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(77)
    shape = (5, 7, 9)
    pred = np.zeros(shape, np.uint16); gt = np.zeros(shape, np.uint16)
    pred[0:3, 0:4, 0:5] = 1; pred[2:5, 4:7, 3:9] = 2; pred[0:2, 5:7, 0:2] = 3
    gt[0:3, 1:5, 1:6] = 1; gt[3:5, 3:7, 2:8] = 2; gt[0:1, 0:1, 8:9] = 3; gt[4:5, 0:2, 0:2] = 4
    pred[rng.random(shape) < 0.05] = 0
    pa = np.stack([pred == i for i in (2, 1, 3)])            # the caller's score order, not id order
    ga = np.stack([gt == i for i in (1, 2, 3, 4)])
    out = dict(pred=pred, gt=gt, pred_ids=np.array([2, 1, 3]), gt_ids=np.array([1, 2, 3, 4]),
               iou=mod.mask_iou_fast(pa, ga), ios=mod.mask_ios_fast(pa, ga), iog=mod.mask_iog_fast(pa, ga),
               iou_slow=mod.mask_iou(pa, ga))
    np.savez_compressed(os.path.join(HERE, "mask_iou.npz"), **out)
    print("mask_iou.npz", out["iou"])


def make_rle():
    sys.path.insert(0, os.path.join(REF, "lib", "utils"))
    spec = importlib.util.spec_from_file_location("ref_mask_3d", os.path.join(REF, "lib", "utils", "mask_3d.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    a = np.array([[[1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0]], [[1, 1, 1, 1, 1, 0], [1, 1, 1, 0, 0, 0]]])
    r = m.binary_mask_to_rle(a)
    np.savez_compressed(os.path.join(HERE, "rle.npz"), mask=a.astype(np.uint8),
                        counts=np.asarray(r["counts"], dtype=np.int64), size=np.asarray(r["size"], dtype=np.int64))
    print("rle.npz", r)


def make_prefilter():
    """The reference's prefilter / normalisation lines run verbatim on seeded volumes:
    binarization_nuclei.py:43-44 (scipy.ndimage, the reference's own dependency; version here recorded in the file),
    infer_simple.py:180-183 (z-score) and :233-238 (PRM -> uint8)."""
    import scipy
    from scipy import ndimage
    rng = np.random.default_rng(4344)
    out = {"scipy_version": np.array(scipy.__version__)}
    for name, dtype, hi, shape in (("u16", np.uint16, 65535, (11, 37, 70)), ("u8", np.uint8, 255, (9, 21, 33)), ("thin", np.uint16, 4000, (3, 5, 130))):
        img = rng.integers(0, hi + 1, shape).astype(dtype)
        img[rng.random(shape) < 0.3] = 0
        g = ndimage.gaussian_filter(img, sigma=1)                 # binarization_nuclei.py:43
        m = ndimage.median_filter(g, size=3)                      # binarization_nuclei.py:44
        out[name + "_img"], out[name + "_gauss"], out[name + "_median"] = img, g, m
        if name != "thin":
            out[name + "_gauss_s2"] = ndimage.gaussian_filter(img, sigma=2)
    im = out["u16_img"]
    mask = im > 0                                                 # infer_simple.py:180-183
    mean_val = np.mean(im[mask])
    std_val = np.std(im[mask])
    out["zs_out"] = (im - mean_val) / std_val
    out["zs_stats"] = np.array([mean_val, std_val])
    prm = (rng.random((3, 6, 10, 12)) ** 3).astype(np.float32) * np.array([1.0, 37.5, 1e-3], np.float32)[:, None, None, None]
    prm[1] -= 11.0
    out["prm_in"] = prm.copy()
    u8 = np.empty(prm.shape, np.uint8)
    for ch in range(prm.shape[0]):                                # infer_simple.py:233-238
        fm_ch = prm[ch, :]
        fm_ch -= np.min(fm_ch)
        fm_ch /= np.max(fm_ch)
        fm_ch *= 255.
        u8[ch] = fm_ch.astype(np.uint8)
    out["prm_u8"] = u8
    np.savez_compressed(os.path.join(HERE, "prefilter.npz"), **out)
    print("prefilter.npz", {k: v.shape for k, v in out.items()})


def make_proposals():
    """lib/modeling/generate_proposals_3d.py run unmodified (core.config imported with `nn` stubbed, the Cython NMS from
    oracle/_ref).  cfg.BBOX_XFORM_CLIP is handed over as a Python float so that numpy >= 2 keeps the float32 arithmetic
    the reference had under numpy 1.x (a float64 numpy scalar would now promote dw/dh/ds to float64)."""
    import oracle
    sys.path.insert(0, os.path.join(REF, "lib"))
    for n in ("cython_nms_3d", "cython_bbox_3d"):
        sys.modules["utils." + n] = oracle.ref_module(n)
    sys.modules["nn"] = types.ModuleType("nn")
    np.float = float
    import core.config as cc
    cc.cfg.BBOX_XFORM_CLIP = float(cc.cfg.BBOX_XFORM_CLIP)
    spec = importlib.util.spec_from_file_location("ref_gp", os.path.join(REF, "lib", "modeling", "generate_proposals_3d.py"))
    gp = importlib.util.module_from_spec(spec); spec.loader.exec_module(gp)
    spec = importlib.util.spec_from_file_location("ref_ga", os.path.join(REF, "lib", "modeling", "generate_anchors.py"))
    ga = importlib.util.module_from_spec(spec); spec.loader.exec_module(ga)
    import torch
    out = {}
    rng = np.random.default_rng(606)
    cases = [("a", 4, (12, 20, 30), np.array([[1., 1.], [1., 0.5]]), (6, 10, 12), 300, 100, 0.5),
             ("b", 8, (16, 24), np.array([[1., 1.]]), (4, 7, 9), 0, 50, 0.3),
             ("c", 4, (12, 20, 30), np.array([[1., 1.], [1., 0.5]]), (5, 8, 8), 200, 0, 0.7)]
    for name, stride, sizes, ratios, (S, H, W), pre, post, thr in cases:
        anchors = ga.generate_anchors_3d(stride=stride, sizes=sizes, aspect_ratios=ratios)
        A = anchors.shape[0]
        cc.cfg.TEST.RPN_PRE_NMS_TOP_N, cc.cfg.TEST.RPN_POST_NMS_TOP_N = pre, post
        cc.cfg.TEST.RPN_NMS_THRESH, cc.cfg.TEST.RPN_MIN_SIZE = thr, 0
        op = gp.GenerateProposalsOp_3d(anchors, 1.0 / stride).eval()
        scores = rng.permutation(2 * A * S * H * W).astype(np.float32).reshape(2, A, S, H, W) / np.float32(2 * A * S * H * W)   # distinct
        deltas = (rng.standard_normal((2, 6 * A, S, H, W)) * 0.4).astype(np.float32)
        deltas[0, 3::6][:, 0, 0, :3] = 9.0                                # beyond BBOX_XFORM_CLIP
        im_info = np.array([[S * stride, H * stride, W * stride, 1.0], [S * stride - 3, H * stride - 5, W * stride, 1.0]], np.float32)
        rois, probs, keep_idx = op(torch.from_numpy(scores), torch.from_numpy(deltas), torch.from_numpy(im_info))
        out.update({name + "_anchors": anchors, name + "_scores": scores, name + "_deltas": deltas, name + "_im_info": im_info,
                    name + "_cfg": np.array([stride, pre, post, thr], np.float64), name + "_rois": rois.astype(np.float32),
                    name + "_probs": probs.astype(np.float32), name + "_keep_idx_last": np.asarray(keep_idx, np.int64)})
        print(name, "A", A, "rois", rois.shape, "last image keep", len(keep_idx))
    np.savez_compressed(os.path.join(HERE, "proposals.npz"), **out)


def make_box_results():
    """box_results_with_nms_and_limit (lib/core/test.py:806-878): core/test.py itself needs cv2 / skimage / pycocotools at
    import, so only that function's source is compiled, against the reference's own cfg and utils.boxes_3d."""
    import ast
    import oracle
    sys.path.insert(0, os.path.join(REF, "lib"))
    for n in ("cython_nms_3d", "cython_bbox_3d"):
        sys.modules["utils." + n] = oracle.ref_module(n)
    sys.modules["nn"] = types.ModuleType("nn")
    np.float = float
    import core.config as cc
    import utils.boxes_3d as box_utils_3d
    src = open(os.path.join(REF, "lib", "core", "test.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "box_results_with_nms_and_limit")
    ns = {"np": np, "cfg": cc.cfg, "box_utils_3d": box_utils_3d}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_core_test", "exec"), ns)
    rng = np.random.default_rng(808)
    out = {}
    for name, R, ncls, per_im, use_idx in (("a", 300, 2, 100, True), ("b", 120, 4, 0, False), ("c", 200, 3, 1000, True)):
        cc.cfg.MODEL.NUM_CLASSES, cc.cfg.TEST.DETECTIONS_PER_IM = ncls, per_im
        cc.cfg.TEST.SCORE_THRESH, cc.cfg.TEST.NMS = 0.05, 0.3
        ctr = rng.uniform(10, 120, (R, 1, 3)) + rng.normal(0, 1.5, (R, ncls, 3))
        half = rng.uniform(4, 14, (R, ncls, 3))
        boxes = np.concatenate([ctr - half, ctr + half], axis=2).reshape(R, ncls * 6).astype(np.float32)
        scores = rng.permutation(R * ncls).reshape(R, ncls).astype(np.float32) / np.float32(R * ncls)
        idx = rng.permutation(50000)[:R].reshape(R, 1) if use_idx else None          # 2-D, the only shape the reference's limit step accepts
        if per_im == 100:
            cc.cfg.TEST.NMS = 0.9                                                     # keep > 100 so that the limit triggers
        s, b, cls_boxes, cls_idx = ns["box_results_with_nms_and_limit"](scores, boxes, idx)
        out.update({name + "_scores": scores, name + "_boxes": boxes, name + "_cfg": np.array([ncls, per_im, cc.cfg.TEST.NMS]),
                    name + "_out_scores": s, name + "_out_boxes": b})
        if idx is not None:
            out[name + "_idx"] = idx
        for j in range(1, ncls):
            out["%s_cls%d" % (name, j)] = cls_boxes[j]
            if idx is not None:
                out["%s_idx%d" % (name, j)] = cls_idx[j]
        print(name, "dets", s.shape)
    np.savez_compressed(os.path.join(HERE, "box_results.npz"), **out)


def make_segm():
    """segm_results (lib/core/test.py:886-945): the function's own source compiled against the reference's cfg and
    utils.boxes_3d (like make_box_results), with `transform.resize` bound to the scipy.ndimage calls skimage's resize
    delegates to (oracle.scipy_resize_reflect_antialias) -- scikit-image itself is not installed and unversioned in the
    reference, so the fixture pins everything except skimage's wrapper."""
    import ast
    sys.path.insert(0, os.path.join(REF, "lib"))
    for n in ("cython_nms_3d", "cython_bbox_3d"):
        sys.modules["utils." + n] = oracle.ref_module(n)
    sys.modules["nn"] = types.ModuleType("nn")
    np.float = float
    import core.config as cc
    import utils.boxes_3d as box_utils_3d
    src = open(os.path.join(REF, "lib", "core", "test.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "segm_results")
    transform = types.SimpleNamespace(resize=lambda img, shape, mode, anti_aliasing: (
        oracle.scipy_resize_reflect_antialias(img, shape) if (mode, anti_aliasing) == ("reflect", True) else 1 / 0))
    ns = {"np": np, "cfg": cc.cfg, "box_utils_3d": box_utils_3d, "transform": transform}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_core_test", "exec"), ns)
    rng = np.random.default_rng(919)
    out = {}
    S, H, W, M = 24, 40, 48, cc.cfg.MRCNN.RESOLUTION
    for name, ncls, cls_specific in (("a", 3, True), ("b", 2, False)):
        cc.cfg.MODEL.NUM_CLASSES, cc.cfg.MRCNN.CLS_SPECIFIC_MASK = ncls, cls_specific
        counts = [0, 5, 4][:ncls] if ncls == 3 else [0, 6]
        n = sum(counts)
        lo = np.stack([rng.uniform(-4, W - 10, n), rng.uniform(-4, H - 10, n), rng.uniform(-3, S - 6, n)], axis=1)
        ext = np.stack([rng.uniform(3, 30, n), rng.uniform(3, 26, n), rng.uniform(0.2, 16, n)], axis=1)
        ext[0] = (40.0, 9.5, 0.4)                          # one-voxel-thin in z after truncation, wide in x
        boxes = np.concatenate([lo, lo + ext], axis=1).astype(np.float32)
        g = np.indices((M, M, M)).astype(np.float64)
        masks_u8 = np.zeros((n, ncls, M, M, M), np.uint8)
        for i in range(n):
            for c in range(ncls):
                ctr = rng.uniform(4, M - 4, 3)
                rad = rng.uniform(2.5, 6.0, 3)
                d2 = sum(((g[k] - ctr[k]) / rad[k]) ** 2 for k in range(3))
                field = 1.0 / (1.0 + np.exp(3.0 * (d2 - 1.0))) + rng.normal(0, 0.08, (M, M, M))
                masks_u8[i, c] = np.clip(np.round(field * 256), 0, 255).astype(np.uint8)
        masks = masks_u8.astype(np.float32) / np.float32(256)
        cls_boxes = [[]] + [np.zeros((c, 7), np.float32) for c in counts[1:]]
        segms = ns["segm_results"](cls_boxes, masks, boxes, S, H, W)
        vols = np.stack([v for j in range(1, ncls) for v in segms[j]])
        out.update({name + "_masks_u8": masks_u8, name + "_boxes": boxes, name + "_counts": np.array(counts),
                    name + "_cfg": np.array([ncls, int(cls_specific), M, S, H, W]), name + "_vols_bits": np.packbits(vols.reshape(n, -1), axis=1)})
        print(name, "dets", n, "foreground voxels", int(vols.sum()))
    np.savez_compressed(os.path.join(HERE, "segm.npz"), **out)


def make_nuclei():
    """tools/binarization_nuclei.py:110-139 (normalise, otsu_py_2d_fast, largest component, hole filling, closing): the
    script body cannot be imported (hard-coded paths, tif IO), so exactly those lines are cut out of the file, dedented and
    executed per crop with the reference's own tools/otsu.py; cc3d.connected_components and skimage's binary_closing (not
    installed, unversioned) are bound to their scipy.ndimage equivalents."""
    import textwrap
    import zlib
    from scipy import ndimage as ndi
    ref_otsu = load_ref_otsu()
    lines = open(os.path.join(REF, "tools", "binarization_nuclei.py")).read().split("\n")
    body = textwrap.dedent("\n".join(lines[109:139]))           # file lines 110..139
    assert body.lstrip().startswith("# normalize gray image") and "binary_closing" in body.split("\n")[-1]
    code = compile(body, "ref_binarization_nuclei_110_139", "exec")
    morphology = types.SimpleNamespace(binary_closing=lambda m: ndi.binary_erosion(
        ndi.binary_dilation(m, structure=ndi.generate_binary_structure(3, 1)), structure=ndi.generate_binary_structure(3, 1), border_value=True))
    connected_components = lambda m: ndi.label(m, structure=np.ones((3, 3, 3), bool))[0]
    rng = np.random.default_rng(139)
    c = synth.postproc_case(139, shape=(32, 96, 128), n_blobs=8, n_dup=0, n_false=0)
    vol = c["volume"].astype(np.int32)
    for bl in c["blobs"][::2]:                                  # hollow every other blob: cavities for the hole filling
        cz, cy, cx = bl["c"]
        zz, yy, xx = np.ogrid[:32, :96, :128]
        r2 = ((zz - cz) / (0.45 * bl["sz"])) ** 2 + ((yy - cy) / (0.45 * bl["sxy"])) ** 2 + ((xx - cx) / (0.45 * bl["sxy"])) ** 2
        vol -= (0.9 * bl["amp"] * np.exp(-0.5 * r2)).astype(np.int32)
    vol = np.clip(vol, 0, 255).astype(np.uint8)
    out = {"count": 0}
    k = 0
    for dtype in (np.uint8, np.uint16):
        v = vol if dtype == np.uint8 else (vol.astype(np.uint16) * 7 + 11)
        for i in ((0, 1, 3, 5) if dtype == np.uint8 else (0, 6)):
            x1, y1, z1, x2, y2, z2 = c["boxes"][i]
            box_img = v[z1:z2 + 1, y1:y2 + 1, x1:x2 + 1].copy()
            box_prm = c["prm"][c["crop_off"][i]:c["crop_off"][i + 1]].reshape(box_img.shape).copy()
            ns = {"np": np, "box_img": box_img.copy(), "box_prm": box_prm.copy(), "otsu_py_2d_fast": ref_otsu.otsu_py_2d_fast,
                  "connected_components": connected_components, "morphology": morphology}
            exec(code, ns)
            out.update({"n%d_img" % k: box_img, "n%d_prm" % k: box_prm, "n%d_crc16" % k: np.array([zlib.crc32(np.ascontiguousarray(ns["box_img"].astype(np.uint16)).tobytes()),
                                                     zlib.crc32(np.ascontiguousarray(ns["box_prm"].astype(np.uint16)).tobytes())], np.int64),
                        "n%d_b" % k: np.int32(ns["b"]),
                        "n%d_mask" % k: np.packbits(ns["largestCC"].ravel())})
            print("nuclei crop", k, dtype.__name__, box_img.shape, "b", ns["b"], "fg", int(ns["largestCC"].sum()),
                  "otsu fg", int((ns["labels_out"] > -1).sum()))
            k += 1
    out["count"] = k
    np.savez_compressed(os.path.join(HERE, "nuclei.npz"), **out)


def _eval_images(rng):
    """two small images: ground-truth / predicted label volumes with shifted, missing and spurious instances + scores + boxes"""
    imgs = []
    for k in range(2):
        S, H, W = 12, 40, 48
        gt = np.zeros((S, H, W), np.uint16); pred = np.zeros((S, H, W), np.uint16)
        gt_boxes, det_boxes, score = [], [], []
        n = 7 + k
        for i in range(n):
            z, y, x = int(rng.integers(0, S - 6)), int(rng.integers(0, H - 12)), int(rng.integers(0, W - 12))
            d, h, w = int(rng.integers(4, 7)), int(rng.integers(7, 12)), int(rng.integers(7, 12))
            region = gt[z:z + d, y:y + h, x:x + w]
            region[region == 0] = i + 1
            gt_boxes.append([x, y, z, x + w - 1, y + h - 1, z + d - 1])
            if i % 4 != 3:                                            # every fourth instance is missed
                dz, dy, dx = (int(v) for v in rng.integers(-1, 2, 3))
                z2, y2, x2 = max(z + dz, 0), max(y + dy, 0), max(x + dx, 0)
                region = pred[z2:z2 + d, y2:y2 + h, x2:x2 + w]
                region[region == 0] = 10 + i
                det_boxes.append([x2, y2, z2, min(x2 + w, W) - 1, min(y2 + h, H) - 1, min(z2 + d, S) - 1])
                score.append([10 + i, float(rng.random())])
        pred[0:2, 0:3, W - 4:W][pred[0:2, 0:3, W - 4:W] == 0] = 99   # a spurious prediction
        det_boxes.append([W - 4, 0, 0, W - 1, 2, 1]); score.append([99, 0.05 + 0.1 * k])
        present = set(np.unique(pred).tolist())
        keep = [j for j, sc in enumerate(score) if int(sc[0]) in present]     # fully overwritten predictions carry no score row
        imgs.append(dict(gt=gt, pred=pred, gt_boxes=np.array(gt_boxes, np.float32), det_boxes=np.array(det_boxes, np.float32)[keep],
                         score=np.array(score, np.float64)[keep]))
    return imgs


def make_eval():
    """tools/evaluation/eval_instance_segmentation_soma.py: calc_instance_segmentation_voc_prec_rec + voc_ap run unmodified
    (skimage.io.imread / np.load bound to in-memory volumes, numba stubbed for mask_iou.py), and
    tools/evaluation/evaluation_nuclei_f1score_seg.py: the per-image body (file lines 82-133) cut out of the script and executed."""
    import tempfile
    import textwrap
    nb = types.ModuleType("numba"); nb.jit = lambda *a, **k: (lambda f: f); sys.modules["numba"] = nb
    store = {}
    sk, skio = types.ModuleType("skimage"), types.ModuleType("skimage.io")
    skio.imread = lambda path: store[path]
    sk.io = skio
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, skio
    sys.path.insert(0, os.path.join(REF, "tools", "evaluation"))
    spec = importlib.util.spec_from_file_location("ref_eval_soma", os.path.join(REF, "tools", "evaluation", "eval_instance_segmentation_soma.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(216)
    imgs = _eval_images(rng)
    out = {"count": len(imgs)}
    with tempfile.TemporaryDirectory() as tmp:
        names = []
        for k, im in enumerate(imgs):
            name = "img%d" % k
            names.append(name)
            store[os.path.join(tmp, name + ".tif")] = im["pred"]
            store[os.path.join("gt", name, name + ".tif")] = im["gt"]
            np.save(os.path.join(tmp, name + ".npy"), im["score"])
            for key, v in im.items():
                out["%s_%s" % (name, key)] = v
        prec, rec = mod.calc_instance_segmentation_voc_prec_rec(tmp, "gt", names, 0.3)
        _, _, ap = mod.voc_ap(rec, prec, use_07_metric=False)
    out.update(prec=prec, rec=rec, ap=np.float64(ap))
    print("soma eval: ap", ap, "rows", len(prec))
    lines = open(os.path.join(REF, "tools", "evaluation", "evaluation_nuclei_f1score_seg.py")).read().split("\n")
    body = textwrap.dedent("\n".join(lines[81:133]))                  # file lines 82..133
    assert body.startswith("tp = np.zeros") and body.rstrip().split("\n")[-1].lstrip().startswith("tp_pixel +=")
    code = compile(body, "ref_evaluation_nuclei_82_133", "exec")
    for k, im in enumerate(imgs):
        store.clear()
        store["GT"] = im["gt"]; store["PRED"] = im["pred"]
        io = types.SimpleNamespace(imread=lambda path: store["GT"] if "man_seg" in path else store["PRED"])
        ns = {"np": np, "os": os, "io": io, "src_path": "s", "res_path": "r", "track": "01", "img_name": "01_t000", "ovthresh": 0.4,
              "dets_bbox": im["det_boxes"].astype(float), "gt_bbox": im["gt_boxes"].astype(float),
              "detected": np.zeros(len(im["gt_boxes"]), dtype=bool), "gt_pixel": 0, "pre_pixel": 0, "tp_pixel": 0, "print": lambda *a, **k: None}
        exec(code, ns)
        out.update({"img%d_tp" % k: ns["tp"], "img%d_fp" % k: ns["fp"],
                    "img%d_pixels" % k: np.array([ns["tp_pixel"], ns["gt_pixel"], ns["pre_pixel"]], np.int64)})
        print("nuclei eval img", k, "tp", int(ns["tp"].sum()), "fp", int(ns["fp"].sum()), "pixels", ns["tp_pixel"], ns["gt_pixel"], ns["pre_pixel"])
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)


def make_nuclei_script():
    """tools/binarization_nuclei.py file lines 73-148 (edge filter, nms_3d_volume, score cut, the whole per-instance loop with
    tile-relative clamping, first-come paste and the survivor table) cut out of the script and executed on one small synthetic
    volume: PRM tifs are served from memory, otsu is the reference's tools/otsu.py, nms_3d_volume the reference's Cython
    (oracle/_ref); cc3d / skimage closing are bound to scipy.ndimage as in make_nuclei."""
    import textwrap
    from scipy import ndimage as ndi
    ref_otsu = load_ref_otsu()
    lines = open(os.path.join(REF, "tools", "binarization_nuclei.py")).read().split("\n")
    body = textwrap.dedent("\n".join(lines[72:148]))             # file lines 73..148
    assert body.startswith("# remove broken boxes at edges") and "id_det = np.concatenate" in body.rstrip().split("\n")[-1]
    code = compile(body, "ref_binarization_nuclei_73_148", "exec")
    cross = ndi.generate_binary_structure(3, 1)
    morphology = types.SimpleNamespace(binary_closing=lambda m: ndi.binary_erosion(ndi.binary_dilation(m, structure=cross), structure=cross, border_value=True))
    connected_components = lambda m: ndi.label(m, structure=np.ones((3, 3, 3), bool))[0]
    nms_mod = oracle.ref_module("cython_nms_3d")
    box_utils_3d = types.SimpleNamespace(nms_3d_volume=lambda d, t: nms_mod.nms_3d_volume(np.ascontiguousarray(d, dtype=np.float32), np.float32(t)))
    from b200seg.binarization import dets_to_boxes, crop_offsets
    rng = np.random.default_rng(148)
    S, H, W, norm_side = 16, 160, 256, 64
    vol = rng.uniform(0, 40, (S, H, W))
    blobs, bx = [], []
    zz, yy, xx = np.ogrid[:S, :H, :W]
    for gy in (40, 120):
        for gx in (32, 96, 160, 224):
            cz, cy, cx = 8 + rng.uniform(-1, 1), gy + rng.uniform(-6, 6), gx + rng.uniform(-6, 6)
            sz, sxy, amp = rng.uniform(2.5, 3.5), rng.uniform(8, 10), rng.uniform(120, 200)
            vol += amp * np.exp(-0.5 * (((zz - cz) / sz) ** 2 + ((yy - cy) / sxy) ** 2 + ((xx - cx) / sxy) ** 2))
            blobs.append(dict(c=(cz, cy, cx), sz=sz, sxy=sxy, amp=amp))
            bx.append([cx - 2 * sxy, cy - 2 * sxy, cz - 2 * sz, cx + 2 * sxy, cy + 2 * sxy, cz + 2 * sz])
    owners = list(range(8)) + [1, 5] + [0, 3]
    bx = np.array(bx)
    allb = np.concatenate([bx, bx[[1, 5]] + rng.uniform(-2, 2, (2, 6)), np.array([[100., 60., 2., 140., 100., 9.], [200., 20., 5., 236., 58., 12.]])])
    scores = np.array([0.9, 0.5, 0.8, 0.3, 0.7, 0.95, 0.6, 0.45, 0.55, 0.35, 0.25, 0.85])   # the false box with a constant PRM crop stays below the 0.4 cut (the script divides 0 by 0 there)
    dets = np.hstack([allb, scores[:, None]]).astype(np.float32)
    img = (np.clip(vol, 0, 255).astype(np.uint8) // 32 * 32).astype(np.uint8)       # coarse gray levels: a small fixture
    boxes0 = dets_to_boxes(dets, (S, H, W))
    off0 = crop_offsets(boxes0)
    c = dict(boxes=boxes0, crop_off=off0, prm=np.concatenate([synth.prm_crop(blobs[owners[i]], boxes0[i]).ravel() for i in range(len(dets))]))
    n = len(dets)
    # tile of every detection: the 32-aligned tile origin that keeps the box centre inside
    cx, cy = (dets[:, 0] + dets[:, 3]) / 2, (dets[:, 1] + dets[:, 4]) / 2
    tw = np.clip((cx // 32 - 1) * 32, 0, W - norm_side).astype(int)
    th = np.clip((cy // 32 - 1) * 32, 0, H - norm_side).astype(int)
    instance_idex = np.stack([np.arange(n), np.zeros(n, int), tw, th, np.zeros(n, int)], axis=1).astype(int)
    tiles = np.zeros((n, S, norm_side, norm_side), np.uint8)
    for i in range(n):                                            # the PRM "tif" of instance i: its response inside its tile
        ob, full = c["boxes"][i], None
        full = c["prm"][c["crop_off"][i]:c["crop_off"][i + 1]].reshape(ob[5] - ob[2] + 1, ob[4] - ob[1] + 1, ob[3] - ob[0] + 1)
        for z in range(ob[2], ob[5] + 1):
            for y in range(max(ob[1], th[i]), min(ob[4] + 1, th[i] + norm_side)):
                x0, x1 = max(ob[0], tw[i]), min(ob[3] + 1, tw[i] + norm_side)
                if x1 > x0:
                    tiles[i, z, y - th[i], x0 - tw[i]:x1 - tw[i]] = full[z - ob[2], y - ob[1], x0 - ob[0]:x1 - ob[0]]
    io = types.SimpleNamespace(imread=lambda path: tiles[int(path.split("/")[-2])].copy())
    ns = {"np": np, "os": os, "io": io, "dets": dets.copy(), "instance_idex": instance_idex.copy(), "width": W, "nms_thresh": 0.15,
          "img": img, "norm_side": norm_side, "slices": S, "prm_path": "p", "img_name": "x", "box_utils_3d": box_utils_3d,
          "otsu_py_2d_fast": ref_otsu.otsu_py_2d_fast, "connected_components": connected_components, "morphology": morphology}
    exec(code, ns)
    print("nuclei script: visited", len(ns["dets"]), "of", n, "kept rows", ns["id_det"].shape, "labels", len(np.unique(ns["seg"])) - 1)
    np.savez_compressed(os.path.join(HERE, "nuclei_script.npz"), img=img, dets=dets, instance_idex=instance_idex, tiles=tiles,
                        cfg=np.array([W, norm_side, S]), seg=ns["seg"], id_det=ns["id_det"], visited_dets=ns["dets"])


def make_soma_script():
    """tools/binarization_soma.py file lines 57-104 (nms_3d, visit in descending score, mask_id bookkeeping, the per-instance
    loop with tile crops, normalisation, otsu_py_2d_fast, largest component, first-come paste and the score table) cut out of
    the script and executed on one small synthetic volume: PRM tifs served from memory, nms_3d = the reference's Cython
    (oracle/_ref), otsu = the reference's tools/otsu.py; skimage.measure.label bound to scipy.ndimage.label (3x3x3)."""
    import textwrap
    from scipy import ndimage as ndi
    ref_otsu = load_ref_otsu()
    lines = open(os.path.join(REF, "tools", "binarization_soma.py")).read().split("\n")
    body = textwrap.dedent("\n".join(lines[56:104]))             # file lines 57..104
    assert body.startswith("keep = box_utils_3d.nms_3d(dets, nms_thresh)") and body.rstrip().split("\n")[-1].lstrip().startswith("scores = np.concatenate")
    code = compile(body, "ref_binarization_soma_57_104", "exec")
    nms_mod = oracle.ref_module("cython_nms_3d")
    box_utils_3d = types.SimpleNamespace(nms_3d=lambda d, t: nms_mod.nms_3d(np.ascontiguousarray(d, dtype=np.float32), np.float32(t)))
    label = lambda m: ndi.label(m, structure=np.ones((3, 3, 3), bool))[0]
    rng = np.random.default_rng(104)
    shape = (40, 120, 144)
    c = synth.postproc_case(104, shape=shape, n_blobs=9, n_dup=4, n_false=3)
    img = (c["volume"] // 32 * 32).astype(np.uint8)               # coarse gray levels: a small fixture
    dets, n = c["dets"], len(c["dets"])
    # tile origins like the script's (ws, hs, ss): any tile that contains the box will do
    b = c["boxes"]
    ws = np.minimum(b[:, 0] // 48 * 48, 48); hs = np.minimum(b[:, 1] // 40 * 40, 40); ss = np.minimum(b[:, 2] // 16 * 16, 16)
    instance_idex = np.stack([np.arange(n), np.zeros(n, int), ws, hs, ss], axis=1).astype(int)
    tiles = []
    for i in range(n):                                            # the PRM "tif" of instance i = its response inside its tile
        t = np.zeros((min(64, shape[0] - ss[i]), min(160, shape[1] - hs[i]), min(160, shape[2] - ws[i])), np.uint8)
        ob = b[i]
        full = c["prm"][c["crop_off"][i]:c["crop_off"][i + 1]].reshape(ob[5] - ob[2] + 1, ob[4] - ob[1] + 1, ob[3] - ob[0] + 1)
        t[ob[2] - ss[i]:ob[5] + 1 - ss[i], ob[1] - hs[i]:ob[4] + 1 - hs[i], ob[0] - ws[i]:ob[3] + 1 - ws[i]] = full
        tiles.append(t)
    io = types.SimpleNamespace(imread=lambda path: tiles[int(path.split("/")[-2])].copy())
    ns = {"np": np, "os": os, "io": io, "dets": dets.copy(), "instance_idex": instance_idex.copy(), "nms_thresh": 0.23, "img": img,
          "seg": np.zeros(img.shape, np.uint16), "mask_id": 0, "prm_path": "p", "im_name": "x", "box_utils_3d": box_utils_3d,
          "otsu_py_2d_fast": ref_otsu.otsu_py_2d_fast, "label": label}
    exec(code, ns)
    print("soma script: visited", len(ns["dets"]), "of", n, "score rows", ns["scores"].shape, "labels", len(np.unique(ns["seg"])) - 1)
    np.savez_compressed(os.path.join(HERE, "soma_script.npz"), img=img, dets=dets, boxes=c["boxes"], prm=c["prm"], crop_off=c["crop_off"],
                        seg=ns["seg"], scores=ns["scores"], visited_dets=ns["dets"])


if __name__ == "__main__":
    make_soma_script()
    make_nuclei_script()
    make_eval()
    make_nuclei()
    make_segm()
    make_box_results()
    make_proposals()
    make_prefilter()
    make_nms_iou()
    make_otsu()
    make_peaks()
    make_rle()
    make_mask_iou()
