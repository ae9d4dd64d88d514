"""CPU: the oracle (oracle/oracle.c) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py) and, when present, against the reference Cython built in oracle/_ref."""
import numpy as np
import pytest

import oracle
from b200seg import synth


def test_nms_golden(golden):
    g = golden("nms_iou.npz")
    ids = sorted({int(k[3:].split("_")[0]) for k in g.files if k.startswith("nms")})
    assert len(ids) >= 8
    for i in ids:
        d, thr = g["nms%d_dets" % i], float(g["nms%d_thr" % i])
        assert np.array_equal(oracle.nms_3d(d, thr), g["nms%d_keep" % i]), i
        if "nms%d_keepvol" % i in g.files:
            assert np.array_equal(oracle.nms_3d_volume(d, thr), g["nms%d_keepvol" % i]), i


def test_iou_golden(golden):
    g = golden("nms_iou.npz")
    for i in range(4):
        out = oracle.bbox_overlaps_3d(g["iou%d_boxes" % i], g["iou%d_query" % i])
        assert np.array_equal(out.view(np.uint32), g["iou%d_out" % i].view(np.uint32)), i
    # the reference's self-overlap is 0.9999999, not 1.0 (mixed precision); golden case 1 has identical boxes
    assert g["iou1_out"][0, 0] < 1.0


def test_otsu_golden(golden):
    g = golden("otsu.npz")
    n = int(g["count"])
    assert n >= 10
    for i in range(n):
        img, prm = g["otsu%d_img" % i], g["otsu%d_prm" % i]
        mask, k, b, hist = oracle.otsu_py_2d_fast(img, prm, want_hist=True)
        assert k == int(g["otsu%d_k" % i]) == -1
        assert b == int(g["otsu%d_b" % i]), i
        assert np.array_equal(np.packbits(mask.ravel() > 0), g["otsu%d_mask" % i]), i
        assert np.array_equal(hist.astype(np.uint32), g["otsu%d_hist" % i]), i
        assert set(np.unique(mask)) <= {0, 255}


def test_otsu_constant_crop_raises():
    img = np.full((4, 5, 6), 100, np.uint16)
    with pytest.raises(UnboundLocalError):       # the reference fails on k_max (otsu.py:277)
        oracle.otsu_py_2d_fast(img, img)


def test_peaks_golden(golden):
    g = golden("peaks.npz")
    for i in range(int(g["count"])):
        x, win = g["pk%d_in" % i], int(g["pk%d_win" % i])
        p, agg, _ = oracle.peak_stimulation_3d(x, win_size=win, filter_mode="median")
        assert np.array_equal(p, g["pk%d_peaks_med" % i]), i
        np.testing.assert_allclose(agg, g["pk%d_agg_med" % i], rtol=1e-5, atol=1e-6, equal_nan=True)
        p0, agg0, _ = oracle.peak_stimulation_3d(x, win_size=win, filter_mode=None)
        assert np.array_equal(p0, g["pk%d_peaks_none" % i]), i
        np.testing.assert_allclose(agg0, g["pk%d_agg_none" % i], rtol=1e-5, atol=1e-6, equal_nan=True)


def test_oracle_vs_reference_cython_random():
    """Extra pinning where the reference build is available (authoring container + GPU box)."""
    nms = oracle.ref_module("cython_nms_3d")
    bb = oracle.ref_module("cython_bbox_3d")
    if nms is None or bb is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    for t in range(60):
        n = int(rng.choice([1, 3, 50, 200, 600]))
        d = synth.random_dets(rng, n, extent=(150, 150, 60), integer=(t % 3 == 0))
        thr = np.float32(rng.choice([0.15, 0.23, 0.5]))
        assert np.array_equal(nms.nms_3d(d, thr), oracle.nms_3d(d, thr))
        q = synth.random_dets(rng, int(rng.integers(1, 40)), extent=(150, 150, 60), integer=(t % 2 == 0))[:, :6].copy()
        a, o = bb.bbox_overlaps_3d(d[:, :6].copy(), q), oracle.bbox_overlaps_3d(d[:, :6], q)
        assert np.array_equal(a.view(np.uint32), o.view(np.uint32))


def test_paste_first_come_wins():
    seg = np.zeros((4, 6, 8), np.uint16)
    boxes = np.array([[1, 1, 0, 4, 3, 2], [3, 2, 1, 6, 5, 3]], np.int32)
    m0 = np.ones((3, 3, 4), np.uint8)
    m1 = np.ones((3, 4, 4), np.uint8); m1[0, 0, 0] = 0
    surv = oracle.paste_labels(seg, boxes, np.array([1, 2], np.uint16), [m0, m1])
    # numpy restatement of binarization_soma.py:100-102
    ref = np.zeros_like(seg)
    for i, (b, m) in enumerate(zip(boxes, [m0, m1])):
        sub = ref[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
        z = sub == 0
        sub[z] = (m.astype(np.uint16) * (i + 1))[z]
    assert np.array_equal(seg, ref)
    assert surv.tolist() == [True, True]
    # an instance completely covered by an earlier one does not survive
    seg2 = np.zeros((2, 2, 2), np.uint16)
    bx = np.array([[0, 0, 0, 1, 1, 1], [0, 0, 0, 0, 0, 0]], np.int32)
    s2 = oracle.paste_labels(seg2, bx, np.array([1, 2], np.uint16), [np.ones((2, 2, 2), np.uint8), np.ones((1, 1, 1), np.uint8)])
    assert s2.tolist() == [True, False] and (seg2 == 1).all()


def test_roialign_oracle_properties():
    """No CPU reference exists for RoIAlign3D; check invariants of the restatement:
    constant features -> constant output inside the volume; layout quirk (H,W,S)."""
    feat = np.full((1, 2, 6, 7, 8), 3.5, np.float32)
    rois = np.array([[0, 4, 4, 4, 20, 20, 16]], np.float32)
    out = oracle.roialign3d_fwd(feat, rois, 7, 0.25, 2)
    np.testing.assert_allclose(out, 3.5, rtol=1e-6)
    # a feature map that varies only along z: output must vary along the LAST output axis (ps fastest)
    f2 = np.broadcast_to(np.arange(6, dtype=np.float32)[None, None, :, None, None], (1, 1, 6, 7, 8)).copy()
    o2 = oracle.roialign3d_fwd(f2, rois, 7, 0.25, 2)[0, 0]
    assert np.allclose(o2, o2[:1, :1, :]) and not np.allclose(o2[0, 0, 0], o2[0, 0, -1])
    # backward of ones sums to (#bins fully inside) per channel: sum of all tap weights / count = 1 per bin
    g = np.ones((1, 1, 7, 7, 7), np.float32)
    gi = oracle.roialign3d_bwd(g, rois, (1, 1, 6, 7, 8), 0.25, 2)
    np.testing.assert_allclose(gi.sum(), 343.0, rtol=1e-5)


def test_rle_golden_and_reference_cython(golden):
    """lib/utils/mask_3d.py:75-79 is the only known-answer vector of the reference; the Cython twin built by
    oracle/build_ref.py (when present) pins the restatement on random masks."""
    d = golden("rle.npz")
    r = oracle.binary_mask_to_rle(d["mask"])
    assert r["counts"] == d["counts"].tolist() and r["size"] == d["size"].tolist()
    assert np.array_equal(oracle.rle_to_binary_mask(r), d["mask"])
    ref = oracle.ref_module("cython_mask_3d")
    rng = np.random.default_rng(3)
    for sh in [(3, 4, 5), (1, 1, 1), (7, 2, 9), (5, 6, 1)]:
        for dens in (0.0, 0.2, 0.7, 1.0):
            m = (rng.random(sh) < dens).astype(np.uint8)
            a = oracle.binary_mask_to_rle(m)
            assert sum(a["counts"]) == m.size
            assert np.array_equal(oracle.rle_to_binary_mask(a), m)
            if ref is not None:
                b = ref.binary_mask_to_rle(m)
                assert a["counts"] == [int(x) for x in b["counts"]] and a["size"] == list(b["size"])


def test_mask_overlaps_golden(golden):
    """tools/evaluation/mask_iou.py:10-109 run as plain Python (numba stubbed) generated the fixture."""
    d = golden("mask_iou.npz")
    pa = np.stack([d["pred"] == i for i in d["pred_ids"]])
    ga = np.stack([d["gt"] == i for i in d["gt_ids"]])
    iou, ios, iog = oracle.mask_overlaps(pa, ga)
    for k, v in (("iou", iou), ("ios", ios), ("iog", iog), ("iou_slow", iou)):
        assert np.array_equal(v, d[k], equal_nan=True), k


def test_prefilter_golden(golden):
    """scipy.ndimage (the reference's dependency for binarization_nuclei.py:43-44) and the reference's inline numpy
    lines generated the fixture; the oracle restatement must reproduce every array bit for bit."""
    d = golden("prefilter.npz")
    for n in ("u16", "u8", "thin"):
        g = oracle.gaussian_filter(d[n + "_img"], 1)
        assert g.dtype == d[n + "_img"].dtype and np.array_equal(g, d[n + "_gauss"]), n
        assert np.array_equal(oracle.median_filter3(g), d[n + "_median"]), n
        if n != "thin":
            assert np.array_equal(oracle.gaussian_filter(d[n + "_img"], 2), d[n + "_gauss_s2"]), n
    z, mu, sd = oracle.zscore_norm(d["u16_img"])
    assert np.array_equal(z, d["zs_out"]) and mu == d["zs_stats"][0] and sd == d["zs_stats"][1]
    assert np.array_equal(oracle.prm_to_uint8(d["prm_in"]), d["prm_u8"])


def test_prefilter_oracle_vs_scipy_live():
    """Where scipy is importable the restatement is also checked live on fresh random volumes (odd sizes, dims < radius)."""
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(5)
    for shape, dtype, hi in [((6, 9, 11), np.uint16, 65535), ((2, 3, 40), np.uint8, 255), ((17, 8, 5), np.uint16, 700)]:
        img = rng.integers(0, hi + 1, shape).astype(dtype)
        for sigma in (0.5, 1, 1.5):
            assert np.array_equal(oracle.gaussian_filter(img, sigma), ndimage.gaussian_filter(img, sigma=sigma)), (shape, sigma)
        assert np.array_equal(oracle.median_filter3(img), ndimage.median_filter(img, size=3)), shape


def test_generate_proposals_golden(golden):
    """lib/modeling/generate_proposals_3d.py, run unmodified by tests/golden/make_golden.py, generated the fixture."""
    d = golden("proposals.npz")
    for name in "abc":
        stride, pre, post, thr = d[name + "_cfg"]
        rois, probs = [], []
        for i in range(2):
            p, s, k = oracle.generate_proposals(d[name + "_scores"][i], d[name + "_deltas"][i], d[name + "_im_info"][i], d[name + "_anchors"],
                                                stride, int(pre), int(post), float(thr))
            rois.append(np.hstack([np.full((len(p), 1), i, np.float32), p]))
            probs.append(s)
        assert np.array_equal(np.vstack(rois), d[name + "_rois"]), name
        assert np.array_equal(np.concatenate(probs)[:, None], d[name + "_probs"]), name
        assert np.array_equal(k, d[name + "_keep_idx_last"]), name


def _box_results_check(fn, d, name):
    ncls, per_im, nms = d[name + "_cfg"]
    ncls = int(ncls)
    idx = d[name + "_idx"] if name + "_idx" in d.files else None
    s, b, cb, ci = fn(d[name + "_scores"], d[name + "_boxes"], idx, ncls, 0.05, float(nms), int(per_im))
    assert np.array_equal(s, d[name + "_out_scores"]) and np.array_equal(b, d[name + "_out_boxes"]), name
    for j in range(1, ncls):
        assert np.array_equal(cb[j], d["%s_cls%d" % (name, j)]), (name, j)
        if idx is not None:
            assert np.array_equal(ci[j], d["%s_idx%d" % (name, j)]), (name, j)


def test_box_results_golden(golden):
    """lib/core/test.py:806-878 compiled from the reference source generated the fixture."""
    d = golden("box_results.npz")
    for name in "abc":
        _box_results_check(oracle.box_results_with_nms_and_limit, d, name)


# ------------------------------------------------------------------------------------------ Mask R-CNN mask paste-back
def _segm_case(g, name):
    ncls, cls_specific, M, S, H, W = (int(v) for v in g[name + "_cfg"])
    counts = [int(c) for c in g[name + "_counts"]]
    masks = g[name + "_masks_u8"].astype(np.float32) / np.float32(256)
    cls_boxes = [[]] + [np.zeros((c, 7), np.float32) for c in counts[1:]]
    vols = np.unpackbits(g[name + "_vols_bits"], axis=1)[:, :S * H * W].reshape(-1, S, H, W)
    return dict(ncls=ncls, cls_specific=bool(cls_specific), M=M, shape=(S, H, W), counts=counts, masks=masks,
                boxes=g[name + "_boxes"], cls_boxes=cls_boxes, vols=vols)


def test_segm_resize_restatement_is_the_scipy_calls():
    """skimage.transform.resize == gaussian_filter('mirror') + zoom(order 1, 'mirror', grid_mode) + clip: the restatement
    must be bit-identical to those scipy calls, upsampling and (anti-aliased) downsampling, every output size 1..40."""
    rng = np.random.default_rng(31)
    shapes = [(o, 16 + (o * 7) % 23, 1 + (o * 5) % 37) for o in range(1, 41)] + [(3, 60, 2), (47, 5, 16), (16, 16, 16)]
    for n_in in (16, 9):
        img = np.zeros((n_in,) * 3, np.float32)
        img[1:-1, 1:-1, 1:-1] = rng.random((n_in - 2,) * 3, dtype=np.float32)
        for shp in shapes:
            a = oracle.resize_reflect_antialias(img, shp)
            b = oracle.scipy_resize_reflect_antialias(img, shp)
            assert a.dtype == b.dtype == np.float32 and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n_in, shp)


def test_segm_results_golden(golden):
    g = golden("segm.npz")
    for name in ("a", "b"):
        c = _segm_case(g, name)
        out = oracle.segm_results(c["cls_boxes"], c["masks"], c["boxes"], *c["shape"], num_classes=c["ncls"],
                                  cls_specific_mask=c["cls_specific"])
        vols = np.stack([v for j in range(1, c["ncls"]) for v in out[j]])
        assert [len(out[j]) for j in range(c["ncls"])] == c["counts"]
        assert vols.sum() > 200 and np.array_equal(vols, c["vols"]), name


# ------------------------------------------------------------------------------------------ nuclei per-instance chain
def _nuclei_crops(g):
    import zlib
    out = []
    for k in range(int(g["count"])):
        img, prm = g["n%d_img" % k], g["n%d_prm" % k]
        mask = np.unpackbits(g["n%d_mask" % k])[:img.size].reshape(img.shape).astype(bool)
        out.append(dict(img=img, prm=prm, b=int(g["n%d_b" % k]), crc=[int(v) for v in g["n%d_crc16" % k]], mask=mask, crc32=zlib.crc32))
    return out


def test_nuclei_chain_golden(golden):
    """oracle.binarize_nuclei against binarization_nuclei.py:110-139 executed from the reference file (fixture)."""
    crops = _nuclei_crops(golden("nuclei.npz"))
    assert len(crops) >= 6 and {c["img"].dtype for c in crops} == {np.dtype(np.uint8), np.dtype(np.uint16)}
    for k, c in enumerate(crops):
        i16, p16 = oracle.nuclei_normalise(c["img"], c["prm"])
        assert [c["crc32"](np.ascontiguousarray(i16).tobytes()), c["crc32"](np.ascontiguousarray(p16).tobytes())] == c["crc"], k
        assert oracle.otsu_py_2d_fast(i16, p16)[2] == c["b"], k
        S, H, W = c["img"].shape
        seg, status, survive, masks = oracle.binarize_nuclei(c["img"], np.array([[0, 0, 0, W - 1, H - 1, S - 1]]), [c["prm"]])
        assert status == [0] and survive == [True] and np.array_equal(masks[0], c["mask"]) and np.array_equal(seg > 0, c["mask"]), k


# ------------------------------------------------------------------------------------------ evaluation records
def _eval_images(g):
    return [dict(gt=g["img%d_gt" % k], pred=g["img%d_pred" % k], gt_boxes=g["img%d_gt_boxes" % k], det_boxes=g["img%d_det_boxes" % k],
                 score=g["img%d_score" % k], tp=g["img%d_tp" % k], fp=g["img%d_fp" % k], pixels=g["img%d_pixels" % k])
            for k in range(int(g["count"]))]


def test_eval_records_golden(golden):
    """oracle.eval_volume_soma / eval_volume_nuclei against the reference's evaluation code run on in-memory volumes (fixture)."""
    g = golden("eval.npz")
    imgs = _eval_images(g)
    score, match, n_pos = [], [], 0
    for im in imgs:
        s, m, n = oracle.eval_volume_soma(im["pred"], im["gt"], im["score"], 0.3)
        score += s; match += m; n_pos += n
        tp, fp, tpp, gtp, prp = oracle.eval_volume_nuclei(im["pred"], im["gt"], im["det_boxes"], im["gt_boxes"], 0.4)
        assert np.array_equal(tp, im["tp"]) and np.array_equal(fp, im["fp"]) and [tpp, gtp, prp] == im["pixels"].tolist()
    order = np.asarray(score).argsort()[::-1]
    m = np.asarray(match, np.int8)[order]
    tp, fp = np.cumsum(m == 1), np.cumsum(m == 0)
    assert np.array_equal(tp / (fp + tp), g["prec"]) and np.array_equal(tp / n_pos, g["rec"])
    assert sum(int(im["tp"].sum()) for im in imgs) >= 4 and sum(int(im["fp"].sum()) for im in imgs) >= 2


def _nuclei_script_inputs(g, nms_volume):
    """Selection (:72-87), clamped boxes (:96-106) and PRM crops of the fixture volume, through the package's host helpers;
    nms_volume = the NMS to use (oracle on CPU, CUDA on the GPU)."""
    from b200seg import binarization_nuclei as bn
    dets, idx5, tiles = g["dets"], g["instance_idex"], g["tiles"]
    W, norm_side, S = (int(v) for v in g["cfg"])
    sel = np.nonzero(bn.nuclei_edge_filter(dets, W))[0]
    d = np.ascontiguousarray(dets[sel], dtype=np.float32)
    keep = np.asarray(nms_volume(d, 0.15), dtype=np.int64)
    sel, d = sel[keep], d[keep]
    sel = sel[d[:, -1] > 0.4]
    boxes = bn.nuclei_boxes(dets[sel], idx5[sel][:, 2:5], norm_side, S)
    crops = []
    for i, b in zip(sel, boxes):
        w, h = idx5[i, 2], idx5[i, 3]
        crops.append(np.ascontiguousarray(tiles[idx5[i, 0]][b[2]:b[5] + 1, b[1] - h:b[4] - h + 1, b[0] - w:b[3] - w + 1]))
    return sel, boxes, crops


def test_nuclei_script_golden(golden):
    """binarization_nuclei.py:73-148 executed from the reference file (selection, tile-relative clamping, per-instance chain,
    first-come paste, survivor table) against the host helpers + the oracle chain."""
    from b200seg import binarization_nuclei as bn
    g = golden("nuclei_script.npz")
    sel, boxes, crops = _nuclei_script_inputs(g, oracle.nms_3d_volume)
    assert np.array_equal(g["dets"][sel], g["visited_dets"]) and len(sel) >= 6
    seg, status, survive, _ = oracle.binarize_nuclei(g["img"], boxes, crops)
    assert status == [0] * len(sel) and np.array_equal(seg, g["seg"])
    rows = bn.id_det_rows(boxes, g["dets"][sel, -1], survive)
    assert rows.dtype == g["id_det"].dtype and np.array_equal(rows, g["id_det"])


def test_soma_script_golden(golden):
    """binarization_soma.py:57-104 executed from the reference file (NMS, descending-score visit order, mask ids that count
    skipped instances, tile crops, normalisation, Otsu, largest component, first-come paste, score table) against the
    oracle chain the GPU tests use as their checker."""
    from helpers import oracle_chain
    g = golden("soma_script.npz")
    case = dict(volume=g["img"], dets=g["dets"], boxes=g["boxes"], prm=g["prm"], crop_off=g["crop_off"])
    oc = oracle_chain(case, nms_thresh=0.23, keep_largest_cc=True)
    assert np.array_equal(g["dets"][oc["order"]], g["visited_dets"])
    assert np.array_equal(oc["seg"], g["seg"]) and len(np.unique(g["seg"])) > 6
    alive = np.asarray(oc["survive"][:len(oc["order"])], bool)
    ids = np.arange(1, len(oc["order"]) + 1)[alive]
    rows = np.stack([ids.astype(np.float64), g["dets"][oc["order"], 6][alive].astype(np.float64)], axis=1)
    assert np.array_equal(rows, g["scores"]) and 3 in oc["status"].values()          # an empty-PRM instance was skipped


# ------------------------------------------------------------------------------------------ round 2: the oracle against the
# reference's own Python, live (the staged files of oracle/_ref/py travel with the repo; skipped when they were never staged)
def test_oracle_otsu_vs_reference_python_1000_crops():
    """oracle.otsu_py_2d_fast == tools/otsu.py:otsu_py_2d_fast (b_max and mask) on >= 10^3 seeded crops: soma-normalised blob
    crops, few-level crops with exact ties in the criterion, noise and correlated attributes (same generator family as the GPU
    stress test, which checks the CUDA path against the oracle on >= 10^4 crops)."""
    from oracle import refpy
    if not refpy.available():
        pytest.skip("oracle/_ref/py not staged")
    ref = refpy.load_otsu().otsu_py_2d_fast
    rng = np.random.default_rng(2718)
    old = np.seterr(all="ignore")
    n_ok = n_raise = 0
    try:
        for i in range(1000):
            shape = tuple(int(v) for v in rng.integers(3, 9, 3))
            kind = i % 4
            if kind == 0:
                zz, yy, xx = np.meshgrid(*[np.arange(s_) for s_ in shape], indexing="ij")
                c = [s_ / 2 + rng.uniform(-1, 1) for s_ in shape]
                r2 = sum(((g_ - c_) / (s_ / 3 + 0.5)) ** 2 for g_, c_, s_ in zip((zz, yy, xx), c, shape))
                v = np.clip(rng.uniform(0, 40, shape) + rng.uniform(80, 200) * np.exp(-0.5 * r2), 0, 255).astype(np.uint8)
                p = (255 * np.exp(-0.5 * r2 / 0.64)).astype(np.uint8)
                if p.max() == 0 or v.max() == 0:
                    p[tuple(s_ // 2 for s_ in shape)] = 200; v[tuple(s_ // 2 for s_ in shape)] = 100
                # fewer levels than the soma range keep the reference's Python loop over b short
                a = (v // 4).astype(np.uint16) + 30; b = (p // 4).astype(np.uint16) + 30
            elif kind == 1:
                a = rng.integers(0, int(rng.integers(2, 6)), shape).astype(np.uint16) * int(rng.integers(1, 12)) + int(rng.integers(0, 300))
                b = rng.integers(0, int(rng.integers(2, 6)), shape).astype(np.uint16) * int(rng.integers(1, 12)) + int(rng.integers(0, 300))
            elif kind == 2:
                a = rng.integers(0, int(rng.integers(5, 70)), shape).astype(np.uint16)
                b = rng.integers(0, int(rng.integers(5, 70)), shape).astype(np.uint16)
            else:
                a = rng.integers(0, 60, shape).astype(np.uint16)
                b = (a // int(rng.integers(1, 5)) + rng.integers(0, 3, shape)).astype(np.uint16)
            try:
                rm, _, rb = ref(a, b)
            except (NameError, UnboundLocalError):
                with pytest.raises(UnboundLocalError):
                    oracle.otsu_py_2d_fast(a, b)
                n_raise += 1
                continue
            om, _, ob = oracle.otsu_py_2d_fast(a, b)
            assert ob == rb, (i, ob, rb)
            assert np.array_equal(om, rm), i
            n_ok += 1
    finally:
        np.seterr(**old)
    assert n_ok >= 900


def test_oracle_chain_vs_reference_script_live():
    """The oracle chain that checks every GPU chain test == the reference script's own lines 57-104 executed on a fresh seeded
    case with the reference's Cython NMS and tools/otsu.py (a second, independent case besides tests/golden/soma_script.npz)."""
    import sys, os
    from oracle import refpy
    if not refpy.available():
        pytest.skip("oracle/_ref/py not staged")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import oracle_chain
    # no false boxes: the script indexes with the raw int() box and breaks on boxes that leave the volume (negative start
    # = empty numpy slice); the detector clips boxes to the tile (boxes_3d.py:144-225), dets_to_boxes mirrors that clip
    c = synth.postproc_case(77, shape=(24, 80, 96), n_blobs=6, n_dup=3, n_false=0)
    c["volume"] = (c["volume"] // 16 * 16).astype(np.uint8)       # coarse gray levels keep the Python Otsu loop short
    r = refpy.run_soma_script(c, 0.23)
    o = oracle_chain(c, 0.23)
    assert np.array_equal(r["seg"], o["seg"])
    assert np.array_equal(r["visited_dets"], c["dets"][o["order"]])
    ids = np.flatnonzero(o["survive"]) + 1
    assert np.array_equal(r["scores"][:, 0], ids.astype(np.float64))
