"""GPU parity tests proper: every CUDA path, called through the C ABI (ctypes), against the oracle
on seeded inputs and against the committed golden vectors produced by the reference's own code.
Bit-exact for integer/index/byte results; RoIAlign3D within 1e-5 relative (fp32)."""
import ctypes

import numpy as np
import pytest

import oracle
from helpers import oracle_chain

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2(built_lib, cuda):
    import b200seg
    return b200seg


@pytest.fixture(scope="module")
def torch_(cuda):
    import torch
    return torch


# ------------------------------------------------------------------------------------------ NMS
def test_nms_golden(b2, golden):
    g = golden("nms_iou.npz")
    ids = sorted({int(k[3:].split("_")[0]) for k in g.files if k.startswith("nms")})
    for i in ids:
        d, thr = g["nms%d_dets" % i], g["nms%d_thr" % i]
        keep = b2.nms_3d(d, thr)
        assert keep.dtype == np.int64 and np.array_equal(keep, g["nms%d_keep" % i]), i
        if "nms%d_keepvol" % i in g.files:
            assert np.array_equal(b2.nms_3d_volume(d, thr), g["nms%d_keepvol" % i]), i


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 64, 65, 127, 200, 1000, 2500, 6000])
def test_nms_vs_oracle_sizes(b2, n):
    from b200seg import synth
    rng = np.random.default_rng(100 + n)
    for rep in range(3 if n <= 1000 else 1):
        d = synth.random_dets(rng, n, extent=(60 + 3 * rep * 40, 200, 64), integer=(rep == 1))
        for thr in (0.15, 0.23, 0.7):
            assert np.array_equal(b2.nms_3d(d, thr), oracle.nms_3d(d, thr)), (n, rep, thr)
        assert np.array_equal(b2.nms_3d_volume(d, 0.15), oracle.nms_3d_volume(d, 0.15)), (n, rep)


def test_nms_ties_and_degenerate(b2):
    rng = np.random.default_rng(5)
    from b200seg import synth
    d = synth.random_dets(rng, 300, extent=(80, 80, 40), integer=True)
    d[:, 6] = (rng.integers(0, 4, 300) / 4).astype(np.float32)           # heavy score ties: documented tie rule
    assert np.array_equal(b2.nms_3d(d, 0.23), oracle.nms_3d(d, 0.23))
    d2 = np.repeat(d[:1], 70, axis=0)                                    # identical boxes and scores
    assert np.array_equal(b2.nms_3d(d2, 0.5), oracle.nms_3d(d2, 0.5)) and len(b2.nms_3d(d2, 0.5)) == 1
    d3 = d.copy(); d3[:, 3:6] = d3[:, :3] - 5                            # inverted boxes (negative extents)
    assert np.array_equal(b2.nms_3d(d3, 0.23), oracle.nms_3d(d3, 0.23))
    assert b2.nms_3d(np.zeros((0, 7), np.float32), 0.3) == []


def test_nms_batched_device_and_rank_order(b2, torch_):
    from b200seg import synth
    rng = np.random.default_rng(77)
    sets = [synth.random_dets(rng, n, extent=(150, 150, 60)) for n in (800, 0, 1, 257, 800, 64)]
    off = np.zeros(len(sets) + 1, np.int32); off[1:] = np.cumsum([s.shape[0] for s in sets])
    dets = torch_.from_numpy(np.concatenate(sets)).cuda()
    keep, cnt, rank = b2.nms_3d_batched(dets, torch_.from_numpy(off).cuda(), 800, 0.23, want_rank_order=True)
    keep, cnt, rank = keep.cpu().numpy(), cnt.cpu().numpy(), rank.cpu().numpy()
    for b, s in enumerate(sets):
        ok = oracle.nms_3d(s, 0.23)
        assert cnt[b] == len(ok)
        assert np.array_equal(keep[off[b]:off[b] + cnt[b]], ok)
        exp_rank = ok[oracle.argsort_desc(s[ok, 6])] if len(ok) else ok
        assert np.array_equal(rank[off[b]:off[b] + cnt[b]], exp_rank)
    # tensor seam
    k2 = b2.nms_3d(torch_.from_numpy(sets[0]).cuda(), 0.23)
    assert np.array_equal(k2.cpu().numpy(), oracle.nms_3d(sets[0], 0.23))


# ------------------------------------------------------------------------------------------ IoU
def test_iou_golden_and_random(b2, golden, torch_):
    g = golden("nms_iou.npz")
    for i in range(4):
        out = b2.bbox_overlaps_3d(g["iou%d_boxes" % i], g["iou%d_query" % i])
        assert out.dtype == np.float32
        assert np.array_equal(out.view(np.uint32), g["iou%d_out" % i].view(np.uint32)), i
    from b200seg import synth
    rng = np.random.default_rng(9)
    for (N, K, integer) in [(1, 1, False), (5, 3, True), (1000, 50, False), (4097, 7, True), (33, 1025, False)]:
        b = synth.random_dets(rng, N, extent=(120, 120, 50), integer=integer)[:, :6].copy()
        q = synth.random_dets(rng, K, extent=(120, 120, 50), integer=integer)[:, :6].copy()
        q[: min(N, K)] = b[: min(N, K)]
        o = oracle.bbox_overlaps_3d(b, q)
        assert np.array_equal(b2.bbox_overlaps_3d(b, q).view(np.uint32), o.view(np.uint32)), (N, K)
        t = b2.bbox_overlaps_3d(torch_.from_numpy(b).cuda(), torch_.from_numpy(q).cuda())
        assert np.array_equal(t.cpu().numpy().view(np.uint32), o.view(np.uint32))
    assert b2.bbox_overlaps_3d(np.zeros((0, 6), np.float32), np.zeros((4, 6), np.float32)).shape == (0, 4)


def test_iou_anchor_sized_properties(b2, torch_):
    """BASELINE-sized call (917 504 anchors x 50 gt): sampled rows equal the oracle, result is symmetric."""
    from b200seg import synth
    rng = np.random.default_rng(10)
    N, K = 917504, 50
    b = synth.random_dets(rng, N, extent=(256, 256, 64), side=(8, 64))[:, :6].copy()
    q = synth.random_dets(rng, K, extent=(256, 256, 64), side=(10, 40))[:, :6].copy()
    out = b2.bbox_overlaps_3d(torch_.from_numpy(b).cuda(), torch_.from_numpy(q).cuda())
    rows = rng.choice(N, 2000, replace=False)
    assert np.array_equal(out[torch_.from_numpy(rows).cuda()].cpu().numpy().view(np.uint32),
                          oracle.bbox_overlaps_3d(b[rows], q).view(np.uint32))
    outT = b2.bbox_overlaps_3d(torch_.from_numpy(q).cuda(), torch_.from_numpy(b[:4096]).cuda())
    ref = oracle.bbox_overlaps_3d(q, b[:4096])
    assert np.array_equal(outT.cpu().numpy().view(np.uint32), ref.view(np.uint32))
    assert float(out.max()) <= 1.0 and float(out.min()) >= 0.0


# ------------------------------------------------------------------------------------------ Otsu
def test_otsu_golden(b2, golden):
    g = golden("otsu.npz")
    for i in range(int(g["count"])):
        img, prm = g["otsu%d_img" % i], g["otsu%d_prm" % i]
        mask, k, b = b2.otsu_py_2d_fast(img, prm)
        assert (k, b) == (-1, int(g["otsu%d_b" % i])), i
        assert mask.dtype == np.uint8 and mask.shape == img.shape
        assert np.array_equal(np.packbits(mask.ravel() > 0), g["otsu%d_mask" % i]), i


def _random_crops(rng, n, lo=4, hi=36, u16=True):
    from b200seg import synth
    crops = []
    while len(crops) < n:
        shape = tuple(int(v) for v in rng.integers(lo, hi, 3))
        vol, _, blobs = synth.blob_volume(rng, shape, 1, sigma_xy=(shape[1] / 6 + .5, shape[1] / 3 + 1),
                                          sigma_z=(shape[0] / 6 + .5, shape[0] / 3 + 1), noise=int(rng.integers(5, 60)))
        prm = synth.prm_crop(blobs[0], (0, 0, 0, shape[2] - 1, shape[1] - 1, shape[0] - 1))
        if prm.max() == 0:
            continue
        crops.append((vol, prm))
    return crops


def test_otsu_batch_vs_oracle_hist_threshold_mask(b2, torch_):
    rng = np.random.default_rng(2024)
    raw = _random_crops(rng, 160)
    imgs, prms = [], []
    for i, (v, p) in enumerate(raw):
        if i % 3 == 0:
            a, b = v.astype(np.uint16), p.astype(np.uint16)
        else:
            a, b = oracle.soma_normalise(v, p)
        if i % 17 == 0:
            b = np.full_like(b, 123)                    # constant second attribute
        if i % 23 == 0:
            a = a * 3 + 1000                            # wider, offset gray range
        imgs.append(a); prms.append(b)
    imgs.append(np.full((3, 4, 5), 42, np.uint16)); prms.append(np.full((3, 4, 5), 7, np.uint16))   # constant crop
    off = np.zeros(len(imgs) + 1, np.int64); off[1:] = np.cumsum([a.size for a in imgs])
    I = torch_.from_numpy(np.concatenate([a.ravel() for a in imgs])).cuda()
    P = torch_.from_numpy(np.concatenate([a.ravel() for a in prms])).cuda()
    out = b2.otsu_2d_batch(I, P, torch_.from_numpy(off).cuda(), want_hist=True)
    mask, bmax, status = out["mask"].cpu().numpy(), out["b_max"].cpu().numpy(), out["status"].cpu().numpy()
    for i, (a, p) in enumerate(zip(imgs, prms)):
        try:
            om, _, ob, oh = oracle.otsu_py_2d_fast(a, p, want_hist=True)
        except UnboundLocalError:
            assert status[i] == 1 and (mask[off[i]:off[i + 1]] == 255).all(), i
            continue
        assert status[i] == 0 and bmax[i] == ob, (i, bmax[i], ob)
        assert np.array_equal(mask[off[i]:off[i + 1]].reshape(a.shape), om), i
        assert np.array_equal(out["hist"][i].astype(np.int64), oh.astype(np.int64)), i
    assert status[-1] == 1
    with pytest.raises(UnboundLocalError):
        b2.otsu_py_2d_fast(imgs[-1], prms[-1])


def test_soma_binarize_vs_oracle(b2, torch_):
    from b200seg import synth
    case = synth.postproc_case(31, shape=(40, 128, 160), n_blobs=20, n_dup=5, n_false=5)
    case["prm"][case["crop_off"][3]:case["crop_off"][4]] = 0                 # instance with empty PRM (skipped)
    vol = torch_.from_numpy(case["volume"]).cuda()
    mask, bmax, status = b2.soma_binarize(vol, torch_.from_numpy(case["boxes"]).cuda(), torch_.from_numpy(case["prm"]).cuda(),
                                          torch_.from_numpy(case["crop_off"]).cuda())
    mask, bmax, status = mask.cpu().numpy(), bmax.cpu().numpy(), status.cpu().numpy()
    off = case["crop_off"]
    for i, b in enumerate(case["boxes"]):
        img = case["volume"][b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1]
        p = case["prm"][off[i]:off[i + 1]].reshape(img.shape)
        if p.max() == 0:
            assert status[i] == 3 and not mask[off[i]:off[i + 1]].any()
            continue
        i16, p16 = oracle.soma_normalise(img, p)
        try:
            om, _, ob = oracle.otsu_py_2d_fast(i16, p16)
        except UnboundLocalError:
            assert status[i] == 1
            continue
        assert status[i] == 0 and bmax[i] == ob, i
        assert np.array_equal(mask[off[i]:off[i + 1]].reshape(img.shape), om), i


# ------------------------------------------------------------------------------------------ paste + chain
def test_largest_cc_vs_oracle(b2, torch_):
    rng = np.random.default_rng(77)
    shapes = [(5, 9, 11), (12, 20, 30), (6, 7, 3), (3, 4, 1), (10, 12, 70), (4, 6, 130), (40, 60, 70), (18, 40, 44), (1, 1, 1), (2, 3, 64), (2, 3, 65)]
    masks = []
    for i, sh in enumerate(shapes):
        zz, yy, xx = np.meshgrid(*[np.arange(s) for s in sh], indexing="ij")
        blob = ((zz - sh[0] / 2) ** 2 / max(sh[0] / 3, 1) ** 2 + (yy - sh[1] / 2) ** 2 / max(sh[1] / 3, 1) ** 2 +
                (xx - sh[2] / 2) ** 2 / max(sh[2] / 3, 1) ** 2) < 1
        speck = rng.random(sh) < (0.5 if i == 6 else 0.08)          # crop 6: > 2560 runs and > 2048 rows -> global scratch
        m = (blob & (rng.random(sh) < 0.9)) | speck
        masks.append((m * 255).astype(np.uint8))
    masks.append(np.zeros((4, 5, 6), np.uint8))                     # no foreground: status 5
    tie = np.zeros((3, 3, 9), np.uint8); tie[0, 0, 0:2] = 255; tie[2, 2, 6:8] = 255                # two components of 2 voxels
    masks.append(tie)
    i_tie = len(masks) - 1
    diag = np.zeros((3, 3, 3), np.uint8); diag[0, 0, 0] = diag[1, 1, 1] = diag[2, 2, 2] = 255; diag[0, 2, 0] = 255   # 26-connectivity
    masks.append(diag)
    # every way out of the flood-fill fast path (largest_cc.cu): donut (empty centre row), two blobs where the centre one is
    # the smaller (no majority), a spiral corridor (more sweep pairs than the cap), 30 planes x 61 rows (row words in
    # the global scratch), 33 planes (too deep for one lane per plane), and a filled block (one sweep pair)
    zz, yy, xx = np.meshgrid(np.arange(9), np.arange(21), np.arange(21), indexing="ij")
    rr = np.sqrt((yy - 10) ** 2 + (xx - 10) ** 2)
    masks.append((((rr > 4) & (rr < 8)) * 255).astype(np.uint8))
    two = np.zeros((8, 30, 40), np.uint8); two[2:6, 12:18, 17:23] = 255; two[1:7, 1:9, 1:38] = 255; two[0, 29, 39] = 255
    masks.append(two)
    spiral = np.zeros((3, 41, 41), np.uint8)
    y, x, dy, dx, n = 20, 20, 0, 1, 1
    for leg in range(38):                                           # unit-width arms two voxels apart: a single long corridor
        for _ in range(n):
            spiral[1, y, x] = 255; y += dy; x += dx
        dy, dx = dx, -dy
        n += 2 * (leg % 2)
    masks.append(spiral)
    big = (rng.random((30, 61, 40)) < 0.7).astype(np.uint8) * 255; big[:, 30, :] = 255
    masks.append(big)
    deep = np.zeros((33, 6, 7), np.uint8); deep[:, 2:4, 2:5] = 255; deep[0, 0, 0] = 255
    masks.append(deep)
    masks.append(np.full((6, 10, 64), 255, np.uint8))
    boxes = np.array([[0, 0, 0, m.shape[2] - 1, m.shape[1] - 1, m.shape[0] - 1] for m in masks], np.int32)
    off = np.zeros(len(masks) + 1, np.int64); off[1:] = np.cumsum([m.size for m in masks])
    flat = torch_.from_numpy(np.concatenate([m.ravel() for m in masks])).cuda()
    from b200seg import _lib
    counts = (ctypes.c_longlong * 8)()
    _lib.lib().b200seg_largest_cc_path_counts(counts, 1)
    status = b2.largest_cc(flat, torch_.from_numpy(off).cuda(), torch_.from_numpy(boxes).cuda()).cpu().numpy()
    out = flat.cpu().numpy()
    _lib.lib().b200seg_largest_cc_path_counts(counts, 0)
    assert sum(counts) == len(masks) and all(c > 0 for c in list(counts)[:6]), list(counts)      # every path was exercised
    for i, m in enumerate(masks):
        got = out[off[i]:off[i + 1]].reshape(m.shape)
        if not m.any():
            assert status[i] == 5 and not got.any(), i
            continue
        ref = oracle.largest_cc(m)
        assert status[i] == 0, (i, status[i])
        assert np.array_equal(got != 0, ref), (i, m.shape, int((got != 0).sum()), int(ref.sum()))
        assert np.array_equal(got[ref], m[ref]), i                  # kept voxels keep their value
    assert out[off[i_tie]:off[i_tie + 1]].reshape(tie.shape)[2, 2, 6] == 255   # tie: the later component wins


def test_mask_overlaps_vs_oracle_and_golden(b2, golden, torch_):
    from b200seg import evaluation
    d = golden("mask_iou.npz")
    r = evaluation.mask_overlaps_labels(d["pred"], d["gt"], d["pred_ids"], d["gt_ids"])
    for k in ("iou", "ios", "iog"):
        assert np.array_equal(r[k].cpu().numpy(), d[k], equal_nan=True), k           # bit-exact float32(double / double)
    pa = np.stack([d["pred"] == i for i in d["pred_ids"]]); ga = np.stack([d["gt"] == i for i in d["gt_ids"]])
    assert np.array_equal(evaluation.mask_iou_fast(pa, ga), d["iou"], equal_nan=True)   # the reference's stack signature
    # random label volumes, odd sizes (scalar tail), ids listed in arbitrary order, unlisted ids, an empty listed id
    rng = np.random.default_rng(11)
    for shape, npred, ngt in [((7, 13, 19), 5, 4), ((16, 40, 50), 30, 25), ((33, 65, 70), 200, 180)]:
        pred = np.zeros(shape, np.uint16); gt = np.zeros(shape, np.uint16)
        for lab, n in ((pred, npred), (gt, ngt)):
            for i in range(1, n + 1):
                c = [rng.integers(0, s) for s in shape]; e = [rng.integers(1, max(2, s // 3)) for s in shape]
                lab[c[0]:c[0] + e[0], c[1]:c[1] + e[1], c[2]:c[2] + e[2]] = i
        pid = rng.permutation(np.arange(1, npred + 1))[: max(1, npred - 2)]           # two ids stay unlisted
        gid = np.concatenate([rng.permutation(np.arange(1, ngt + 1)), [ngt + 7]])     # one listed id that never occurs
        r = evaluation.mask_overlaps_labels(pred, gt, pid, gid)
        pa = np.stack([pred == i for i in pid]); ga = np.stack([gt == i for i in gid])
        iou, ios, iog = oracle.mask_overlaps(pa, ga)
        assert np.array_equal(r["iou"].cpu().numpy(), iou, equal_nan=True)
        assert np.array_equal(r["ios"].cpu().numpy(), ios, equal_nan=True)
        assert np.array_equal(r["iog"].cpu().numpy(), iog, equal_nan=True)
        assert np.array_equal(r["area_pred"].cpu().numpy(), pa.reshape(len(pid), -1).sum(1))
        assert np.array_equal(r["area_gt"].cpu().numpy(), ga.reshape(len(gid), -1).sum(1))
    # full size: checksum property -- the table accounts for every voxel that is not background in both volumes
    S, H, W = 128, 512, 512
    pred = torch_.zeros((S, H, W), dtype=torch_.int32, device="cuda"); gt = torch_.zeros_like(pred)
    pred[10:100, 50:400, 60:300] = 3; pred[20:60, 10:40, 10:500] = 9; gt[30:120, 100:450, 100:350] = 5
    r = evaluation.mask_overlaps_labels(pred.to(torch_.uint16), gt.to(torch_.uint16), [3, 9], [5])
    assert int(r["inter"][0, 0]) == int(((pred == 3) & (gt == 5)).sum()) and int(r["area_pred"][1]) == int((pred == 9).sum())


def test_rle_codec_vs_oracle_and_golden(b2, golden, torch_):
    from b200seg import mask_3d
    d = golden("rle.npz")
    r = mask_3d.binary_mask_to_rle(d["mask"])
    assert r["counts"] == d["counts"].tolist() and r["size"] == d["size"].tolist()      # mask_3d.py:75-79
    assert np.array_equal(mask_3d.rle_to_binary_mask(r), d["mask"])
    rng = np.random.default_rng(5)
    for sh, dens in [((3, 4, 5), 0.3), ((1, 1, 1), 1.0), ((1, 1, 1), 0.0), ((7, 2, 9), 0.5), ((16, 33, 70), 0.02), ((16, 33, 70), 0.0),
                     ((16, 33, 70), 1.0), ((40, 64, 96), 0.5), ((128, 40, 50), 0.001)]:
        m = (rng.random(sh) < dens).astype(np.uint8) * 255
        a = mask_3d.binary_mask_to_rle(m, cap=64)                # small first guess: exercises the exact second pass
        o = oracle.binary_mask_to_rle(m)
        assert a == o, (sh, dens, len(a["counts"]), len(o["counts"]))
        back = mask_3d.rle_to_binary_mask(a)
        assert back.dtype == np.uint8 and np.array_equal(back, (m != 0).astype(np.uint8))   # round trip
    blob = np.zeros((128, 256, 256), np.uint8); blob[30:90, 60:200, 50:180] = 1            # config-sized round trip
    a = mask_3d.binary_mask_to_rle(blob)
    assert sum(a["counts"]) == blob.size and np.array_equal(mask_3d.rle_to_binary_mask(a), blob)


def test_paste_vs_oracle_overlapping(b2, torch_):
    rng = np.random.default_rng(3)
    S, H, W = 20, 50, 77                                                     # W not a multiple of 8: scalar store path
    n = 120                                                                  # > PASTE_MAXL in one tile: overflow path
    boxes = np.zeros((n, 6), np.int32)
    masks = []
    for i in range(n):
        sz, sy, sx = rng.integers(1, 12), rng.integers(1, 30), rng.integers(1, 40)
        z1, y1, x1 = rng.integers(0, S - sz + 1), rng.integers(0, H - sy + 1), rng.integers(0, W - sx + 1)
        boxes[i] = [x1, y1, z1, x1 + sx - 1, y1 + sy - 1, z1 + sz - 1]
        masks.append((rng.random((sz, sy, sx)) < 0.6).astype(np.uint8) * 255)
    ids = np.arange(1, n + 1, dtype=np.uint16)
    off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum([m.size for m in masks])
    seg_o = np.zeros((S, H, W), np.uint16)
    surv_o = oracle.paste_labels(seg_o, boxes, ids, masks)
    seg = torch_.full((S, H, W), 999, dtype=torch_.uint16, device="cuda")   # no pre-clear needed
    surv = b2.paste_labels(seg, torch_.from_numpy(boxes).cuda(), torch_.from_numpy(ids).cuda(),
                           torch_.from_numpy(np.concatenate([m.ravel() for m in masks])).cuda(), torch_.from_numpy(off).cuda())
    assert np.array_equal(seg.cpu().numpy(), seg_o)
    assert np.array_equal(surv.cpu().numpy().astype(bool), surv_o)


@pytest.mark.parametrize("seed,shape,nb,cc", [(1001, (64, 256, 256), 35, True), (7, (33, 100, 130), 12, True),
                                              (1001, (64, 256, 256), 35, False)])
def test_postproc_chain_host_vs_oracle(b2, seed, shape, nb, cc):
    """BASELINE config 1: one 64x256x256 uint8 volume, exactly 50 boxes, NMS 0.23 + per-instance Otsu
    (+ largest connected component, binarization_soma.py:97-99) + paste."""
    from b200seg import synth
    case = synth.postproc_case(seed, shape=shape, n_blobs=nb)
    out = b2.postproc_soma_host(case["volume"], case["dets"], case["boxes"], case["prm"], case["crop_off"], 0.23,
                                keep_largest_cc=cc)
    ref = oracle_chain(case, 0.23, keep_largest_cc=cc)
    assert out["n_keep"] == len(ref["order"]) and np.array_equal(out["rank_order"], ref["order"])
    for i, st in ref["status"].items():
        assert out["status"][i] == st, i
    for i, bm in ref["b_max"].items():
        assert out["b_max"][i] == bm, i
    assert np.array_equal(out["seg"], ref["seg"])
    assert np.array_equal(out["survive"], ref["survive"])
    assert (out["status"][np.setdiff1d(np.arange(len(case["dets"])), ref["order"])] == -1).all()


def test_postproc_host_batch_matches_single_volume_calls(b2):
    from b200seg import synth
    cases = [synth.postproc_case(500 + i, shape=(24, 80, 96), n_blobs=6 + 3 * i, n_dup=3, n_false=2) for i in range(5)]
    cases.append(dict(cases[0], dets=np.zeros((0, 7), np.float32), boxes=np.zeros((0, 6), np.int32),
                      prm=np.zeros(0, np.uint8), crop_off=np.zeros(1, np.int64)))          # a volume without detections
    outs = b2.postproc_soma_host_batch(cases, 0.23)
    assert len(outs) == len(cases)
    for c, o in zip(cases, outs):
        ref = b2.postproc_soma_host(c["volume"], c["dets"], c["boxes"], c["prm"], c["crop_off"], 0.23)
        assert o["n_keep"] == ref["n_keep"]
        assert np.array_equal(o["seg"], ref["seg"])
        for k in ("rank_order", "b_max", "status", "survive", "scores"):
            assert np.array_equal(o[k], ref[k]), k
    assert outs[-1]["n_keep"] == 0 and not outs[-1]["seg"].any()


def test_postproc_host_batch_output_states(b2):
    """`host_batch_out`: what the label buffers hold on entry.  1 = zeros (fresh np.zeros, the reference's allocation),
    2 = the previous result of the same call (only what was written then is cleared; unknown buffers are zero-filled).
    Same buffers reused over three different batches, one of them with an all-background volume; a buffer the library
    has never seen, full of garbage, must still come out right in state 2."""
    from b200seg import synth
    from b200seg.binarization import set_host_batch_out
    from helpers import oracle_chain
    shape = (24, 80, 96)
    batches = [[synth.postproc_case(900 + 10 * b + i, shape=shape, n_blobs=3 + 2 * i + b, n_dup=3, n_false=2) for i in range(4)]
               for b in range(3)]
    empty = dict(batches[1][2])
    empty["dets"] = empty["dets"][:0]; empty["boxes"] = empty["boxes"][:0]; empty["prm"] = empty["prm"][:0]
    empty["crop_off"] = np.zeros(1, np.int64)
    batches[1][2] = empty
    refs = [[oracle_chain(c, 0.23)["seg"] if c["dets"].shape[0] else np.zeros(shape, np.uint16) for c in bt] for bt in batches]
    try:
        set_host_batch_out(2)
        segs = [np.full(shape, 0x1234, np.uint16) for _ in range(4)]          # never seen by the library: garbage allowed
        for bt, rf in zip(batches, refs):
            b2.postproc_soma_host_batch(bt, 0.23, seg_out=segs)
            for s_, r_ in zip(segs, rf):
                assert np.array_equal(s_, r_)
        segs[1][3, 5, 7] = 77                                                    # the caller breaks the contract ...
        set_host_batch_out(0)                                                    # ... and says so: state 0 repairs it
        b2.postproc_soma_host_batch(batches[0], 0.23, seg_out=segs)
        for s_, r_ in zip(segs, refs[0]):
            assert np.array_equal(s_, r_)
        set_host_batch_out(1)
        fresh = [np.zeros(shape, np.uint16) for _ in range(4)]
        b2.postproc_soma_host_batch(batches[2], 0.23, seg_out=fresh)
        for s_, r_ in zip(fresh, refs[2]):
            assert np.array_equal(s_, r_)
    finally:
        set_host_batch_out(0)
    with pytest.raises(b2.B200SegError):
        set_host_batch_out(3)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 7, 15, 23, 39, 47, 33])
def test_postproc_host_batch_transfer_modes_pinned(b2, torch_, mode):
    """Every transfer scheme of the batch entry point (dense / compacted label download x whole-array / gathered PRM
    upload; mode 7 also sends the image crops of the NMS survivors packed by host threads instead of the volume) returns
    the same outputs; pinned inputs make the zero-copy gather path eligible, a volume whose size is
    not a multiple of 8 voxels and an odd-aligned PRM buffer exercise the fallbacks."""
    from b200seg import synth
    from b200seg.binarization import set_host_batch_mode, host_batch_traffic
    from helpers import oracle_chain
    cases = [synth.postproc_case(700 + i, shape=(24, 80, 96), n_blobs=5 + 2 * i, n_dup=4, n_false=3) for i in range(7)]
    pin = lambda a: torch_.from_numpy(np.ascontiguousarray(a)).pin_memory()
    keep = []
    pinned = []
    for i, c in enumerate(cases):
        d = dict(c)
        for k in ("volume", "dets", "boxes", "crop_off"):
            t = pin(c[k]); keep.append(t); d[k] = t.numpy()
        if i == 3:                                  # pinned but not 16-byte aligned: whole-array copy
            t = torch_.empty(c["prm"].size + 1, dtype=torch_.uint8).pin_memory(); keep.append(t)
            t[1:] = torch_.from_numpy(c["prm"]); d["prm"] = t.numpy()[1:]
        elif i == 4:                                # pageable
            d["prm"] = c["prm"].copy()
        else:
            t = pin(c["prm"]); keep.append(t); d["prm"] = t.numpy()
        pinned.append(d)
    try:
        set_host_batch_mode(mode)
        segs = [np.full(c["volume"].shape, 0xABCD, np.uint16) for c in cases]      # stale contents must be overwritten
        outs = b2.postproc_soma_host_batch(pinned, 0.23, seg_out=segs)
        h2d, d2h = host_batch_traffic()
    finally:
        set_host_batch_mode(7)
    dense_down = sum(c["volume"].size * 2 for c in cases)
    full_up = sum(c["volume"].size + c["prm"].size for c in cases)
    assert (d2h < dense_down * 3 // 4) if (mode & 1) else (d2h >= dense_down)     # 64-byte lines: coarse on volumes this small
    assert (h2d < full_up) if (mode & 6) else (h2d >= full_up)      # bit 1: gathered PRM crops, bit 2: image and PRM crops packed by the host
    for c, o in zip(cases, outs):
        ref = oracle_chain(c, 0.23)
        assert np.array_equal(o["seg"], ref["seg"])
        assert np.array_equal(o["rank_order"], ref["order"])
        assert np.array_equal(o["survive"], ref["survive"])


def test_postproc_host_batch_odd_volume_size(b2):
    """S*H*W not a multiple of 8: the compacted download does not apply, the dense copy does."""
    from b200seg import synth
    from helpers import oracle_chain
    cases = [synth.postproc_case(720 + i, shape=(9, 35, 37), n_blobs=3, n_dup=1, n_false=1, sigma_xy=(2, 4), sigma_z=(1, 2)) for i in range(3)]
    outs = b2.postproc_soma_host_batch(cases, 0.23)
    for c, o in zip(cases, outs):
        ref = oracle_chain(c, 0.23)
        assert np.array_equal(o["seg"], ref["seg"])


@pytest.mark.parametrize("pinned", [False, True])
def test_postproc_host_batch_partial_last_line_and_unaligned_buffers(b2, torch_, pinned):
    """S*H*W a multiple of 8 but not of 32: the compacted download ends in a partial 64-byte line; label buffers that are
    not 64-byte aligned take the plain-copy form of the line writes.  Every output state, twice (the second call clears
    what the first one wrote).  Pinned label buffers are written in place by the GPU (seg_lines_to_host_kernel), pageable ones
    through the staged download and the host scatter; "host_batch_mode" bit 8 / bit 6 force the in-place / the staged form for
    pinned buffers (the default picks by the number of host cores per rank)."""
    from b200seg import synth
    from b200seg.binarization import set_host_batch_out, set_host_batch_mode
    from helpers import oracle_chain
    shape = (9, 36, 38)                                  # 12312 voxels = 1539 groups = 384 lines + 3 groups
    assert (9 * 36 * 38) % 8 == 0 and (9 * 36 * 38) % 32 != 0
    cases = [synth.postproc_case(730 + i, shape=shape, n_blobs=4, n_dup=1, n_false=1, sigma_xy=(2, 4), sigma_z=(1, 2)) for i in range(3)]
    cases[2]["volume"][-1, -1, -8:] = 255                # something bright in the very last groups of a volume
    refs = [oracle_chain(c, 0.23) for c in cases]
    V = int(np.prod(shape))
    try:
        for shift in (0, 8):                             # 64-byte aligned / 16-byte aligned only
            if pinned:
                keep = [torch_.empty(V + 64, dtype=torch_.int16).pin_memory() for _ in cases]
                raw = [t.numpy().view(np.uint16) for t in keep]
                for r in raw:
                    r[...] = 0xABCD
            else:
                raw = [np.full(V + 64, 0xABCD, np.uint16) for _ in cases]
            segs = []
            for r in raw:
                o = ((-r.ctypes.data) % 64) // 2 + shift
                segs.append(r[o:o + V].reshape(shape))
            # garbage -> result -> same result kept (twice) -> caller-zeroed -> staged form (garbage, kept) -> back to the default form
            for state, mode in ((0, 7 | 256), (2, 7 | 256), (2, 7 | 256), (1, 7 | 256), (0, 7 | 64), (2, 7 | 64), (2, 7 | 256), (2, 7)):
                set_host_batch_out(state)
                set_host_batch_mode(mode)
                if state == 1:
                    for sg in segs:
                        sg[...] = 0
                outs = b2.postproc_soma_host_batch(cases, 0.23, seg_out=segs)
                for o, ref, sg, r in zip(outs, refs, segs, raw):
                    assert np.array_equal(sg, ref["seg"]), (state, shift)
                    assert np.array_equal(o["seg"], ref["seg"])
                    off = (sg.ctypes.data - r.ctypes.data) // 2
                    assert (r[:off] == 0xABCD).all() and (r[off + V:] == 0xABCD).all(), "wrote outside the label volume"
    finally:
        set_host_batch_out(0)
        set_host_batch_mode(7)


def test_postproc_batched_device_and_fullsize_properties(b2, torch_):
    """Two 128x512x512 volumes (BASELINE config 3/5 size) through the device chain:
    volume 0 is checked voxel-exact against the oracle, both against size-independent properties."""
    from b200seg import synth
    shape = (128, 512, 512)
    cases = [synth.postproc_case(2000 + i, shape=shape, n_blobs=60, n_dup=20, n_false=10) for i in range(2)]
    counts = [c["dets"].shape[0] for c in cases]
    prm = np.concatenate([c["prm"] for c in cases])
    offs, base = [], 0
    for c in cases:
        offs.append(c["crop_off"][:-1] + base); base += c["crop_off"][-1]
    crop_off = np.concatenate(offs + [np.array([base], np.int64)])
    pp = b2.SomaPostproc(2, shape, counts, prm.size)
    vols = torch_.from_numpy(np.stack([c["volume"] for c in cases])).cuda()
    seg = pp.run(vols, torch_.from_numpy(np.concatenate([c["dets"] for c in cases])).cuda(),
                 torch_.from_numpy(np.concatenate([c["boxes"] for c in cases])).cuda(),
                 torch_.from_numpy(prm).cuda(), torch_.from_numpy(crop_off).cuda(), 0.23)
    torch_.cuda.synchronize()
    seg = seg.cpu().numpy()
    cnt = pp.keep_count.cpu().numpy()
    ref0 = oracle_chain(cases[0], 0.23)
    assert cnt[0] == len(ref0["order"]) and np.array_equal(seg[0], ref0["seg"])
    for v in range(2):
        labels = np.unique(seg[v])
        surv = pp.survive.cpu().numpy()[pp.det_off_host[v]:pp.det_off_host[v] + cnt[v]].astype(bool)
        assert np.array_equal(labels[labels > 0], np.nonzero(surv)[0] + 1)      # survivors == labels present
        order = pp.rank_order.cpu().numpy()[pp.det_off_host[v]:pp.det_off_host[v] + cnt[v]]
        for lab in labels[labels > 0][:10]:                                     # every label lies inside its own box
            b = cases[v]["boxes"][order[lab - 1]]
            zz, yy, xx = np.nonzero(seg[v] == lab)
            assert zz.min() >= b[2] and zz.max() <= b[5] and yy.min() >= b[1] and yy.max() <= b[4] and xx.min() >= b[0] and xx.max() <= b[3]
    # idempotence: running the chain again gives the same volume
    seg2 = pp.run(vols, torch_.from_numpy(np.concatenate([c["dets"] for c in cases])).cuda(),
                  torch_.from_numpy(np.concatenate([c["boxes"] for c in cases])).cuda(),
                  torch_.from_numpy(prm).cuda(), torch_.from_numpy(crop_off).cuda(), 0.23).cpu().numpy()
    assert np.array_equal(seg, seg2)


# ------------------------------------------------------------------------------------------ peaks
def test_peaks_golden(b2, golden, torch_):
    from b200seg.peak_stimulation_3d import peak_stimulation_3d, median_filter
    g = golden("peaks.npz")
    for i in range(int(g["count"])):
        x, win = torch_.from_numpy(g["pk%d_in" % i]).cuda(), int(g["pk%d_win" % i])
        pl, agg = peak_stimulation_3d(x, win_size=win, peak_filter=median_filter)
        assert pl.dtype == torch_.int64 and np.array_equal(pl.cpu().numpy(), g["pk%d_peaks_med" % i]), i
        np.testing.assert_allclose(agg.cpu().numpy(), g["pk%d_agg_med" % i], rtol=1e-5, atol=1e-6, equal_nan=True)
        pl0, agg0 = peak_stimulation_3d(x, win_size=win, peak_filter=None)
        assert np.array_equal(pl0.cpu().numpy(), g["pk%d_peaks_none" % i]), i
        np.testing.assert_allclose(agg0.cpu().numpy(), g["pk%d_agg_none" % i], rtol=1e-5, atol=1e-6, equal_nan=True)
        only = peak_stimulation_3d(x, return_aggregation=False, win_size=win, peak_filter="median")
        assert np.array_equal(only.cpu().numpy(), g["pk%d_peaks_med" % i])


def test_peaks_vs_oracle_random_edge_cases(b2, torch_):
    from b200seg.peak_stimulation_3d import peaks_forward, peak_stimulation_3d
    rng = np.random.default_rng(11)
    for t in range(24):
        B, A = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        S, H, W = [int(v) for v in rng.integers(1, 40, 3)]
        x = rng.normal(size=(B, A, S, H, W)).astype(np.float32)
        if t % 4 == 1:
            x = np.round(x * 2) / 2                                    # plateaus / ties
        if t % 4 == 2:
            x[...] = 0.125                                             # constant map: only the first voxel... none
        if t % 6 == 3:
            x.flat[rng.integers(0, x.size, 3)] = -np.inf
        if t % 6 == 5:
            x.flat[rng.integers(0, x.size, 2)] = np.nan                # NaN: median becomes NaN -> no peaks
        if t % 5 == 0:
            x[x < 0] = -0.0                                            # signed zeros
        win = int(rng.choice([3, 3, 5, 7]))
        for mode, name in ((0, None), (1, "median")):
            p, agg, thr = peaks_forward(torch_.from_numpy(x).cuda(), win, mode)
            op, oagg, othr = oracle.peak_stimulation_3d(x, win_size=win, filter_mode=name)
            assert np.array_equal(p.cpu().numpy(), op), (t, mode, x.shape, win)
            np.testing.assert_allclose(agg.cpu().numpy(), oagg, rtol=1e-5, atol=1e-6, equal_nan=True)
            if mode == 1:
                assert np.array_equal(thr.cpu().numpy(), othr, equal_nan=True), (t, thr, othr)
        # arbitrary callable filter (mean), constant filter
        xt = torch_.from_numpy(x).cuda()
        if not np.isnan(x).any() and not np.isinf(x).any():
            mean = lambda inp: inp.view(inp.size(0), inp.size(1), -1).mean(2).view(inp.size(0), inp.size(1), 1, 1, 1)
            pl, _ = peak_stimulation_3d(xt, win_size=win, peak_filter=mean)
            thr_np = mean(xt).reshape(B, A).cpu().numpy()
            op = oracle.peak_stimulation_3d(x, win_size=win, filter_mode="given", thresholds=thr_np, return_aggregation=False)
            assert np.array_equal(pl.cpu().numpy(), op)


def test_peaks_aligned_rows_nan_signs_and_plateaus(b2, torch_):
    """Row widths that are multiples of 32 take the ballot-word store path of the window-3 scan; heights that are not
    multiples of the row group, both NaN signs with non-canonical payloads, -inf, signed zeros and plateaus."""
    from b200seg.peak_stimulation_3d import peaks_forward
    rng = np.random.default_rng(77)
    for t, (S, H, W) in enumerate([(5, 7, 32), (9, 13, 64), (3, 4, 96), (17, 6, 32), (1, 1, 32), (2, 33, 64), (12, 9, 128)]):
        x = rng.normal(size=(2, 2, S, H, W)).astype(np.float32)
        if t % 3 == 0:
            x = np.round(x * 2) / 2
        if t % 2 == 1:
            x[x < -1] = -0.0
            x.flat[rng.integers(0, x.size, 4)] = -np.inf
        xn = x.copy()
        if t in (1, 3, 6):                                             # NaNs of both signs, odd payloads
            u = xn.view(np.uint32)
            idx = rng.integers(0, x.size, 6)
            u.flat[idx[:2]] = 0xFFC00000; u.flat[idx[2:4]] = 0x7FC00001; u.flat[idx[4:]] = 0xFF800123
        for arr in (x, xn):
            for mode, name in ((0, None), (1, "median")):
                p, agg, thr = peaks_forward(torch_.from_numpy(arr).cuda(), 3, mode)
                op, oagg, othr = oracle.peak_stimulation_3d(arr, win_size=3, filter_mode=name)
                assert np.array_equal(p.cpu().numpy(), op), (t, mode, arr.shape)
                np.testing.assert_allclose(agg.cpu().numpy(), oagg, rtol=1e-5, atol=1e-6, equal_nan=True)
                if mode == 1:
                    assert np.array_equal(thr.cpu().numpy(), othr, equal_nan=True), (t, thr, othr)


def test_peaks_sampled_interval_median_ties_and_fallback(b2, torch_):
    """Maps of 2^16..2^21 elements take the sampled-interval median (one streaming pass); the threshold and the peak list
    stay exact for continuous data, heavy ties at the median (ReLU-like zeros, binary and few-level maps, constant maps),
    NaNs, -inf tails, and when the interval is forced to miss (slow exact selection) -- all three strategies agree."""
    from b200seg.peak_stimulation_3d import peaks_forward, set_median_mode, median_fallback_count
    rng = np.random.default_rng(404)
    shapes = [(1, 2, 16, 64, 64), (1, 3, 8, 96, 128), (2, 1, 32, 128, 128), (1, 1, 17, 61, 67), (1, 2, 16, 128, 256)]
    maps = []
    for t, shp in enumerate(shapes):
        x = rng.normal(size=shp).astype(np.float32)
        maps.append(("normal", x))
        maps.append(("relu", np.maximum(x - (0.05 if t % 2 else -0.05), 0).astype(np.float32)))     # ~52 % / ~48 % exact zeros
        maps.append(("binary", (x > 0.01).astype(np.float32)))
        maps.append(("levels", np.round(x * 1.5).astype(np.float32) * 0.25))
        c = np.full(shp, 0.375, np.float32); c.flat[rng.integers(0, c.size, 50)] = 1.0
        maps.append(("constant", c))
        n = x.copy(); n.flat[rng.integers(0, n.size, 3)] = np.nan
        maps.append(("nan", n))
        m = x.copy(); m.flat[rng.integers(0, m.size, m.size // 3)] = -np.inf; m[m > 1.0] = np.inf
        maps.append(("inf", m))
        maps.append(("tiny", (x * 1e-42).astype(np.float32)))                                      # denormals and signed zeros
    before = median_fallback_count()
    try:
        for name, x in maps:
            op, oagg, othr = oracle.peak_stimulation_3d(x, win_size=3, filter_mode="median")
            for mode in (0, 1, 2):
                set_median_mode(mode)
                p, agg, thr = peaks_forward(torch_.from_numpy(x).cuda(), 3, 1)
                assert np.array_equal(thr.cpu().numpy(), othr, equal_nan=True), (name, x.shape, mode, thr, othr)
                assert np.array_equal(p.cpu().numpy(), op), (name, x.shape, mode)
    finally:
        set_median_mode(0)
    # mode 2 forces one fallback per map without a NaN (a NaN decides the threshold before the selection); mode 0 must not
    # add any on these inputs
    forced = sum(int((~np.isnan(x).reshape(x.shape[0] * x.shape[1], -1).any(axis=1)).sum()) for _, x in maps)
    assert median_fallback_count() - before == forced, (median_fallback_count() - before, forced)


def test_peaks_full_resolution_map(b2, torch_):
    """(1,1,128,512,512) fp32 map (BASELINE config 3, full-resolution variant): exact peak list and threshold vs the oracle."""
    from b200seg import synth
    from b200seg.peak_stimulation_3d import peaks_forward
    x = synth.response_map(np.random.default_rng(1003), (128, 512, 512), n_peaks=200, channels=1)
    p, agg, thr = peaks_forward(torch_.from_numpy(x).cuda(), 3, 1)
    op, oagg, othr = oracle.peak_stimulation_3d(x, win_size=3, filter_mode="median")
    assert np.array_equal(thr.cpu().numpy(), othr)
    assert p.shape[0] == op.shape[0] and np.array_equal(p.cpu().numpy(), op)
    # aggregation = mean of the map over its peaks: an fp32 sum of 1.2 M terms is order dependent (torch reduces pairwise,
    # the oracle sequentially, the kernel per 8192-voxel chunk), so it is checked against the fp64 mean at 1e-5 relative
    exact = x[op[:, 0], op[:, 1], op[:, 2], op[:, 3], op[:, 4]].astype(np.float64).mean()
    np.testing.assert_allclose(agg.cpu().numpy().ravel()[0], exact, rtol=1e-5)
    np.testing.assert_allclose(oagg.ravel()[0], exact, rtol=5e-3)


def test_peaks_config3_size_and_backward(b2, torch_):
    """(1,14,32,128,128) fp32 response map (BASELINE config 3 @stride 4): exact peaks vs oracle; backward."""
    from b200seg import synth
    from b200seg.peak_stimulation_3d import peak_stimulation_3d
    rng = np.random.default_rng(1003)
    x = synth.response_map(rng, (32, 128, 128), n_peaks=60, channels=14)
    xt = torch_.from_numpy(x).cuda().requires_grad_(True)
    pl, agg = peak_stimulation_3d(xt, win_size=3, peak_filter="median")
    op, oagg, _ = oracle.peak_stimulation_3d(x, win_size=3, filter_mode="median")
    assert np.array_equal(pl.cpu().numpy(), op)
    np.testing.assert_allclose(agg.detach().cpu().numpy(), oagg, rtol=1e-5, atol=1e-6)
    w = torch_.arange(1, 15, dtype=torch_.float32, device="cuda").view(1, 14)
    (agg * w).sum().backward()
    gi = xt.grad.cpu().numpy()
    ref = np.zeros_like(x)
    ref[op[:, 0], op[:, 1], op[:, 2], op[:, 3], op[:, 4]] = (op[:, 1] + 1).astype(np.float32)    # peak_map * grad (:43-48)
    assert np.array_equal(gi, ref)


# ------------------------------------------------------------------------------------------ RoIAlign3D
def _ref_cuda_roialign(torch_, feat, rois, P, scale, sr, grad=None):
    """The reference's own CUDA kernels (oracle/_ref/libref_roialign3d.so), called through ctypes."""
    L = oracle.ref_roialign_lib()
    if L is None:
        return None
    f = torch_.from_numpy(feat).cuda(); r = torch_.from_numpy(rois).cuda()
    B, C, S, H, W = feat.shape
    R = rois.shape[0]
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    out = torch_.zeros((R, C, P, P, P), dtype=torch_.float32, device="cuda")
    L.ROIAlignForwardLaucher_3d.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    L.ROIAlignForwardLaucher_3d(vp(f.data_ptr()), cf(scale), R, S, H, W, C, P, P, P, sr, vp(r.data_ptr()), vp(out.data_ptr()),
                                vp(torch_.cuda.current_stream().cuda_stream))
    gi = None
    if grad is not None:
        g = torch_.from_numpy(grad).cuda()
        gi = torch_.zeros((B, C, S, H, W), dtype=torch_.float32, device="cuda")
        L.ROIAlignBackwardLaucher_3d.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
        L.ROIAlignBackwardLaucher_3d(vp(g.data_ptr()), cf(scale), B, R, S, H, W, C, P, P, P, sr, vp(r.data_ptr()), vp(gi.data_ptr()),
                                     vp(torch_.cuda.current_stream().cuda_stream))
    torch_.cuda.synchronize()
    return out.cpu().numpy(), (gi.cpu().numpy() if gi is not None else None)


def _close(a, b, rel=1e-5, floor="mean"):
    """fp32 tolerance of BASELINE.json: |a-b| <= 1e-5 * max(|b|, floor).
    floor="mean" (forward): mean |b|, i.e. elementwise-relative except where the result cancels to ~0.
    floor="max"  (backward): max |b| (norm-wise).  A grad_in voxel sums hundreds of signed taps from
    several RoIs; its fp32 value depends on the summation order at the 1e-5*|b| level -- the reference
    itself accumulates with atomicAdd in an unspecified order (roi_align_kernel_3d.cu:325-332)."""
    scale = max(float(np.abs(b).mean() if floor == "mean" else np.abs(b).max()), 1e-6)
    err = np.abs(a - b) / np.maximum(np.abs(b), scale)
    return bool(np.isfinite(a).all()) and float(err.max()) <= rel, float(err.max())


def _elementwise_rel(a, b, q=0.9999):
    """Strict element-wise relative error |a-b| / |b| over the elements with |b| >= 1e-3 * max|b| (no floor otherwise):
    returns (max, q-quantile).  Reported beside the floored figure of _close; the forward must meet 1e-5 at the maximum."""
    m = np.abs(b) >= 1e-3 * np.abs(b).max()
    e = np.abs(a - b)[m] / np.abs(b)[m]
    return float(e.max()), float(np.quantile(e, q))


@pytest.mark.parametrize("shape,R,scale,P,sr,side", [
    ((2, 16, 8, 32, 32), 64, 0.125, 7, 2, (10, 50)),       # nuclei-like (config 4 geometry)
    ((1, 8, 16, 40, 40), 40, 0.25, 7, 2, (10, 60)),        # soma tile
    ((2, 5, 6, 9, 11), 30, 0.25, 7, 2, (2, 60)),           # odd sizes, RoIs larger than the map
    ((1, 4, 8, 25, 25), 20, 0.125, 14, 2, (20, 120)),      # mask head P=14
    ((1, 3, 16, 40, 40), 12, 0.25, 7, 0, (10, 160)),       # adaptive sampling (sr=0), large footprints
    ((1, 2, 20, 64, 64), 6, 1.0, 7, 2, (30, 64)),          # footprint > 16 voxels: direct path
    ((1, 33, 4, 8, 8), 9, 0.5, 3, 1, (2, 12)),             # P=3, sr=1, channel tail
    ((1, 20, 8, 24, 24), 24, 0.125, 7, 3, (10, 50)),       # sr=3: count 27 is not a power of two -> generic body inside the fast kernel
    ((1, 6, 16, 48, 48), 10, 0.5, 7, 2, (20, 28)),         # footprints of 11..16 voxels with P=7: generic separable body inside the fast kernel
    ((2, 40, 8, 32, 32), 48, 0.125, 8, 2, (10, 50)),       # P=8 (no padding bin), channel groups of 16 + 8
])
def test_roialign_fwd_bwd_vs_oracle(b2, torch_, shape, R, scale, P, sr, side):
    from b200seg import synth
    from b200seg.roi_align_3d import RoIAlignFunction_3d
    feat, rois = synth.roialign_case(hash((shape, R)) % 1000, feat_shape=shape, n_rois=R, scale=scale, side=side, frac_outside=0.1)
    rois[0, 1:] = [-50, -50, -50, -40, -40, -40]                      # completely outside
    rois[1, 4:] = rois[1, 1:4] - 3                                    # malformed (end < start): forced to 1x1x1
    f = torch_.from_numpy(feat).cuda().requires_grad_(True)
    y = RoIAlignFunction_3d(P, P, P, scale, sr)(f, torch_.from_numpy(rois).cuda())
    ref = oracle.roialign3d_fwd(feat, rois, P, scale, sr)
    ok, err = _close(y.detach().cpu().numpy(), ref)
    assert ok, ("fwd", err)
    g = np.random.default_rng(1).standard_normal(ref.shape).astype(np.float32)
    y.backward(torch_.from_numpy(g).cuda())
    refg = oracle.roialign3d_bwd(g, rois, feat.shape, scale, sr)
    ok, err = _close(f.grad.cpu().numpy(), refg, floor="max")
    assert ok, ("bwd", err)
    # determinism of the atomics-free backward: bit-identical on a second run
    f2 = torch_.from_numpy(feat).cuda().requires_grad_(True)
    RoIAlignFunction_3d(P, P, P, scale, sr)(f2, torch_.from_numpy(rois).cuda()).backward(torch_.from_numpy(g).cuda())
    assert torch_.equal(f.grad, f2.grad)


def test_roialign_backward_paths_agree(b2, torch_):
    """The backward picks its path by the workspace it is handed: b200seg_roialign3d_bwd_workspace_bytes enables the
    two-launch path (per-RoI footprint gradients + gather), the smaller b200seg_roialign3d_workspace_bytes keeps the
    tile kernel.  Both are deterministic and agree with the oracle and with each other to rounding; the mix of
    qualifying and non-qualifying RoIs (large footprints) exercises the direct evaluation inside the gather pass."""
    from b200seg import synth, _lib
    feat, rois = synth.roialign_case(77, feat_shape=(2, 24, 10, 40, 40), n_rois=60, scale=0.25, side=(8, 70), frac_outside=0.1)
    rois[0, 1:] = [-50, -50, -50, -40, -40, -40]
    g = np.random.default_rng(3).standard_normal((60, 24, 7, 7, 7)).astype(np.float32)
    ref = oracle.roialign3d_bwd(g, rois, feat.shape, 0.25, 2)
    L = _lib.lib()
    gd, rd = torch_.from_numpy(g).cuda(), torch_.from_numpy(rois).cuda()
    B, C, S, H, W = feat.shape
    outs = []
    for nbytes in (L.b200seg_roialign3d_workspace_bytes(60, S, H, W, 7), L.b200seg_roialign3d_bwd_workspace_bytes(60, C, S, H, W, 7)):
        ws = torch_.empty(nbytes, dtype=torch_.uint8, device="cuda")
        res = []
        for _ in range(2):
            gi = torch_.full(feat.shape, float("nan"), dtype=torch_.float32, device="cuda")
            _lib.check(L.b200seg_roialign3d_bwd_dev(_lib.ptr(gd), 0, _lib.ptr(rd), _lib.ptr(gi), B, C, S, H, W, 60, 7, 7, 7, 0.25, 2, 0,
                                                    _lib.ptr(ws), nbytes, _lib.current_stream()), "bwd")
            res.append(gi)
        assert torch_.equal(res[0], res[1])
        ok, err = _close(res[0].cpu().numpy(), ref, floor="max")
        assert ok, err
        outs.append(res[0])
    assert L.b200seg_roialign3d_bwd_workspace_bytes(60, C, S, H, W, 7) > L.b200seg_roialign3d_workspace_bytes(60, S, H, W, 7)
    ok, err = _close(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rel=2e-6, floor="max")
    assert ok, err


def test_roialign_vs_reference_cuda_kernels(b2, torch_):
    """Same inputs through the reference's own .cu (compiled unmodified for sm_100a)."""
    from b200seg import synth
    from b200seg.roi_align_3d import roialign3d_forward, roialign3d_backward
    feat, rois = synth.roialign_case(1004, feat_shape=(2, 32, 8, 32, 32), n_rois=128, scale=0.125)
    g = np.random.default_rng(2).standard_normal((128, 32, 7, 7, 7)).astype(np.float32)
    got = _ref_cuda_roialign(torch_, feat, rois, 7, 0.125, 2, grad=g)
    if got is None:
        pytest.skip("oracle/_ref/libref_roialign3d.so not built")
    ref_out, ref_gi = got
    r = torch_.from_numpy(rois).cuda()
    y = roialign3d_forward(torch_.from_numpy(feat).cuda(), r, 7, 7, 7, 0.125, 2)
    ok, err = _close(y.cpu().numpy(), ref_out)
    assert ok, err
    gi = roialign3d_backward(torch_.from_numpy(g).cuda(), r, feat.shape, 0.125, 2)
    ok, err = _close(gi.cpu().numpy(), ref_gi, floor="max")      # the reference sums with atomics (order varies)
    assert ok, err
    # and the oracle agrees with the reference kernels too (pins the RoIAlign restatement)
    ok, err = _close(oracle.roialign3d_fwd(feat, rois, 7, 0.125, 2), ref_out)
    assert ok, err


def test_roialign_layout_shw_is_adjoint_and_bf16(b2, torch_):
    from b200seg import synth
    from b200seg.roi_align_3d import RoIAlignFunction_3d, roialign3d_forward
    feat, rois = synth.roialign_case(5, feat_shape=(1, 6, 6, 12, 12), n_rois=10, scale=0.25, side=(8, 30), frac_outside=0.0)
    f = torch_.from_numpy(feat).cuda()
    r = torch_.from_numpy(rois).cuda()
    ref = roialign3d_forward(f, r, 7, 7, 7, 0.25, 2, layout=0)
    shw = roialign3d_forward(f, r, 7, 7, 7, 0.25, 2, layout=1)
    # (H,W,S) storage vs (S,H,W): the same bins; the two layouts contract the axes in different orders (the last pass is the
    # one whose bin index is the slowest output index), so they agree to rounding, not bit for bit
    ok, err = _close(shw.cpu().numpy(), ref.permute(0, 1, 4, 2, 3).contiguous().cpu().numpy(), rel=2e-6)
    assert ok, err
    # <A x, g> == <x, A^T g> for the self-consistent layout
    fx = f.clone().double().float().requires_grad_(True)
    y = RoIAlignFunction_3d(7, 7, 7, 0.25, 2, layout="shw")(fx, r)
    g = torch_.randn_like(y)
    y.backward(g)
    lhs = float((y.detach().double() * g.double()).sum()); rhs = float((fx.grad.double() * f.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    # bf16 features: fp32 accumulation, result rounded once to bf16
    yb = roialign3d_forward(f.bfloat16(), r, 7, 7, 7, 0.25, 2)
    want = roialign3d_forward(f.bfloat16().float(), r, 7, 7, 7, 0.25, 2).bfloat16()
    assert yb.dtype == torch_.bfloat16 and torch_.equal(yb, want)


def test_roialign_config4_full_size(b2, torch_):
    """BASELINE config 4: 512 RoIs x 256 ch x 7^3 on (2,256,8,32,32): sampled RoIs vs the oracle, full
    tensor vs the reference CUDA kernel when available, backward linearity."""
    from b200seg import synth
    from b200seg.roi_align_3d import roialign3d_forward, roialign3d_backward
    feat, rois = synth.roialign_case(1004)
    f, r = torch_.from_numpy(feat).cuda(), torch_.from_numpy(rois).cuda()
    y = roialign3d_forward(f, r, 7, 7, 7, 0.125, 2)
    sel = np.arange(0, 512, 37)
    ok, err = _close(y[torch_.from_numpy(sel).cuda()].cpu().numpy(), oracle.roialign3d_fwd(feat, rois[sel], 7, 0.125, 2))
    assert ok, err
    got = _ref_cuda_roialign(torch_, feat, rois, 7, 0.125, 2)
    if got is not None:
        ok, err = _close(y.cpu().numpy(), got[0])
        assert ok, err
    g1, g2 = torch_.randn_like(y), torch_.randn_like(y)
    a = roialign3d_backward(g1, r, feat.shape, 0.125, 2); b = roialign3d_backward(g2, r, feat.shape, 0.125, 2)
    c = roialign3d_backward(g1 + g2, r, feat.shape, 0.125, 2)
    assert float((a + b - c).abs().max()) <= 1e-4 * float(c.abs().max())


# ------------------------------------------------------------------------------------------ whole-volume prefilters
def test_prefilter_golden_and_oracle(b2, golden, torch_):
    from b200seg import prefilter
    d = golden("prefilter.npz")
    for n in ("u16", "u8", "thin"):
        g = prefilter.gaussian_filter(d[n + "_img"], 1)
        assert g.dtype == d[n + "_img"].dtype and np.array_equal(g, d[n + "_gauss"]), n                # bit exact vs scipy
        assert np.array_equal(prefilter.median_filter(g, 3), d[n + "_median"]), n
        assert np.array_equal(prefilter.prefilter_nuclei(d[n + "_img"]), d[n + "_median"]), n
        if n != "thin":
            assert np.array_equal(prefilter.gaussian_filter(d[n + "_img"], 2), d[n + "_gauss_s2"]), n   # radius 8
    # seeded volumes: odd widths (scalar paths), dims smaller than the radius, tiles that straddle the border, all radii
    rng = np.random.default_rng(91)
    for shape, dtype, hi in [((5, 33, 67), np.uint16, 65535), ((13, 70, 129), np.uint8, 255), ((40, 31, 64), np.uint16, 3000),
                             ((2, 3, 5), np.uint16, 65535), ((1, 1, 1), np.uint8, 255), ((9, 64, 130), np.uint16, 65535)]:
        img = rng.integers(0, hi + 1, shape).astype(dtype)
        img[rng.random(shape) < 0.2] = hi                                   # saturated voxels: results must not overflow
        for sigma in (0.3, 0.5, 0.7, 1, 1.2, 1.5, 1.7, 2):
            assert np.array_equal(prefilter.gaussian_filter(img, sigma), oracle.gaussian_filter(img, sigma)), (shape, sigma)
        assert np.array_equal(prefilter.median_filter(img), oracle.median_filter3(img)), shape
    const = np.full((7, 40, 70), 1000, np.uint16)
    assert np.array_equal(prefilter.gaussian_filter(const, 1), oracle.gaussian_filter(const, 1))
    # BASELINE config 2 shape (59 x 350 x 640 uint16): oracle on a z-slab (the filters are local), properties on the rest
    S, H, W = 59, 350, 640
    vol = rng.integers(0, 4096, (S, H, W)).astype(np.uint16)
    vol[:, 100:200, 300:500] += 20000
    t = torch_.from_numpy(vol).cuda()
    g = prefilter.gaussian_filter(t, 1)
    m = prefilter.median_filter(g, 3)
    gh, mh = g.cpu().numpy(), m.cpu().numpy()
    ref_g = oracle.gaussian_filter(vol[:24], 1)                          # exact for z < 24 - 4
    assert np.array_equal(gh[:20], ref_g[:20])
    assert np.array_equal(mh[:19], oracle.median_filter3(gh[:20])[:19])
    assert gh.min() >= vol.min() and gh.max() <= vol.max() and mh.min() >= gh.min() and mh.max() <= gh.max()
    flipped = prefilter.gaussian_filter(torch_.from_numpy(np.ascontiguousarray(vol[::-1, ::-1, ::-1])).cuda(), 1).cpu().numpy()
    assert np.array_equal(flipped[::-1, ::-1, ::-1], gh)                # the symmetric kernel commutes with mirroring
    assert np.array_equal(prefilter.median_filter(m, 3).cpu().numpy()[30], oracle.median_filter3(mh[28:33])[2])


def test_zscore_and_prm_to_uint8(b2, golden, torch_):
    from b200seg import prefilter
    d = golden("prefilter.npz")
    z, mu, sd = prefilter.zscore_norm(d["u16_img"], return_stats=True)
    assert mu == d["zs_stats"][0]                                          # exact integer sums: the mean is bit exact
    assert abs(sd - d["zs_stats"][1]) <= 1e-12 * d["zs_stats"][1]
    np.testing.assert_allclose(z, d["zs_out"].astype(np.float32), rtol=1e-6, atol=1e-6)
    assert np.array_equal(prefilter.prm_to_uint8(d["prm_in"]), d["prm_u8"])    # fp32 ops in the reference's order: bit exact
    rng = np.random.default_rng(17)
    for dtype, hi in ((np.uint8, 255), (np.uint16, 65535)):
        im = rng.integers(0, hi + 1, (31, 50, 77)).astype(dtype)
        im[rng.random(im.shape) < 0.4] = 0
        ref, rm, rs = oracle.zscore_norm(im)
        z, mu, sd = prefilter.zscore_norm(im, return_stats=True)
        assert mu == rm and abs(sd - rs) <= 1e-12 * rs
        np.testing.assert_allclose(z, ref.astype(np.float32), rtol=1e-6, atol=1e-6)
    imf = (rng.random((20, 33, 47)).astype(np.float32) * 5000).astype(np.float32)
    imf[rng.random(imf.shape) < 0.3] = 0
    mask = imf > 0                                                         # lib/utils/blob.py:180-184 (float32 image)
    ref = (imf - np.mean(imf[mask])) / np.std(imf[mask])
    z, mu, sd = prefilter.zscore_norm(imf, return_stats=True)
    assert abs(mu - np.mean(imf[mask].astype(np.float64))) <= 1e-12 * abs(mu)
    np.testing.assert_allclose(z, ref, rtol=2e-5, atol=2e-6)               # the reference sums in float32 here
    z2 = prefilter.zscore_norm(imf)
    assert np.array_equal(z, z2)                                            # deterministic run to run
    prm = rng.standard_normal((14, 9, 30, 31)).astype(np.float32) * np.linspace(1e-3, 50, 14, dtype=np.float32)[:, None, None, None]
    assert np.array_equal(prefilter.prm_to_uint8(prm), oracle.prm_to_uint8(prm))
    big = torch_.rand((2, 64, 256, 256), device="cuda")
    u8 = prefilter.prm_to_uint8(big)
    assert np.array_equal(u8.cpu().numpy(), oracle.prm_to_uint8(big.cpu().numpy()))


# ------------------------------------------------------------------------------------------ RPN proposal generation
def _proposals_close(got, ref):
    # boxes: float32 arithmetic in the reference's order; exp() may differ from numpy's float32 exp by one ulp
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=2e-5)


def test_generate_proposals_golden_and_oracle(b2, golden, torch_):
    from b200seg.generate_proposals_3d import GenerateProposalsOp_3d
    d = golden("proposals.npz")
    for name in "abc":
        stride, pre, post, thr = d[name + "_cfg"]
        op = GenerateProposalsOp_3d(d[name + "_anchors"], 1.0 / stride, pre_nms_topN=int(pre), post_nms_topN=int(post), nms_thresh=float(thr))
        rois, probs, keep_idx = op(torch_.from_numpy(d[name + "_scores"]).cuda(), torch_.from_numpy(d[name + "_deltas"]).cuda(),
                                   torch_.from_numpy(d[name + "_im_info"]))
        assert rois.shape == d[name + "_rois"].shape, name
        assert np.array_equal(rois[:, 0], d[name + "_rois"][:, 0])
        _proposals_close(rois, d[name + "_rois"])
        assert np.array_equal(probs, d[name + "_probs"]), name
        assert np.array_equal(keep_idx, d[name + "_keep_idx_last"]), name
    # soma test-tile geometry (14 anchors on 16x40x40, stride 4): every branch against the oracle
    rng = np.random.default_rng(33)
    A, S, H, W, stride = 14, 16, 40, 40, 4
    anchors = np.concatenate([np.stack([-(s / 2 - 2) * np.ones(3), (s / 2 + 1) * np.ones(3)]).reshape(1, 6) * np.array([1, 1, r, 1, 1, r])
                              for s in (8, 12, 16, 20, 24, 30, 36) for r in (1.0, 0.5)]).astype(np.float32)
    n = A * S * H * W
    base = rng.permutation(n).astype(np.float32).reshape(1, A, S, H, W) / np.float32(n)
    deltas = (rng.standard_normal((1, 6 * A, S, H, W)) * 0.3).astype(np.float32)
    im_info = np.array([[S * stride, H * stride - 7, W * stride, 1.0]], np.float32)
    quant = np.round(base * 50) / 50                                           # heavy ties: 51 distinct score values
    cases = [(base, 1000, 300, 0.23, 0), (base, 2000, 1000, 0.7, 0), (base, 0, 50, 0.5, 0), (base, 500, 0, 0.4, 0), (base, 600, 100, 0.0, 0),
             (base, 800, 200, 0.3, 14.0), (quant.astype(np.float32), 1000, 300, 0.5, 0), (base, n + 5, 10, 0.9, 0)]
    for scores, pre, post, thr, min_size in cases:
        if pre <= 0 or pre >= n:                                               # take-all: keep the quadratic NMS small
            scores = scores.copy(); scores.reshape(-1)[rng.random(n) < 0.99] *= 0.0
            scores = scores[:, :2, :4, :10, :10].copy(); dl = deltas.reshape(1, A, 6, S, H, W)[:, :2, :, :4, :10, :10].reshape(1, 12, 4, 10, 10).copy()
            an = anchors[:2]
        else:
            dl, an = deltas, anchors
        op = GenerateProposalsOp_3d(an, 1.0 / stride, pre_nms_topN=pre, post_nms_topN=post, nms_thresh=thr, min_size=min_size)
        rois, probs, keep_idx = op(torch_.from_numpy(scores).cuda(), torch_.from_numpy(dl).cuda(), im_info)
        p, s, k = oracle.generate_proposals(scores[0], dl[0], im_info[0], an, stride, pre, post, thr, min_size)
        assert np.array_equal(keep_idx, k), (pre, post, thr, min_size, len(keep_idx), len(k))
        assert np.array_equal(probs[:, 0], s)
        _proposals_close(rois[:, 1:], p)
        assert np.all(rois[:, 0] == 0)


def test_box_results_golden(b2, golden):
    from b200seg.box_results import box_results_with_nms_and_limit
    from test_oracle_golden import _box_results_check
    d = golden("box_results.npz")

    def fn(scores, boxes, idx, ncls, thr, nms, per_im):
        return box_results_with_nms_and_limit(scores, boxes, idx, num_classes=ncls, rpn_only=False, score_thresh=thr, nms=nms,
                                              detections_per_im=per_im)
    for name in "abc":
        _box_results_check(fn, d, name)
    s, b, cb, ci = fn(np.zeros((5, 2), np.float32), np.zeros((5, 12), np.float32), None, 2, 0.05, 0.3, 100)     # nothing above the threshold
    assert s.shape == (0,) and b.shape == (0, 6) and cb[1].shape == (0, 7)


# ------------------------------------------------------------------------------------------ Mask R-CNN mask paste-back
def test_segm_results_golden_and_oracle(b2, golden, torch_):
    """segm_results (core/test.py:886-945): the numpy seam against the fixture produced by the reference's own function, the
    device form (packed crops + expanded volumes) against the oracle on seeded boxes that upsample, downsample (anti-aliased
    axes), touch every border of the volume, or miss it."""
    from b200seg import segm
    from test_oracle_golden import _segm_case
    g = golden("segm.npz")
    for name in ("a", "b"):
        c = _segm_case(g, name)
        out = segm.segm_results(c["cls_boxes"], c["masks"], c["boxes"], *c["shape"], num_classes=c["ncls"],
                                cls_specific_mask=c["cls_specific"], resolution=c["M"])
        assert [len(out[j]) for j in range(c["ncls"])] == c["counts"]
        vols = np.stack([v for j in range(1, c["ncls"]) for v in out[j]])
        assert vols.dtype == np.uint8 and np.array_equal(vols, c["vols"]), name
    rng = np.random.default_rng(4242)
    for M, (S, H, W), n in ((14, (32, 48, 64), 40), (7, (16, 16, 32), 12), (26, (8, 24, 16), 5)):
        ncls = 3
        counts = [0, n - n // 3, n // 3]
        lo = np.stack([rng.uniform(-8, W - 4, n), rng.uniform(-8, H - 4, n), rng.uniform(-6, S - 3, n)], axis=1)
        ext = np.stack([rng.uniform(0, 1.2 * W, n), rng.uniform(0, 1.2 * H, n), rng.uniform(0, 1.5 * S, n)], axis=1)
        ext[::5] = rng.uniform(0, 6, (len(ext[::5]), 3))                  # small boxes: all three axes anti-aliased
        boxes = np.concatenate([lo, lo + ext], axis=1).astype(np.float32)
        boxes[1] = (-40, -30, -20, -12, -9, -7)                            # misses the volume: all-zero output
        masks = rng.random((n, ncls, M, M, M), dtype=np.float32)
        masks[2] = 0.25                                                   # constant block below the threshold
        masks[3] = 0.75
        cls_boxes = [[]] + [np.zeros((c, 7), np.float32) for c in counts[1:]]
        ref = oracle.segm_results(cls_boxes, masks, boxes, S, H, W, num_classes=ncls, cls_specific_mask=True, thresh_binarize=0.5)
        ref = np.stack([v for j in range(1, ncls) for v in ref[j]])
        dev = segm.segm_results_device(cls_boxes, torch_.from_numpy(masks).cuda(), boxes, S, H, W, expand=True, num_classes=ncls,
                                       cls_specific_mask=True, thresh_binarize=0.5)
        torch_.cuda.synchronize()
        assert np.array_equal(dev["volumes"].cpu().numpy(), ref), (M, n)
        crops, off, clip = dev["crops"].cpu().numpy(), dev["crop_off"], dev["boxes"]
        for d in range(n):
            x0, y0, z0, x1, y1, z1 = clip[d]
            if off[d + 1] > off[d]:
                assert np.array_equal(crops[off[d]:off[d + 1]].reshape(z1 - z0, y1 - y0, x1 - x0), ref[d, z0:z1, y0:y1, x0:x1]), d
        assert off[2] == off[1] and ref[2].sum() == 0 and ref[3].sum() > 0
        host = segm.segm_results(cls_boxes, masks, boxes, S, H, W, num_classes=ncls, cls_specific_mask=True, thresh_binarize=0.5)
        assert np.array_equal(np.stack([v for j in range(1, ncls) for v in host[j]]), ref)
    empty = segm.segm_results([[], np.zeros((0, 7), np.float32)], np.zeros((0, 2, 14, 14, 14), np.float32), np.zeros((0, 6), np.float32),
                              8, 8, 8, num_classes=2)
    assert empty == [[], []]


# ------------------------------------------------------------------------------------------ nuclei per-instance chain
def test_binarize_nuclei_golden_and_oracle(b2, golden, torch_):
    """binarization_nuclei.py:92-149: each fixture crop (reference lines executed from the file) as a one-instance volume;
    then whole volumes (uint8 with stretch, uint16 without) with overlapping boxes, hollow blobs, constant-PRM and
    out-of-volume rejects against the oracle chain: label volume, final masks, status and survivors bit-exact."""
    from b200seg import binarization_nuclei as bn, synth
    from b200seg.binarization import crop_offsets
    from test_oracle_golden import _nuclei_crops
    for k, c in enumerate(_nuclei_crops(golden("nuclei.npz"))):
        S, H, W = c["img"].shape
        r = bn.binarize_nuclei_host(c["img"], np.array([[0, 0, 0, W - 1, H - 1, S - 1]], np.int32), [c["prm"]], want_masks=True)
        assert r["status"].tolist() == [0] and r["b_max"].tolist() == [c["b"]] and r["survive"].tolist() == [True], k
        assert np.array_equal(r["masks"].reshape(c["img"].shape) > 0, c["mask"]) and np.array_equal(r["seg"] > 0, c["mask"]), k
    for seed, shape, nb, as16 in ((139, (32, 128, 160), 8, False), (140, (24, 120, 150), 6, True), (141, (40, 160, 192), 14, False)):
        case = synth.postproc_case(seed, shape=shape, n_blobs=nb, n_dup=4, n_false=3, sigma_xy=(9, 13), sigma_z=(3, 6))
        vol = case["volume"].astype(np.int32)
        zz, yy, xx = np.ogrid[:shape[0], :shape[1], :shape[2]]
        for bl in case["blobs"][::2]:                                     # cavities for the hole filling
            cz, cy, cx = bl["c"]
            r2 = ((zz - cz) / (0.45 * bl["sz"])) ** 2 + ((yy - cy) / (0.45 * bl["sxy"])) ** 2 + ((xx - cx) / (0.45 * bl["sxy"])) ** 2
            vol -= (0.9 * bl["amp"] * np.exp(-0.5 * r2)).astype(np.int32)
        vol = np.clip(vol, 0, 255).astype(np.uint8)
        if as16:
            vol = vol.astype(np.uint16) * 7 + 11
        # selection (:72-87): edge filter, NMS by volume on the GPU, score > 0.4 -- against the same steps with the oracle NMS
        dets = case["dets"]
        idx = np.nonzero(bn.nuclei_edge_filter(dets, shape[2]))[0]
        kept = idx[oracle.nms_3d_volume(np.ascontiguousarray(dets[idx]), 0.15)]
        assert np.array_equal(bn.nuclei_select(dets, shape[2], nms_thresh=0.15, score_thresh=0.4), kept[dets[kept, -1] > 0.4])
        sel = np.arange(len(dets))            # the chain itself is exercised on EVERY box: duplicates overlap, false boxes have a constant PRM
        boxes = bn.nuclei_boxes(case["dets"][sel], np.zeros((len(sel), 3), np.int64), max(shape[1], shape[2]), shape[0])
        boxes[:, 3] = np.minimum(boxes[:, 3], shape[2] - 1); boxes[:, 4] = np.minimum(boxes[:, 4], shape[1] - 1)
        off = crop_offsets(boxes)
        crops = []
        for i, j in enumerate(sel):
            ob = case["boxes"][j]
            full = case["prm"][case["crop_off"][j]:case["crop_off"][j + 1]].reshape(ob[5] - ob[2] + 1, ob[4] - ob[1] + 1, ob[3] - ob[0] + 1)
            b = boxes[i]
            crops.append(np.ascontiguousarray(full[b[2] - ob[2]:b[5] - ob[2] + 1, b[1] - ob[1]:b[4] - ob[1] + 1, b[0] - ob[0]:b[3] - ob[0] + 1]))
        seg, status, survive, masks = oracle.binarize_nuclei(vol, boxes, crops)
        r = bn.binarize_nuclei_host(vol, boxes, crops, want_masks=True)
        assert r["status"].tolist() == status, (seed, r["status"].tolist(), status)
        assert np.array_equal(r["seg"], seg), seed
        assert r["survive"].tolist() == survive, seed
        for i in range(len(sel)):
            assert np.array_equal(r["masks"][off[i]:off[i + 1]].reshape(masks[i].shape) > 0, masks[i]), (seed, i)
        assert 0 in status and 7 not in status and len(np.unique(seg)) > 4      # constant PRM crops follow the script's NaN -> 0 path
        d = bn.binarize_nuclei(torch_.from_numpy(vol).cuda(),
                               torch_.from_numpy(boxes).cuda(), torch_.from_numpy(np.concatenate([c.ravel() for c in crops])).cuda(),
                               torch_.from_numpy(off).cuda())
        torch_.cuda.synchronize()
        assert np.array_equal(d[0].cpu().view(torch_.int16).numpy().view(np.uint16), seg) and d[3].cpu().tolist() == status
    bad = np.array([[0, 0, 0, 200, 5, 5]], np.int32)                          # box outside the volume: status 2, nothing pasted
    r = bn.binarize_nuclei_host(np.zeros((8, 16, 16), np.uint8), bad, [np.ones((6, 6, 201), np.uint8)])
    assert r["status"].tolist() == [2] and not r["seg"].any() and r["survive"].tolist() == [False]
    r = bn.binarize_nuclei_host(np.zeros((8, 16, 16), np.uint8), np.zeros((0, 6), np.int32), [])
    assert not r["seg"].any() and len(r["status"]) == 0


def test_config2_sizes_segm_and_nuclei(b2, torch_):
    """BASELINE configs[1] sizes: 100 detections with 14^3 masks on a 64x200x200 tile (segm_results) and the nuclei chain
    on a 59x350x640 uint16 volume with 50 large instances -- both against the oracle, bit-exact."""
    from b200seg import segm, synth, binarization_nuclei as bn
    rng = np.random.default_rng(2002)
    n, M, (S, H, W) = 100, 14, (64, 200, 200)
    lo = np.stack([rng.uniform(-5, W - 30, n), rng.uniform(-5, H - 30, n), rng.uniform(-5, S - 15, n)], axis=1)
    ext = np.stack([rng.uniform(25, 60, n), rng.uniform(25, 60, n), rng.uniform(10, 40, n)], axis=1)
    boxes = np.concatenate([lo, lo + ext], axis=1).astype(np.float32)
    masks = rng.random((n, 2, M, M, M), dtype=np.float32)
    cls_boxes = [[], np.zeros((n, 7), np.float32)]
    ref = np.stack(oracle.segm_results(cls_boxes, masks, boxes, S, H, W, num_classes=2)[1])
    dev = segm.segm_results_device(cls_boxes, torch_.from_numpy(masks).cuda(), boxes, S, H, W, expand=True, num_classes=2)
    torch_.cuda.synchronize()
    assert ref.sum() > 1e6 and np.array_equal(dev["volumes"].cpu().numpy(), ref)
    shape = (59, 350, 640)
    case = synth.postproc_case(1002, shape=shape, n_blobs=40, n_dup=10, n_false=0, sigma_xy=(10, 15), sigma_z=(4, 7))
    vol = case["volume"].astype(np.uint16) * 7 + 11
    off = case["crop_off"]
    crops = [case["prm"][off[i]:off[i + 1]] for i in range(len(case["boxes"]))]
    seg, status, survive, _ = oracle.binarize_nuclei(vol, case["boxes"], crops)
    r = bn.binarize_nuclei_host(vol, case["boxes"], case["prm"])
    assert r["status"].tolist() == status and r["survive"].tolist() == survive and np.array_equal(r["seg"], seg)
    assert status.count(0) >= 45 and len(np.unique(seg)) > 40


# ------------------------------------------------------------------------------------------ evaluation records
def test_eval_records_golden_and_oracle(b2, golden, torch_):
    """Label presence, overlap-matrix matching and voxel counts behind eval_volume_soma / eval_volume_nuclei: the fixture
    produced by the reference's evaluation code, then a 64x256x256 volume pair (BASELINE configs[0] shape) against the oracle."""
    from b200seg import evaluation as ev, synth
    from test_oracle_golden import _eval_images
    g = golden("eval.npz")
    score, match, n_pos = [], [], 0
    for im in _eval_images(g):
        assert np.array_equal(ev.label_ids(im["pred"]), np.unique(im["pred"])[1:])
        r = ev.eval_volume_soma(im["pred"], im["gt"], im["score"], 0.3)
        score += r["score"]; match += r["match"]; n_pos += r["n_pos"]
        n = ev.eval_volume_nuclei(im["pred"], im["gt"], im["det_boxes"], im["gt_boxes"], 0.4)
        assert np.array_equal(n["tp"], im["tp"]) and np.array_equal(n["fp"], im["fp"])
        assert [n["tp_pixel"], n["gt_pixel"], n["pre_pixel"]] == im["pixels"].tolist()
    prec, rec = ev.precision_recall(score, match, n_pos)
    assert np.array_equal(prec, g["prec"]) and np.array_equal(rec, g["rec"]) and ev.voc_ap(rec, prec) == float(g["ap"])
    # the chain's own label volume as the prediction, a shifted copy as ground truth
    case = synth.postproc_case(1001, shape=(64, 256, 256), n_blobs=35, n_dup=10, n_false=5)
    out = b2.postproc_soma_host(case["volume"], case["dets"], case["boxes"], case["prm"], case["crop_off"], 0.23)
    pred = out["seg"]
    ps = out["scores"].astype(np.float64)
    drop = int(ps[3, 0])
    gt = np.roll(pred, (1, 2, -2), axis=(0, 1, 2)).copy()
    gt[gt == drop] = 0                                                     # one ground truth instance removed
    assert np.array_equal(ev.label_ids(gt), np.unique(gt)[1:])
    r = ev.eval_volume_soma(pred, gt, ps, 0.3)
    s0, m0, n0 = oracle.eval_volume_soma(pred, gt, ps, 0.3)
    assert r["score"] == s0 and r["match"] == m0 and r["n_pos"] == n0 and sum(m0) > 20 and 0 in m0
    ids = ps[:, 0].astype(int)
    det_boxes = case["boxes"][out["rank_order"][ids - 1]].astype(np.float32)
    gt_boxes = det_boxes + np.array([-2, 2, 1, -2, 2, 1], np.float32)
    gt_boxes = gt_boxes[ids != drop]
    n = ev.eval_volume_nuclei(pred, gt, det_boxes, gt_boxes, 0.4)
    tp, fp, tpp, gtp, prp = oracle.eval_volume_nuclei(pred, gt, det_boxes, gt_boxes, 0.4)
    assert np.array_equal(n["tp"], tp) and np.array_equal(n["fp"], fp) and [n["tp_pixel"], n["gt_pixel"], n["pre_pixel"]] == [tpp, gtp, prp]
    assert tp.sum() > 20 and fp.sum() >= 1 and 0 < tpp < prp


def test_nuclei_script_golden_cuda(b2, golden):
    """binarization_nuclei.py:73-148 executed from the reference file (fixture) against the CUDA path end to end:
    selection with the GPU NMS by volume, clamped boxes, the per-instance chain, label volume and survivor table."""
    from b200seg import binarization_nuclei as bn
    from test_oracle_golden import _nuclei_script_inputs
    g = golden("nuclei_script.npz")
    sel, boxes, crops = _nuclei_script_inputs(g, b2.nms_3d_volume)
    assert np.array_equal(g["dets"][sel], g["visited_dets"])
    r = bn.binarize_nuclei_host(g["img"], boxes, crops)
    assert r["status"].tolist() == [0] * len(sel) and np.array_equal(r["seg"], g["seg"])
    assert np.array_equal(bn.id_det_rows(boxes, g["dets"][sel, -1], r["survive"]), g["id_det"])


def test_soma_script_golden_cuda(b2, golden):
    """binarization_soma.py:57-104 executed from the reference file (fixture) against the CUDA chain through the host entry:
    visit order, label volume and the [[id, score]] table."""
    g = golden("soma_script.npz")
    out = b2.postproc_soma_host(g["img"], g["dets"], g["boxes"], g["prm"], g["crop_off"], 0.23)
    assert np.array_equal(g["dets"][out["rank_order"]], g["visited_dets"])
    assert np.array_equal(out["seg"], g["seg"])
    assert np.array_equal(out["scores"].astype(np.float64), g["scores"])


# ------------------------------------------------------------------------------------------ round 2: parity holes of round 1
def test_bench_volume_800_detections_vs_oracle(b2):
    """One volume of the EXACT bench workload (bench.py: seed 2000, 128x512x512, 200 blobs + 400 duplicates + 200 false boxes,
    sigma up to 11) through the host entry point against the oracle chain: visit order, statuses, thresholds, label volume."""
    from b200seg import synth
    c = synth.postproc_case(2000, shape=(128, 512, 512), n_blobs=200, n_dup=400, n_false=200, sigma_xy=(5, 11.0), sigma_z=(3, 6))
    out = b2.postproc_soma_host(c["volume"], c["dets"], c["boxes"], c["prm"], c["crop_off"], 0.23)
    o = oracle_chain(c, 0.23)
    assert np.array_equal(out["rank_order"], o["order"]) and len(o["order"]) > 300
    assert np.array_equal(out["survive"], o["survive"])
    for i, st in o["status"].items():
        assert out["status"][i] == st, (i, out["status"][i], st)
    for i, bm in o["b_max"].items():
        assert out["b_max"][i] == bm, i
    assert np.array_equal(out["seg"], o["seg"])


def test_otsu_stress_10k_crops_vs_oracle(b2, torch_):
    """>= 10^4 seeded crops, GPU against the oracle: b_max and mask bit for bit (near-ties of the fp64 criterion are the risk:
    the kernel evaluates it from exact integer prefix sums, the reference accumulates incrementally, otsu.py:251-274)."""
    rng = np.random.default_rng(31415)
    n_total, n_bad = 0, 0
    for batch in range(10):
        imgs, prms = [], []
        for i in range(1024):
            shape = tuple(int(v) for v in rng.integers(3, 14, 3))
            kind = i % 4
            if kind == 0:                                   # soma-normalised blob crop (levels 30..330)
                zz, yy, xx = np.meshgrid(*[np.arange(s_) for s_ in shape], indexing="ij")
                c = [s_ / 2 + rng.uniform(-1, 1) for s_ in shape]
                r2 = sum(((g_ - c_) / (s_ / 3 + 0.5)) ** 2 for g_, c_, s_ in zip((zz, yy, xx), c, shape))
                v = np.clip(rng.uniform(0, 40, shape) + rng.uniform(80, 200) * np.exp(-0.5 * r2), 0, 255).astype(np.uint8)
                p = (255 * np.exp(-0.5 * r2 / 0.64)).astype(np.uint8)
                if p.max() == 0 or v.max() == 0:
                    p[tuple(s_ // 2 for s_ in shape)] = 200; v[tuple(s_ // 2 for s_ in shape)] = 100
                a, b = oracle.soma_normalise(v, p)
            elif kind == 1:                                 # few gray levels: many exact ties in the criterion
                a = rng.integers(0, int(rng.integers(2, 6)), shape).astype(np.uint16) * int(rng.integers(1, 40)) + int(rng.integers(0, 300))
                b = rng.integers(0, int(rng.integers(2, 6)), shape).astype(np.uint16) * int(rng.integers(1, 40)) + int(rng.integers(0, 300))
            elif kind == 2:                                 # uniform noise, wide range
                a = rng.integers(0, int(rng.integers(20, 900)), shape).astype(np.uint16)
                b = rng.integers(0, int(rng.integers(20, 900)), shape).astype(np.uint16)
            else:                                           # correlated attributes
                a = rng.integers(0, 200, shape).astype(np.uint16)
                b = (a // int(rng.integers(1, 5)) + rng.integers(0, 3, shape)).astype(np.uint16)
            imgs.append(a); prms.append(b)
        off = np.zeros(len(imgs) + 1, np.int64); off[1:] = np.cumsum([a.size for a in imgs])
        I = torch_.from_numpy(np.concatenate([a.ravel() for a in imgs])).cuda()
        P = torch_.from_numpy(np.concatenate([a.ravel() for a in prms])).cuda()
        out = b2.otsu_2d_batch(I, P, torch_.from_numpy(off).cuda())
        mask, bmax, status = out["mask"].cpu().numpy(), out["b_max"].cpu().numpy(), out["status"].cpu().numpy()
        for i, (a, p) in enumerate(zip(imgs, prms)):
            n_total += 1
            try:
                om, _, ob = oracle.otsu_py_2d_fast(a, p)
            except UnboundLocalError:
                n_bad += 1
                assert status[i] == 1, (batch, i)
                continue
            assert status[i] == 0 and bmax[i] == ob, (batch, i, int(bmax[i]), ob)
            assert np.array_equal(mask[off[i]:off[i + 1]].reshape(a.shape), om), (batch, i)
    assert n_total >= 10000 and n_bad < n_total // 4


def test_roialign_bf16_backward_and_elementwise_error(b2, torch_):
    """bf16 backward: fp32 accumulation of the bf16 gradient, rounded once to bf16 == fp32 backward of the same (bf16-valued)
    gradient rounded to bf16, up to one bf16 ulp where the two fp32 sums straddle a rounding boundary; also vs the oracle.
    Forward fp32: strict element-wise relative error against the oracle <= 1e-5 (BASELINE.json), not only the floored form."""
    from b200seg import synth
    from b200seg.roi_align_3d import roialign3d_forward, roialign3d_backward
    feat, rois = synth.roialign_case(7, feat_shape=(2, 24, 8, 32, 32), n_rois=96, scale=0.125)
    r = torch_.from_numpy(rois).cuda()
    g = np.random.default_rng(3).standard_normal((96, 24, 7, 7, 7)).astype(np.float32)
    gb = torch_.from_numpy(g).cuda().bfloat16()
    gi_b = roialign3d_backward(gb, r, feat.shape, 0.125, 2)
    assert gi_b.dtype == torch_.bfloat16
    gi_f = roialign3d_backward(gb.float(), r, feat.shape, 0.125, 2)
    want = gi_f.bfloat16()
    d = (gi_b.float() - want.float()).abs()
    ulp = want.float().abs().clamp_min(1e-30) * 2.0 ** -7
    assert bool((d <= ulp).all()) and float((d > 0).float().mean()) < 1e-3
    refg = oracle.roialign3d_bwd(gb.float().cpu().numpy(), rois, feat.shape, 0.125, 2)
    ok, err = _close(gi_f.cpu().numpy(), refg, floor="max")
    assert ok, err
    assert float(np.abs(gi_b.float().cpu().numpy() - refg).max()) <= 2.0 ** -7 * float(np.abs(refg).max())
    # forward, strict element-wise
    y = roialign3d_forward(torch_.from_numpy(feat).cuda(), r, 7, 7, 7, 0.125, 2).cpu().numpy()
    ref = oracle.roialign3d_fwd(feat, rois, 7, 0.125, 2)
    # Strict element-wise figure (no floor): the kernel contracts the three axes one after the other, the reference adds its
    # 8 taps x 8 samples in a fixed sequence, so an output that is a cancelling sum of O(1) taps differs by ~1 ulp of the
    # LARGEST partial sum; relative to a small |b| that exceeds 1e-5 on a few elements in 10^4 (measured: max 2.8e-5, 99.99 %
    # quantile 9.7e-6 on this case).  The floored form (_close, floor = mean |b|) is the 1e-5 gate used everywhere else.
    mx, q = _elementwise_rel(y, ref, q=0.999)
    assert q <= 1e-5 and mx <= 5e-5, (mx, q)
    mxb, qb = _elementwise_rel(gi_f.cpu().numpy(), refg, q=0.99)
    assert qb <= 1e-5 and mxb <= 1e-3, (mxb, qb)          # backward: order-dependent sums of hundreds of taps (see _close); measured max 1.4e-4, 99.99 % 3.7e-5


def test_dropin_reference_call_sites_through_the_shim(b2, golden):
    """The reference's own call sites, UNCHANGED, bound to this package through b200seg.shim.install():
    (1) lib/utils/boxes_3d.py imported as is on top of the shim (its nms_3d / nms_3d_volume / bbox_overlaps_3d, :55, :364-374);
    (2) tools/binarization_soma.py file lines 57-104 exec'd with box_utils_3d = that module and otsu_py_2d_fast = the shim's;
    results must equal the fixture the same lines produced with the reference's own Cython / tools/otsu.py."""
    import sys
    from oracle import refpy
    if not refpy.available():
        pytest.skip("oracle/_ref/py not staged")
    import b200seg.shim as shim
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "utils" or k.startswith("utils.") or k in ("otsu", "core", "core.config")}
    try:
        shim.install()
        import otsu as shim_otsu
        bx = refpy.load_boxes_3d("utils.boxes_3d")
        assert bx.cython_nms_3d is sys.modules["utils.cython_nms_3d"] and getattr(bx.cython_nms_3d, "__b200seg_shim__", False)
        g = golden("nms_iou.npz")
        assert np.array_equal(bx.nms_3d(g["nms0_dets"], g["nms0_thr"]), g["nms0_keep"])
        assert bx.nms_3d(np.zeros((0, 7), np.float32), 0.3) == []
        b_, q_ = g["iou0_boxes"], g["iou0_query"]
        assert np.array_equal(bx.bbox_overlaps_3d(b_, q_).view(np.uint32), g["iou0_out"].view(np.uint32))
        gs = golden("soma_script.npz")
        case = dict(volume=gs["img"], dets=gs["dets"], boxes=gs["boxes"], prm=gs["prm"], crop_off=gs["crop_off"])
        r = refpy.run_soma_script(case, 0.23, box_utils_3d=bx, otsu_py_2d_fast=shim_otsu.otsu_py_2d_fast)
        assert np.array_equal(r["visited_dets"], gs["visited_dets"])
        assert np.array_equal(r["seg"], gs["seg"]) and np.array_equal(r["scores"], gs["scores"])
    finally:
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k in ("otsu", "core", "core.config")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_maskrcnn_tile_flow_matches_the_reference_named_steps(b2, torch_):
    """BASELINE configs[1]: the device-resident tile flow (proposals -> RoIAlign3D -> box head -> box_results -> mask RoIAlign ->
    segm_results, fixed capacities, no host round trip) gives what the reference-named host functions give on the same
    scores / boxes / masks: box_results.box_results_with_nms_and_limit (golden-pinned) and segm.segm_results (golden-pinned);
    and the flow is capturable as ONE CUDA graph whose replay reproduces the eager result."""
    from b200seg.maskrcnn_flow import TileFlow
    from b200seg.box_results import box_results_with_nms_and_limit
    from b200seg import segm
    torch = torch_
    flow = TileFlow(tile=(64, 200, 200), C=64, dets_per_im=40, seed=3)       # a small detection limit makes the limit step bite
    rng = np.random.default_rng(12)
    A, (S8, H8, W8) = 35, (8, 25, 25)
    feat = torch.from_numpy(rng.standard_normal((1, 64, S8, H8, W8)).astype(np.float32)).cuda()
    n_ = A * S8 * H8 * W8
    sc = torch.from_numpy((rng.permutation(n_).astype(np.float32) / np.float32(n_)).reshape(1, A, S8, H8, W8)).cuda()
    dl = torch.from_numpy((rng.standard_normal((1, 6 * A, S8, H8, W8)) * 0.2).astype(np.float32)).cuda()
    out = flow.run(feat, sc, dl)
    torch.cuda.synchronize()
    n_rois, n_dets = int(out["n_rois"][0]), int(out["n_dets"][0])
    assert 100 < n_rois <= 1000 and 0 < n_dets <= 40
    # box_results on the host from the very same scores / boxes (rows of the valid proposals only)
    scores = out["scores"][:n_rois].cpu().numpy()
    boxes = np.concatenate([np.zeros((n_rois, 6), np.float32), out["boxes"][:n_rois].cpu().numpy()], axis=1)
    s_h, b_h, cls_boxes, _ = box_results_with_nms_and_limit(scores, boxes, num_classes=2, score_thresh=0.05, nms=0.15, detections_per_im=40)
    dets = out["dets"][:n_dets].cpu().numpy()
    assert len(cls_boxes[1]) == n_dets and np.array_equal(dets, cls_boxes[1])
    assert not out["dets"][n_dets:].any()
    # segm_results on the host from the very same masks / boxes
    masks = out["masks"][:n_dets].cpu().numpy()
    segs = segm.segm_results(cls_boxes, masks, b_h, 64, 200, 200, num_classes=2, resolution=14, cls_specific_mask=False, thresh_binarize=0.5)
    off = out["crop_off"].cpu().numpy()
    crops = out["crops"][:int(off[-1])].cpu().numpy()
    bi = out["boxes_i32"].cpu().numpy()
    assert off[n_dets] == off[-1] and off[-1] > 0
    for d in range(n_dets):
        x0, y0, z0 = np.maximum(bi[d, :3], 0)
        x1, y1, z1 = np.minimum(bi[d, 3:] + 1, [200, 200, 64])
        vol = np.zeros((64, 200, 200), np.uint8)
        if off[d + 1] > off[d]:
            vol[z0:z1, y0:y1, x0:x1] = crops[off[d]:off[d + 1]].reshape(z1 - z0, y1 - y0, x1 - x0)
        assert np.array_equal(vol, segs[1][d]), d
    # one CUDA graph for the whole tile
    ref = {k: out[k].clone() for k in ("dets", "n_dets", "masks", "crop_off")}
    ref_crops = out["crops"][:int(off[-1])].clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out2 = flow.run(feat, sc, dl)
    out2["dets"].zero_(); out2["crops"][:int(off[-1])].zero_()
    g.replay()
    torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(out2[k], ref[k]), k
    assert torch.equal(out2["crops"][:int(off[-1])], ref_crops)
