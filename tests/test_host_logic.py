"""CPU: host-side logic -- box truncation/clipping, crop packing, the import shim, and the
multi-process sharding/gather path on gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def test_dets_to_boxes_and_offsets():
    import b200seg
    dets = np.array([[1.9, 2.2, 0.5, 10.7, 9.1, 3.9, .5], [-3.2, -1, -2, 300, 400, 99, .4]], np.float32)
    b = b200seg.dets_to_boxes(dets, (8, 16, 32))
    assert b.dtype == np.int32
    assert b[0].tolist() == [1, 2, 0, 10, 9, 3]                 # int() truncation (binarization_soma.py:78)
    assert b[1].tolist() == [0, 0, 0, 31, 15, 7]                # clipped to the volume
    off = b200seg.crop_offsets(b)
    assert off.tolist() == [0, 10 * 8 * 4, 10 * 8 * 4 + 32 * 16 * 8]


def test_synth_case_is_deterministic():
    from b200seg import synth
    a, b = synth.postproc_case(1001), synth.postproc_case(1001)
    assert a["dets"].shape == (50, 7) and np.array_equal(a["volume"], b["volume"]) and np.array_equal(a["prm"], b["prm"])
    assert len(np.unique(a["dets"][:, 6])) == 50                # distinct scores (tie rule never exercised)
    assert a["crop_off"][-1] == a["prm"].size


def test_shim_installs_reference_names():
    import b200seg.shim as shim
    names = shim.install()
    assert "utils.cython_nms_3d" in names
    import utils.cython_nms_3d as m1
    import utils.cython_bbox_3d as m2
    from modeling.roi_xfrom.roi_align_3d.functions.roi_align_3d import RoIAlignFunction_3d
    from modeling.roi_xfrom.roi_align_3d.modules.roi_align_3d import RoIAlign_3d, RoIAlignAvg_3d, RoIAlignMax_3d
    from prm.peak_stimulation_3d import peak_stimulation_3d
    from utils.cython_mask_3d import binary_mask_to_rle, rle_to_binary_mask
    from otsu import otsu_py, otsu_py_2d, otsu_py_2d_fast      # binarization_soma.py:19, binarization_nuclei.py:12
    import pytest
    with pytest.raises(NotImplementedError):
        otsu_py(None)
    assert callable(otsu_py_2d)
    assert callable(m1.nms_3d) and callable(m1.nms_3d_volume) and callable(m2.bbox_overlaps_3d)
    f = RoIAlignFunction_3d(7, 7, 7, 0.25, 2)
    assert (f.aligned_slices, f.spatial_scale, f.sampling_ratio) == (7, 0.25, 2)
    assert RoIAlign_3d(7, 7, 7, 0.125, 2).spatial_scale == 0.125
    for k in list(sys.modules):
        if getattr(sys.modules[k], "__b200seg_shim__", False):
            del sys.modules[k]


def test_reference_wrapper_semantics_empty_and_dtype():
    import b200seg
    assert b200seg.nms_3d(np.zeros((0, 7), np.float32), 0.5) == []        # boxes_3d.py:366-367
    assert b200seg.nms_3d_volume(np.zeros((0, 7), np.float32), 0.5) == []
    with pytest.raises(ValueError):                                          # Cython buffer dtype check
        b200seg.nms_3d(np.zeros((2, 7), np.float64), 0.5)


def test_roialign_cpu_raises_not_implemented():
    import torch
    from b200seg.roi_align_3d import RoIAlignFunction_3d
    with pytest.raises(NotImplementedError):                                 # functions/roi_align_3d.py:31-32
        RoIAlignFunction_3d(7, 7, 7, 0.25, 2)(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 7))


def test_shard_indices_match_array_split():
    from b200seg.dist import shard_indices
    for n, w in [(64, 8), (64, 3), (5, 8), (0, 2)]:
        got = np.concatenate([shard_indices(n, w, r) for r in range(w)]) if w else []
        assert np.array_equal(got, np.arange(n))
        assert all(np.array_equal(shard_indices(n, w, r), np.array_split(np.arange(n), w)[r]) for r in range(w))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from b200seg.dist import shard_indices, gather_detections, allreduce_counts
    mine = shard_indices(5, world, rank)
    rng_d = [np.full((int(v) + 1, 7), float(v), np.float32) for v in mine]       # volume v has v+1 detections
    res = gather_detections(rng_d, [int(v) for v in mine])
    cnt = allreduce_counts([len(mine), sum(d.shape[0] for d in rng_d), 1])
    q.put((rank, {k: v.shape[0] for k, v in res.items()}, {k: float(v[0, 0]) for k, v in res.items()}, cnt.tolist()))
    dist.destroy_process_group()


def test_gather_detections_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shapes, vals, cnt in outs:
        assert shapes == {0: 1, 1: 2, 2: 3, 3: 4, 4: 5}
        assert vals == {v: float(v) for v in range(5)}
        assert cnt == [5, 15, 2]


def test_segm_host_logic_matches_oracle_and_scipy():
    """expand_boxes / clipped crops / the anti-aliasing tap table of segm.py against the oracle restatement and scipy."""
    import oracle
    from b200seg import segm
    rng = np.random.default_rng(12)
    boxes = (rng.uniform(-20, 200, (50, 6))).astype(np.float32)
    boxes[:, 3:] = boxes[:, :3] + rng.uniform(0, 60, (50, 3)).astype(np.float32)
    scale = (14 + 2.0) / 14
    a, b = segm.expand_boxes(boxes, scale), oracle.expand_boxes(boxes, scale)
    assert a.dtype == np.float64 and np.array_equal(a, b)
    ib = a.astype(np.int32)
    clip, off = segm.clipped_boxes(ib, 64, 100, 150)
    for d in range(len(ib)):
        x0, x1 = max(ib[d, 0], 0), min(ib[d, 3] + 1, 150)
        y0, y1 = max(ib[d, 1], 0), min(ib[d, 4] + 1, 100)
        z0, z1 = max(ib[d, 2], 0), min(ib[d, 5] + 1, 64)
        vol = (x1 - x0) * (y1 - y0) * (z1 - z0) if (x1 > x0 and y1 > y0 and z1 > z0) else 0
        assert off[d + 1] - off[d] == vol
        if vol:
            assert clip[d].tolist() == [x0, y0, z0, x1, y1, z1]
    tab = segm.gauss_table(14)
    assert tab.shape == (16, 32)
    for o in range(1, 16):
        sigma = max(0, (16 / o - 1) / 2)
        w, R = oracle.gaussian_kernel1d(sigma) if sigma > 0 else (np.ones(1), 0)
        if R >= 1:
            assert np.array_equal(tab[o, :R + 1], w[:R + 1]) and not tab[o, R + 1:].any()
        else:
            assert not tab[o].any()


def test_nuclei_host_logic():
    """Edge filter (binarization_nuclei.py:72-77, as written) and box clamping (:96-106) against a literal per-row evaluation."""
    from b200seg import binarization_nuclei as bn
    rng = np.random.default_rng(13)
    n, width, norm_side, slices = 200, 640, 200, 59
    lo = np.stack([rng.uniform(0, 600, n), rng.uniform(0, 600, n), rng.uniform(0, 50, n)], axis=1)
    dets = np.concatenate([lo, lo + rng.uniform(5, 60, (n, 3)), rng.random((n, 1))], axis=1).astype(np.float32)
    keep = bn.nuclei_edge_filter(dets, width)
    for i, d in enumerate(dets):
        c1 = d[0] > 10 and d[3] < width - 10 and d[3] - d[0] + 1 < 32
        c2 = d[1] > 10 and d[4] < width - 10 and d[4] - d[1] + 1 < 32
        assert keep[i] == (not (c1 or c2))
    tile = np.stack([rng.integers(0, 440, n), rng.integers(0, 150, n), np.zeros(n, np.int64)], axis=1)
    boxes = bn.nuclei_boxes(dets, tile, norm_side, slices)
    for i in range(n):
        w, h, s = tile[i]
        rel = (dets[i] - np.array([w, h, s, w, h, s, 0]))[:6].astype(int)
        x1, y1, z1 = max(0, rel[0]), max(0, rel[1]), max(0, rel[2])
        x2, y2, z2 = min(norm_side - 1, rel[3]), min(norm_side - 1, rel[4]), min(slices - 1, rel[5])
        assert boxes[i].tolist() == [x1 + w, y1 + h, z1 + s, x2 + w, y2 + h, z2 + s]
    rows = bn.id_det_rows(boxes[:5], dets[:5, -1], [True, False, True, True, False])
    assert rows.shape == (3, 8) and rows[:, 0].tolist() == [1.0, 3.0, 4.0] and rows.dtype == np.float64


def test_eval_host_logic_golden(golden):
    """The host parts of evaluation.py (greedy matching, precision / recall, AP, box matching) against the fixture produced
    by the reference's evaluation code; no GPU involved (overlaps come from the oracle here)."""
    import oracle
    from b200seg import evaluation as ev
    g = golden("eval.npz")
    score, match, n_pos = [], [], 0
    for k in range(int(g["count"])):
        pred, gt, ps = g["img%d_pred" % k], g["img%d_gt" % k], g["img%d_score" % k]
        ps = ps[ps[:, 1].argsort()[::-1]]
        gt_ids = np.unique(gt)[1:]
        iou = oracle.mask_overlaps(np.stack([pred == i for i in ps[:, 0]]), np.stack([gt == i for i in gt_ids]))[0]
        m = ev.match_by_iou(iou, 0.3)
        assert m == oracle.eval_volume_soma(pred, gt, ps, 0.3)[1]
        score += ps[:, 1].tolist(); match += m; n_pos += len(gt_ids)
        tp, fp, matched = ev.match_boxes_tp_fp(g["img%d_det_boxes" % k], g["img%d_gt_boxes" % k], 0.4)
        assert np.array_equal(tp, g["img%d_tp" % k]) and np.array_equal(fp, g["img%d_fp" % k]) and np.array_equal(matched, tp == 1)
    prec, rec = ev.precision_recall(score, match, n_pos)
    assert np.array_equal(prec, g["prec"]) and np.array_equal(rec, g["rec"]) and ev.voc_ap(rec, prec) == float(g["ap"])
    assert ev.match_by_iou(np.zeros((3, 0)), 0.3) == [0, 0, 0]
    tp, fp, matched = ev.match_boxes_tp_fp(np.zeros((2, 6)), np.zeros((0, 6)), 0.4)
    assert not tp.any() and not fp.any() and not matched.any()          # the script leaves tp / fp at 0 without ground truth


def test_shim_keeps_the_reference_packages_importable(tmp_path):
    """install() must not shadow the reference's own `utils` / `modeling` / `prm` packages.  A fake lib tree
    with utils/boxes_3d.py (importing the shimmed Cython names) and utils/net.py must import after install()."""
    import subprocess
    lib = tmp_path / "lib"
    (lib / "utils").mkdir(parents=True)
    (lib / "utils" / "__init__.py").write_text("")
    (lib / "utils" / "boxes_3d.py").write_text(
        "import utils.cython_bbox_3d as cython_bbox_3d\nimport utils.cython_nms_3d as cython_nms_3d\n"
        "bbox_overlaps_3d = cython_bbox_3d.bbox_overlaps_3d\n"
        "def nms_3d(dets, thresh):\n    return [] if dets.shape[0] == 0 else cython_nms_3d.nms_3d(dets, thresh)\n")
    (lib / "utils" / "net.py").write_text("X = 1\n")
    (lib / "prm").mkdir()
    (lib / "prm" / "__init__.py").write_text("")
    (lib / "prm" / "peak_backprop_3d.py").write_text("Y = 2\n")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import b200seg.shim as shim; shim.install()\n"
            "import utils.boxes_3d, utils.net, prm.peak_backprop_3d, prm.peak_stimulation_3d\n"
            "import utils, prm\n"
            "assert utils.__file__.startswith(%r) and utils.net.X == 1 and prm.peak_backprop_3d.Y == 2\n"
            "assert getattr(utils.boxes_3d.cython_nms_3d, '__b200seg_shim__', False)\n"
            "assert hasattr(prm.peak_stimulation_3d, 'peak_stimulation_3d')\n"
            "import modeling.roi_xfrom.roi_align_3d.functions.roi_align_3d as f\n"
            "assert hasattr(f, 'RoIAlignFunction_3d')\n"
            "import numpy as np; assert utils.boxes_3d.nms_3d(np.zeros((0, 7), np.float32), 0.3) == []\n"
            "print('ok')\n") % (ROOT, str(lib), str(lib))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
