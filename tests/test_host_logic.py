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
