"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/b200seg.h
declares (no compute without a GPU); the product path fails loudly without a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200seg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built_lib):
    L = built_lib.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    # and the ctypes table binds exactly the header's surface
    assert sorted(built_lib.SIGNATURES) == syms


def test_version_and_error_string(built_lib):
    L = built_lib.lib()
    assert L.b200seg_version() >= 100
    assert isinstance(L.b200seg_last_error(), bytes)
    assert L.b200seg_launch_count() >= 0


def test_argument_validation_without_gpu(built_lib):
    L = built_lib.lib()
    # invalid arguments are rejected before any CUDA call
    assert L.b200seg_iou3d_dev(None, -1, None, 3, None, None) == -1
    assert b"negative" in L.b200seg_last_error()
    assert L.b200seg_peaks3d_dev(None, 1, 1, 4, 4, 4, 4, 0, None, None, 0, None, None, None, None, 0, None) == -1
    assert L.b200seg_roialign3d_fwd_dev(None, 0, None, None, 1, 1, 4, 4, 4, 1, 32, 7, 7, 0.25, 2, 0, None) == -1
    assert L.b200seg_nms3d_workspace_bytes(2, 1000) > 2 * 1000 * 16 * 8
    assert L.b200seg_roialign3d_workspace_bytes(512, 8, 32, 32, 7) >= 512 * (32 + 72 * 8 * 4)
    assert L.b200seg_paste_labels_workspace_bytes(2, 128, 512, 512, 800) >= 2 * 8192 * 25 * 4
    # entry points added later in the round: same contract (negative code + message, empty work is a no-op)
    assert L.b200seg_segm_paste_dev(None, None, None, 3, 0, None, 0.5, 8, 8, 8, None, None, None) == -1
    assert b"segm_paste" in L.b200seg_last_error()
    assert L.b200seg_segm_paste_dev(None, None, None, 0, 14, None, 0.5, 8, 8, 8, None, None, None) == 0
    assert L.b200seg_segm_paste_dev(None, None, None, 2, 14, None, 0.5, 8, 8, 8, None, None, None) == -1      # null pointers
    assert L.b200seg_segm_gauss_table_size(14) == 16 * 32
    assert L.b200seg_segm_expand_dev(None, None, None, 0, 8, 8, 8, None, None) == 0
    assert L.b200seg_binarize_nuclei_dev(None, 3, 8, 8, 8, None, None, None, 0, 0, None, None, None, None, None, None, 0, None) == -1
    assert b"binarize_nuclei" in L.b200seg_last_error()
    assert L.b200seg_binarize_nuclei_workspace_bytes(50, 3000000, 59, 350, 640) > 3000000 * (2 + 2 + 1 + 8)
    assert L.b200seg_eval_voxel_counts_workspace_bytes(128 * 512 * 512) >= 128 * 512 * 512 // 8
    assert L.b200seg_eval_voxel_counts_dev(None, None, 8, 8, 8, None, 0, None, None, 0, None) == -1
    assert L.b200seg_label_presence_dev(None, -1, None, None) == -1
    assert L.b200seg_largest_cc_ex_dev(None, None, -1, 1, None, 4, None, None, None, None, 1, None, 0, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import b200seg
    d = np.zeros((3, 7), np.float32)
    with pytest.raises(b200seg.B200SegError):
        b200seg.nms_3d(d, 0.5)
    with pytest.raises(b200seg.B200SegError):
        b200seg.bbox_overlaps_3d(d[:, :6].copy(), d[:, :6].copy())
    with pytest.raises(b200seg.B200SegError):
        b200seg.otsu_py_2d_fast(np.arange(8, dtype=np.uint16).reshape(2, 2, 2), np.arange(8, dtype=np.uint16).reshape(2, 2, 2))


def test_product_does_not_import_oracle():
    """The package must never route through oracle/ (grep of the product sources)."""
    pkg = os.path.join(ROOT, "instanceseg-without-voxelwise-labeling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
