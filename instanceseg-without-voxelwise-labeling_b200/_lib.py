"""ctypes binding of libb200seg.so (include/b200seg.h).  No CPU fallback: if the library has not
been built, or a call fails, an exception is raised."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200seg.so")


class B200SegError(RuntimeError):
    pass


_lib = None

_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/b200seg.h
SIGNATURES = {
    "b200seg_last_error": (C.c_char_p, []),
    "b200seg_version": (_i, []),
    "b200seg_launch_count": (_ll, []),
    "b200seg_set_option": (_i, [C.c_char_p, _i]),
    "b200seg_nms3d_workspace_bytes": (_sz, [_i, _i]),
    "b200seg_nms3d_dev": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_nms3d_host": (_i, [_vp, _i, _f, _i, _vp, C.POINTER(C.c_int)]),
    "b200seg_iou3d_dev": (_i, [_vp, _ll, _vp, _ll, _vp, _vp]),
    "b200seg_iou3d_host": (_i, [_vp, _ll, _vp, _ll, _vp]),
    "b200seg_roialign3d_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200seg_roialign3d_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "b200seg_roialign3d_fwd_dev": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "b200seg_roialign3d_bwd_dev": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "b200seg_peaks3d_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200seg_peaks3d_dev": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_peaks3d_fallback_count": (_i, [C.POINTER(C.c_ulonglong)]),
    "b200seg_peaks3d_bwd_dev": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200seg_otsu2d_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200seg_otsu2d_host": (_i, [_vp, _vp, _ll, _vp, C.POINTER(C.c_int)]),
    "b200seg_soma_binarize_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200seg_largest_cc_workspace_bytes": (_sz, [_ll]),
    "b200seg_largest_cc_dev": (_i, [_vp, _vp, _ll, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_paste_labels_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "b200seg_paste_labels_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_mask_overlaps_workspace_bytes": (_sz, [_i, _i]),
    "b200seg_mask_overlaps_dev": (_i, [_vp, _vp, _ll, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_gaussian3d_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "b200seg_median3d_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "b200seg_zscore_workspace_bytes": (_sz, []),
    "b200seg_zscore_norm_dev": (_i, [_vp, _i, _vp, _ll, _vp, _vp, _sz, _vp]),
    "b200seg_prm_to_u8_workspace_bytes": (_sz, [_i]),
    "b200seg_prm_to_u8_dev": (_i, [_vp, _vp, _i, _ll, _vp, _sz, _vp]),
    "b200seg_generate_proposals_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _f]),
    "b200seg_generate_proposals_capacity": (_i, [_i, _i, _i, _i, _i, _i, _f]),
    "b200seg_generate_proposals_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _f, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_largest_cc_path_counts": (_i, [_vp, _i]),
    "b200seg_rle3d_workspace_bytes": (_sz, [_i, _i, _i, _ll]),
    "b200seg_rle3d_encode_dev": (_i, [_vp, _i, _i, _i, _vp, _ll, _vp, _vp, _sz, _vp]),
    "b200seg_rle3d_decode_dev": (_i, [_vp, _ll, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "b200seg_largest_cc_ex_dev": (_i, [_vp, _vp, _ll, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "b200seg_binarize_nuclei_workspace_bytes": (_sz, [_i, _ll, _i, _i, _i]),
    "b200seg_binarize_nuclei_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_binarize_nuclei_host": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "b200seg_label_presence_dev": (_i, [_vp, _ll, _vp, _vp]),
    "b200seg_eval_voxel_counts_workspace_bytes": (_sz, [_ll]),
    "b200seg_eval_voxel_counts_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "b200seg_segm_gauss_table_size": (_i, [_i]),
    "b200seg_segm_paste_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp, _f, _i, _i, _i, _vp, _vp, _vp]),
    "b200seg_segm_expand_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "b200seg_segm_paste_host": (_i, [_vp, _ll, _vp, _vp, _i, _i, _vp, _f, _i, _i, _i, _vp, _vp]),
    "b200seg_postproc_soma_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _ll]),
    "b200seg_postproc_soma_dev": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _f, _i,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200seg_postproc_soma_host_batch": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200seg_postproc_soma_host_batch_traffic": (None, [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    "b200seg_postproc_soma_host": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _i, _vp, C.POINTER(C.c_int),
                                        _vp, _vp, _vp, _vp]),
}


def lib():
    """Load the library (once).  Raises B200SegError when it is missing -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200SegError(
                "libb200seg.so is not built (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python instanceseg-without-voxelwise-labeling_b200/build.py`; there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().b200seg_last_error().decode(errors="replace")
        raise B200SegError("%s failed (code %d): %s" % (what or "b200seg call", code, msg))


def launch_count():
    return int(lib().b200seg_launch_count())


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
