"""b200seg -- B200-native (sm_100a) operator layer of the 3D Mask R-CNN + PRM + 2D-Otsu pipeline of
MeowMeowLady/InstanceSeg-Without-Voxelwise-Labeling: RoIAlign3D fwd/bwd, 3D NMS, 3D box IoU, PRM
peak stimulation, per-instance 2D-Otsu binarization and label paste-back, as hand-written CUDA
behind a C ABI (include/b200seg.h) with the reference's Python names on top.  No CPU fallback."""
from . import _lib
from ._lib import B200SegError, launch_count
from .boxes_3d import nms_3d, nms_3d_volume, bbox_overlaps_3d, nms_3d_batched
from .otsu import otsu_py_2d_fast, otsu_2d_batch
from .binarization import (soma_binarize, largest_cc, paste_labels, SomaPostproc, postproc_soma_host, postproc_soma_host_batch,
                           dets_to_boxes, crop_offsets)

__all__ = ["nms_3d", "nms_3d_volume", "bbox_overlaps_3d", "nms_3d_batched", "otsu_py_2d_fast", "otsu_2d_batch",
           "soma_binarize", "largest_cc", "paste_labels", "SomaPostproc", "postproc_soma_host", "postproc_soma_host_batch", "dets_to_boxes", "crop_offsets",
           "B200SegError", "launch_count"]


def __getattr__(name):      # torch-dependent names are resolved lazily
    if name in ("RoIAlignFunction_3d", "RoIAlign_3d", "RoIAlignAvg_3d", "RoIAlignMax_3d"):
        import importlib
        return getattr(importlib.import_module(__name__ + ".roi_align_3d"), name)
    if name == "PeakStimulation":
        import importlib
        return getattr(importlib.import_module(__name__ + ".peak_stimulation_3d"), name)
    raise AttributeError(name)
