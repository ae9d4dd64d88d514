"""Make the reference's import names resolve to this package, so that lib/modeling/model_builder.py,
lib/utils/boxes_3d.py, lib/prm/*, tools/binarization_*.py and tools/infer_simple.py keep working
unchanged.

    import b200seg.shim; b200seg.shim.install()

registers these modules in sys.modules:
    utils.cython_nms_3d          -> nms_3d, nms_3d_volume
    utils.cython_bbox_3d         -> bbox_overlaps_3d
    modeling.roi_xfrom.roi_align_3d.functions.roi_align_3d -> RoIAlignFunction_3d
    modeling.roi_xfrom.roi_align_3d.modules.roi_align_3d   -> RoIAlign_3d, RoIAlignAvg_3d, RoIAlignMax_3d
    prm.peak_stimulation_3d      -> peak_stimulation_3d, PeakStimulation
    utils.cython_mask_3d         -> binary_mask_to_rle, rle_to_binary_mask
    modeling.generate_proposals_3d -> GenerateProposalsOp_3d
    b200seg_core_test            -> box_results_with_nms_and_limit, segm_results (core/test.py has many other members: bind
                                    these names into it, see INTEGRATION.md, instead of replacing the module)
    otsu                         -> otsu_py_2d_fast, otsu_py_2d (+ import-compat stubs otsu_py, otsu_mat)
"""
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__b200seg_shim__ = True
    return m


def install(overwrite=True):
    from . import boxes_3d, roi_align_3d, peak_stimulation_3d, otsu, mask_3d, generate_proposals_3d, box_results, segm
    mods = {
        "utils.cython_nms_3d": _module("utils.cython_nms_3d", nms_3d=boxes_3d._nms_numpy and
                                       (lambda dets, thresh: boxes_3d._nms_numpy(dets, thresh, False)),
                                       nms_3d_volume=lambda dets, thresh: boxes_3d._nms_numpy(dets, thresh, True)),
        "utils.cython_bbox_3d": _module("utils.cython_bbox_3d", bbox_overlaps_3d=boxes_3d.bbox_overlaps_3d),
        "modeling.roi_xfrom.roi_align_3d.functions.roi_align_3d":
            _module("modeling.roi_xfrom.roi_align_3d.functions.roi_align_3d",
                    RoIAlignFunction_3d=roi_align_3d.RoIAlignFunction_3d),
        "modeling.roi_xfrom.roi_align_3d.modules.roi_align_3d":
            _module("modeling.roi_xfrom.roi_align_3d.modules.roi_align_3d",
                    RoIAlign_3d=roi_align_3d.RoIAlign_3d, RoIAlignAvg_3d=roi_align_3d.RoIAlignAvg_3d,
                    RoIAlignMax_3d=roi_align_3d.RoIAlignMax_3d),
        "prm.peak_stimulation_3d": _module("prm.peak_stimulation_3d",
                                           peak_stimulation_3d=peak_stimulation_3d.peak_stimulation_3d,
                                           PeakStimulation=peak_stimulation_3d.PeakStimulation),
        "utils.cython_mask_3d": _module("utils.cython_mask_3d", binary_mask_to_rle=mask_3d.binary_mask_to_rle,
                                        rle_to_binary_mask=mask_3d.rle_to_binary_mask),
        "modeling.generate_proposals_3d": _module("modeling.generate_proposals_3d",
                                                  GenerateProposalsOp_3d=generate_proposals_3d.GenerateProposalsOp_3d),
        "b200seg_core_test": _module("b200seg_core_test", box_results_with_nms_and_limit=box_results.box_results_with_nms_and_limit,
                                     segm_results=segm.segm_results),
        "otsu": _module("otsu", otsu_py_2d_fast=otsu.otsu_py_2d_fast, otsu_py_2d=otsu.otsu_py_2d,
                        otsu_py=otsu.otsu_py, otsu_mat=otsu.otsu_mat),
    }
    installed = []
    for name, mod in mods.items():
        if overwrite or name not in sys.modules:
            parts = name.split(".")
            # Parents first.  The reference's own packages (lib/utils, lib/modeling, lib/prm: they have __init__.py and hold
            # many modules this layer does not replace -- utils.boxes_3d, utils.net, modeling.model_builder, ...) must stay
            # importable, so a parent is taken from sys.path when it is there and only fabricated when it is not; a
            # fabricated parent still points its __path__ at any matching directory found on sys.path.
            for i in range(1, len(parts)):
                parent = ".".join(parts[:i])
                if parent not in sys.modules:
                    sys.modules[parent] = _import_or_fabricate(parent)
            sys.modules[name] = mod
            installed.append(name)
            for i in range(1, len(parts)):
                setattr(sys.modules[".".join(parts[:i])], parts[i], sys.modules[".".join(parts[:i + 1])])
    return installed


def _import_or_fabricate(parent):
    import importlib
    import os
    try:
        return importlib.import_module(parent)          # the real package (e.g. the reference's lib/utils) wins
    except Exception:
        pass
    pm = types.ModuleType(parent)
    rel = parent.split(".")
    grand = ".".join(rel[:-1])
    roots = list(getattr(sys.modules.get(grand), "__path__", [])) if grand else list(sys.path)
    pm.__path__ = [d for d in (os.path.join(r or ".", rel[-1]) for r in roots) if os.path.isdir(d)]
    pm.__b200seg_shim__ = True
    return pm
