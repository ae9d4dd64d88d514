"""The reference's own PYTHON on this path, run as a checker / CPU baseline (TEST INFRASTRUCTURE ONLY).

`oracle/build_ref.py` stages the few reference files this needs, byte for byte, under `oracle/_ref/py/` (git-ignored build
output like the compiled `_ref` modules: it travels to the GPU box, it never enters the history):

    tools/otsu.py                      otsu_py_2d_fast  (tools/otsu.py:199-284), imported with matplotlib / skimage stubbed
    tools/binarization_soma.py         file lines 57-104 are cut out and exec'd (the script itself has hard-coded empty paths)
    tools/binarization_nuclei.py       file lines 73-148, the same way
    lib/utils/boxes_3d.py              the unchanged call site of the NMS / IoU seams (boxes_3d.py:55, :364-374)
    lib/prm/peak_stimulation_3d.py     the reference peak finder (CPU torch)

Only tests/, bench.py's cpu_baseline / --impl reference legs and tests/golden/make_golden.py may import this module.
Where the reference calls un-vendored packages that are not installed here (skimage.measure.label, cc3d, skimage
binary_closing, skimage.io) the documented scipy stand-ins of DESIGN.md section 2 are bound, exactly as in the fixtures.
"""
import importlib.util
import os
import sys
import textwrap
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(HERE, "_ref", "py")
REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
FILES = {
    "otsu.py": "tools/otsu.py",
    "binarization_soma.py": "tools/binarization_soma.py",
    "binarization_nuclei.py": "tools/binarization_nuclei.py",
    "boxes_3d.py": "lib/utils/boxes_3d.py",
    "peak_stimulation_3d.py": "lib/prm/peak_stimulation_3d.py",
}


def stage():
    """Copy the files above from the reference checkout into oracle/_ref/py/ (authoring container only)."""
    import shutil
    if not os.path.isdir(REF):
        return [p for p in (os.path.join(STAGE, n) for n in FILES) if os.path.exists(p)]
    os.makedirs(STAGE, exist_ok=True)
    out = []
    for name, rel in FILES.items():
        dst = os.path.join(STAGE, name)
        shutil.copyfile(os.path.join(REF, rel), dst)
        out.append(dst)
    return out


def path(name):
    p = os.path.join(STAGE, name)
    if os.path.exists(p):
        return p
    q = os.path.join(REF, FILES[name])
    if os.path.exists(q):
        return q
    raise FileNotFoundError("reference file %s is not staged (run oracle/build_ref.py where /root/reference exists)" % name)


def available():
    try:
        for n in FILES:
            path(n)
        return True
    except FileNotFoundError:
        return False


_mods = {}


def load_otsu():
    """tools/otsu.py as a module (matplotlib / skimage stubbed, np.float restored: the file predates numpy 1.24)."""
    if "otsu" not in _mods:
        for m in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io", "skimage.exposure"):
            sys.modules.setdefault(m, types.ModuleType(m))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["skimage"].io = sys.modules["skimage.io"]
        sys.modules["skimage"].exposure = sys.modules["skimage.exposure"]
        if not hasattr(np, "float"):
            np.float = float
        spec = importlib.util.spec_from_file_location("ref_otsu", path("otsu.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods["otsu"] = mod
    return _mods["otsu"]


def load_peaks():
    if "peaks" not in _mods:
        spec = importlib.util.spec_from_file_location("ref_peak_stimulation_3d", path("peak_stimulation_3d.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods["peaks"] = mod
    return _mods["peaks"]


def load_boxes_3d(as_name="utils.boxes_3d"):
    """lib/utils/boxes_3d.py imported UNCHANGED under `as_name`; `utils.cython_nms_3d` / `utils.cython_bbox_3d` must already
    resolve (to the _ref Cython build, or to b200seg.shim's modules for the drop-in test).  core.config is stubbed with the
    two fields the module reads at import time."""
    if "core.config" not in sys.modules:
        core = sys.modules.setdefault("core", types.ModuleType("core"))
        cfgm = types.ModuleType("core.config")
        cfgm.cfg = types.SimpleNamespace()
        sys.modules["core.config"] = cfgm
        core.config = cfgm
    spec = importlib.util.spec_from_file_location(as_name, path("boxes_3d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_cython_namespace():
    """A `box_utils_3d` stand-in bound to the reference's own Cython NMS (oracle/_ref)."""
    import oracle
    nms_mod = oracle.ref_module("cython_nms_3d")
    f32 = lambda d: np.ascontiguousarray(d, dtype=np.float32)
    return types.SimpleNamespace(nms_3d=lambda d, t: nms_mod.nms_3d(f32(d), np.float32(t)),
                                 nms_3d_volume=lambda d, t: nms_mod.nms_3d_volume(f32(d), np.float32(t)))


def scipy_label(m):
    from scipy import ndimage as ndi
    return ndi.label(m, structure=np.ones((3, 3, 3), bool))[0]


# ---------------------------------------------------------------------------------------------------------------------
# tools/binarization_soma.py:57-104 on an in-memory case (synth.postproc_case layout)
# ---------------------------------------------------------------------------------------------------------------------
_code = {}


def _soma_code():
    if "soma" not in _code:
        lines = open(path("binarization_soma.py")).read().split("\n")
        head = textwrap.dedent("\n".join(lines[56:62]))            # file lines 57..62: NMS + descending-score visit order
        loop = textwrap.dedent("\n".join(lines[63:104]))           # file lines 64..104: score table + the per-instance loop
        assert head.startswith("keep = box_utils_3d.nms_3d(dets, nms_thresh)"), head[:80]
        assert loop.startswith("scores = np.zeros((0, 2)") and loop.rstrip().split("\n")[-1].lstrip().startswith("scores = np.concatenate")
        _code["soma"] = (compile(head, "ref_binarization_soma_57_62", "exec"), compile(loop, "ref_binarization_soma_64_104", "exec"))
    return _code["soma"]


def soma_tiles(case, tile=(64, 160, 160)):
    """Tile origins (ws, hs, ss) such that every box lies inside its 64 x 160 x 160 tile (the script's hard-coded crop,
    binarization_soma.py:71-72), and a lazy `io.imread` that serves the PRM "tif" of an instance = its response in that tile."""
    b = case["boxes"]
    S, H, W = case["volume"].shape
    ts, th, tw = tile
    ws = np.clip(b[:, 0] - (tw - (b[:, 3] - b[:, 0] + 1)) // 2, 0, max(0, W - tw))
    hs = np.clip(b[:, 1] - (th - (b[:, 4] - b[:, 1] + 1)) // 2, 0, max(0, H - th))
    ss = np.clip(b[:, 2] - (ts - (b[:, 5] - b[:, 2] + 1)) // 2, 0, max(0, S - ts))
    n = len(b)
    idx = np.stack([np.arange(n), np.arange(n), ws, hs, ss], axis=1).astype(int)

    def imread(p):
        i = int(os.path.basename(p).split(".")[0])
        ob = b[i]
        t = np.zeros((min(ts, S - ss[i]), min(th, H - hs[i]), min(tw, W - ws[i])), np.uint8)
        full = case["prm"][case["crop_off"][i]:case["crop_off"][i + 1]].reshape(ob[5] - ob[2] + 1, ob[4] - ob[1] + 1, ob[3] - ob[0] + 1)
        t[ob[2] - ss[i]:ob[5] + 1 - ss[i], ob[1] - hs[i]:ob[4] + 1 - hs[i], ob[0] - ws[i]:ob[3] + 1 - ws[i]] = full
        return t
    return idx, types.SimpleNamespace(imread=imread)


def run_soma_script(case, nms_thresh=0.23, box_utils_3d=None, otsu_py_2d_fast=None, label=None, max_instances=None):
    """Executes the reference script's own lines on `case`.  Bindings default to the reference's code (Cython NMS from
    oracle/_ref, tools/otsu.py, scipy stand-in for skimage.measure.label); the drop-in test passes this repo's shim functions
    instead.  max_instances: visit only the first k detections of the script's own visit order (bounded CPU sample).
    Returns dict(seg, scores, visited_dets, n_after_nms, t_nms, t_loop)."""
    head, loop = _soma_code()
    idx, io = soma_tiles(case)
    img = case["volume"]
    ns = {"np": np, "os": os, "io": io, "dets": case["dets"].copy(), "instance_idex": idx.copy(), "nms_thresh": nms_thresh, "img": img,
          "seg": np.zeros(img.shape, np.uint16), "mask_id": 0, "prm_path": "p", "im_name": "x",
          "box_utils_3d": box_utils_3d or ref_cython_namespace(),
          "otsu_py_2d_fast": otsu_py_2d_fast or load_otsu().otsu_py_2d_fast, "label": label or scipy_label}
    old = np.seterr(all="ignore")
    try:
        t0 = time.perf_counter()
        exec(head, ns)
        t1 = time.perf_counter()
        n_after = len(ns["dets"])
        if max_instances is not None:
            ns["dets"] = ns["dets"][:max_instances]
            ns["instance_idex"] = ns["instance_idex"][:max_instances]
        exec(loop, ns)
        t2 = time.perf_counter()
    finally:
        np.seterr(**old)
    return dict(seg=ns["seg"], scores=ns["scores"], visited_dets=ns["dets"], n_after_nms=n_after, n_visited=len(ns["dets"]),
                t_nms=t1 - t0, t_loop=t2 - t1)
