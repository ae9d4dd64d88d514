"""Build the reference's own native code into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The sources are compiled from where they lie under /root/reference; nothing is copied
into the repository and the only outputs are shared objects under oracle/_ref/ (git-ignored,
NOT gpurun-ignored, so they travel to the GPU box together with our own built .so files).

What is built (reference file -> output):
  lib/utils/cython_nms_3d.pyx   -> oracle/_ref/cython_nms_3d.<ext>.so
        The .pyx does not cythonize against numpy 2 (`np.int_t`, `dtype=np.int` were removed).
        A two-token patch (np.int_t -> np.int64_t, dtype=np.int -> dtype=np.int64) is applied
        to a scratch copy under /tmp; arithmetic is untouched.
  lib/utils/cython_bbox_3d.pyx  -> oracle/_ref/cython_bbox_3d.<ext>.so   (unmodified)
  lib/utils/cython_mask_3d.pyx  -> oracle/_ref/cython_mask_3d.<ext>.so   (unmodified)
  lib/modeling/roi_xfrom/roi_align_3d/src/roi_align_kernel_3d.cu
                                -> oracle/_ref/libref_roialign3d.so       (unmodified, sm_100a)
        exports ROIAlignForwardLaucher_3d / ROIAlignBackwardLaucher_3d; it is the GPU oracle
        and "the kernel to beat" for RoIAlign3D (the reference has no CPU RoIAlign).

The C is compiled with plain `gcc -O2` and no -march flag, i.e. baseline x86-64 without FMA
contraction, which is what a stock `python setup.py build_ext` of the reference produces.

Besides the shared objects, oracle/refpy.py:stage() copies five reference .py files byte for byte into oracle/_ref/py/
(same status: git-ignored build output that travels to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
load anything from oracle/.
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("B200SEG_REFERENCE", "/root/reference")
EXT = sysconfig.get_config_var("EXT_SUFFIX")


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def _cython_module(name, src_path, scratch):
    import numpy
    c_file = os.path.join(scratch, name + ".c")
    _run([sys.executable, "-m", "cython", "-3", src_path, "-o", c_file])
    out = os.path.join(OUT, name + EXT)
    _run(["gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing", "-w",
          "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
          "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
          c_file, "-o", out])
    return out


def build_ref(force=False):
    """Returns the list of built artefacts; no-op (returns existing) when the reference is absent."""
    os.makedirs(OUT, exist_ok=True)
    wanted = [os.path.join(OUT, n + EXT) for n in ("cython_nms_3d", "cython_bbox_3d", "cython_mask_3d")]
    wanted.append(os.path.join(OUT, "libref_roialign3d.so"))
    if not os.path.isdir(REF):
        return [w for w in wanted if os.path.exists(w)]
    # the reference's Python on this path (otsu.py, the two binarization scripts, boxes_3d.py, peak_stimulation_3d.py),
    # staged byte for byte next to the compiled modules: the GPU box has no /root/reference (oracle/refpy.py)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import refpy
    staged = refpy.stage()
    if not force and all(os.path.exists(w) for w in wanted):
        return wanted + staged
    scratch = tempfile.mkdtemp(prefix="b200seg_ref_")
    try:
        utils = os.path.join(REF, "lib", "utils")
        # numpy-2 patch of the NMS module, on a scratch copy only
        with open(os.path.join(utils, "cython_nms_3d.pyx")) as f:
            src = f.read()
        src = src.replace("np.int_t", "np.int64_t")
        src = re.sub(r"dtype=np\.int\b", "dtype=np.int64", src)
        patched = os.path.join(scratch, "cython_nms_3d.pyx")
        with open(patched, "w") as f:
            f.write(src)
        _cython_module("cython_nms_3d", patched, scratch)
        for name in ("cython_bbox_3d", "cython_mask_3d"):
            tmp = os.path.join(scratch, name + ".pyx")
            shutil.copy(os.path.join(utils, name + ".pyx"), tmp)   # cython writes beside the source
            _cython_module(name, tmp, scratch)
        cu_dir = os.path.join(REF, "lib", "modeling", "roi_xfrom", "roi_align_3d", "src")
        _run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared",
              "-Xcompiler", "-fPIC", "-I", cu_dir,
              os.path.join(cu_dir, "roi_align_kernel_3d.cu"),
              "-o", os.path.join(OUT, "libref_roialign3d.so")])
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return wanted + staged


if __name__ == "__main__":
    for p in build_ref(force="--force" in sys.argv):
        print("built", p)
