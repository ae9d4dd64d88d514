"""Build libb200seg.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python -m b200seg.build            # or: python instanceseg-without-voxelwise-labeling_b200/build.py

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box
with the gpurun snapshot.  One translation unit per operator, linked into one C-ABI library
(include/b200seg.h); no torch headers are involved.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200seg.so")
SOURCES = ["api.cu", "nms3d.cu", "iou3d.cu", "roialign3d.cu", "peaks3d.cu", "otsu2d.cu", "soma_binarize.cu", "largest_cc.cu", "paste.cu", "rle3d.cu", "mask_iou.cu", "prefilter.cu", "proposals.cu", "segm_paste.cu", "nuclei.cu", "eval.cu", "pipeline.cu", "host_batch.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(a, deps):
    return os.path.exists(a) and all(os.path.getmtime(a) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    common = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "otsu_common.cuh"), os.path.join(HERE, "..", "include", "b200seg.h")]
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or not _newer(obj, [src] + common):
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        def run(cmd):
            r = subprocess.run(cmd, capture_output=True, text=True)
            return cmd, r
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or not _newer(LIB, objs):
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
                              ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
