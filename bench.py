#!/usr/bin/env python
"""bench.py -- headline benchmark of the b200seg hot path (contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code for the path, all host cores

A "step" = one pass of the post-processing pipeline of BASELINE.json configs[2]/[4] over the batch of `--volumes`
(64) synthetic 128x512x512 uint8 volumes (~200 blob instances, 800 candidate detections, one (14,32,128,128) fp32
response map each): PRM peak stimulation (lib/prm/peak_stimulation_3d.py, median filter) -> tools/binarization_soma.py
:57-104 (3D NMS -> visit order -> per-instance crop/normalise/2D-Otsu -> largest connected component -> label
paste-back).  STRONG scaling by default: the 64 volumes are split over the N ranks (np.array_split, the reference's
my_subprocess.py:56), so N = 1 processes all 64; `--scaling weak --volumes-per-rank V` keeps V volumes per rank.
No data-path collective; the only exchange is the all-gather of the surviving detections at the end of each step.

One JSON line is printed by rank 0:
  value          Gvox/s, inputs resident in HBM, CUDA-event timed, max over ranks (peak stimulation included;
                 `value_without_peaks` = the binarization chain alone)
  e2e            the binarization chain through the host-buffer C-ABI call (pinned host memory in, H2D + D2H inside)
  roofline       dominant kernel: algorithmic bytes per launch / measured launch time vs the measured HBM peak
  cpu_baseline   kind "reference": the reference's own Python (tools/otsu.py + Cython NMS + the script's numpy, staged in
                 oracle/_ref) on a bounded sample; `port` = the C restatement (oracle/) beside it
  parity_checked number of this run's volumes whose GPU result was compared bit for bit with the oracle chain
  ops            per-operator numbers (RoIAlign3D fwd/bwd incl. the reference's own CUDA kernel, peaks, IoU, NMS, ...),
                 each with its roofline fraction and, where the reference has a CPU implementation, its time beside it
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (128, 512, 512)
NMS_THRESH = 0.23
CASE_KW = dict(shape=SHAPE, n_blobs=200, n_dup=400, n_false=200, sigma_xy=(5, float(os.environ.get("B200SEG_BENCH_SIGMA_MAX", "11"))),
               sigma_z=(3, 6))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_case(seed):
    from b200seg import synth
    return synth.postproc_case(seed, **CASE_KW)


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code for the chain, all host cores
# ------------------------------------------------------------------------------------------------
_CASES = {}
REF_SAMPLE_INSTANCES = 24          # detections of the script's own visit order processed per volume and step (bounded sample)


def volume_split(n_volumes, world, rank):
    """Volumes of rank `rank`: np.array_split(range(n), world)[rank] (my_subprocess.py:56)."""
    return [int(v) for v in np.array_split(np.arange(n_volumes), world)[rank]]


def load_oracle_in_parent():
    """Load the oracle libraries in THIS process (before forking workers) so that a loaded-library record shows them."""
    import oracle
    oracle.lib()
    try:
        oracle.ref_module("cython_nms_3d"); oracle.ref_module("cython_bbox_3d")
    except Exception:
        pass
    return oracle


def _ref_worker(job):
    """One volume, the reference's own code: tools/binarization_soma.py:57-104 (Cython NMS + tools/otsu.py + numpy paste)
    on the first `k` detections of its visit order.  Returns (seconds, volume fraction done, seconds of the C port or None)."""
    seed, k, want_port = job
    from oracle import refpy
    if seed not in _CASES:
        _CASES[seed] = make_case(seed)
    c = _CASES[seed]
    t = time.perf_counter()
    r = refpy.run_soma_script(c, NMS_THRESH, max_instances=k)
    dt = time.perf_counter() - t
    frac = r["n_visited"] / max(r["n_after_nms"], 1)
    tp = None
    if want_port:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle_chain
        t = time.perf_counter()
        oracle_chain(c, NMS_THRESH)
        tp = time.perf_counter() - t
    return dt, frac, tp


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import refpy
    load_oracle_in_parent()
    if not refpy.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/py (staged reference Python) is missing: run build() where /root/reference exists"}))
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, args.volumes))
    seeds = [args.seed_base + i for i in range(workers)]
    V = int(np.prod(SHAPE))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        first = pool.map(_ref_worker, [(s_, 2, True) for s_ in seeds])                # builds the cases + times the C port once
        for _ in range(max(0, args.warmup - 1)):
            pool.map(_ref_worker, [(s_, 2, False) for s_ in seeds])
        t0 = time.perf_counter()
        fr = []
        for _ in range(args.steps):
            fr.append(pool.map(_ref_worker, [(s_, REF_SAMPLE_INSTANCES, False) for s_ in seeds]))
        dt = time.perf_counter() - t0
    # volumes' worth of work done per step = sum over workers of the fraction of the visit list they processed
    vol_equiv = float(np.mean([sum(f for _, f, _ in step) for step in fr]))
    val = vol_equiv * V * args.steps / dt / 1e9
    port = workers * V / max(t for _, _, t in first) / 1e9
    line = {"impl": "reference", "metric": "postproc_gvox_per_s", "value": val, "unit": "Gvox/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world, len(volume_split(args.volumes, world, 0)) if args.scaling == "strong" else args.volumes_per_rank),
            "cpu_baseline": {"value": val, "unit": "Gvox/s", "cores": workers, "kind": "reference",
                             "sample": "per step every one of %d worker processes runs tools/binarization_soma.py:57-104 (reference Cython nms_3d, "
                                       "tools/otsu.py otsu_py_2d_fast, the script's numpy crop / paste / np.unique; skimage label -> scipy) on one "
                                       "128x512x512 volume for the first %d of ~390 detections of its visit order; value = volume fractions done x "
                                       "voxels / time; peak stimulation not included; host has %d cores" % (workers, REF_SAMPLE_INSTANCES, cores),
                             "port": {"value": port, "unit": "Gvox/s", "cores": workers,
                                      "note": "the oracle's C restatement of the same chain, whole volumes, one per worker"}},
            "e2e": {"value": val, "unit": "Gvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, world, vpr):
    nd = CASE_KW["n_blobs"] + CASE_KW["n_dup"] + CASE_KW["n_false"]
    return {"workload": "postproc_soma_chain: PRM peak stimulation (14x32x128x128 fp32 map, win 3, median) + 3D NMS(0.23) + per-instance 2D-Otsu + "
                        "largest connected component + label paste-back on synthetic uint8 128x512x512 volumes, ~200 blobs, %d candidate "
                        "detections each (BASELINE configs[2]/[4])" % nd,
            "scaling": args.scaling, "global_volumes_per_step": vpr * world if args.scaling == "weak" else args.volumes,
            "volumes_per_rank": vpr, "volume_shape": list(SHAPE), "response_map_shape": [14, 32, 128, 128],
            "dets_per_volume": nd, "nms_thresh": NMS_THRESH,
            "l2": "per-step inputs+outputs (%.0f MB per rank) exceed the 126 MB L2; no explicit flush in the chain loop, "
                  "explicit 256 MB flush between iterations of the per-operator timings" % (vpr * 3 * np.prod(SHAPE) / 1e6),
            "parallelism": "volumes sharded by rank (np.array_split), dp%d" % world}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_op(torch, fn, iters, flush):
    """CUDA-event time of fn() per call, L2 flushed (256 MB write) before each timed call."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def _ref_roialign_lib():
    """The reference's own CUDA kernels (roi_align_kernel_3d.cu compiled unmodified into oracle/_ref): "the kernel to beat"."""
    import ctypes
    p_ = os.path.join(ROOT, "oracle", "_ref", "libref_roialign3d.so")
    if not os.path.exists(p_):
        return None
    L = ctypes.CDLL(p_)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.ROIAlignForwardLaucher_3d.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    L.ROIAlignBackwardLaucher_3d.argtypes = [vp, cf, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp]
    return L


def bench_ops(torch, peak, with_cpu=True):
    """Per-operator timings on the BASELINE shapes (rank-local).  with_cpu: also time the reference's CPU implementation
    of the operators that have one (Cython NMS / IoU from oracle/_ref, CPU-torch peak_stimulation_3d) on this host."""
    import ctypes
    import b200seg
    from b200seg import synth
    from b200seg.roi_align_3d import roialign3d_forward, roialign3d_backward
    from b200seg.peak_stimulation_3d import peaks_forward, PeaksPlan
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops = {}

    def entry(ms, bytes_alg, extra):
        gbs = bytes_alg / (ms * 1e-3) / 1e9
        d = {"ms": ms, "alg_bytes": bytes_alg, "achieved_gbs": gbs, "roofline_frac": gbs / peak}
        d.update(extra)
        return d

    def cpu_time(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
        return best * 1e3

    ref_nms = ref_bbox = None
    if with_cpu:
        try:
            import oracle
            ref_nms, ref_bbox = oracle.ref_module("cython_nms_3d"), oracle.ref_module("cython_bbox_3d")
        except Exception:
            pass

    # RoIAlign3D, config 4: 512 RoIs x 256 ch x 7^3 on (2,256,8,32,32), scale 1/8, sr 2
    feat, rois = synth.roialign_case(1004)
    f, r = torch.from_numpy(feat).to(dev), torch.from_numpy(rois).to(dev)
    R, C, P = rois.shape[0], feat.shape[1], 7
    out_b, feat_b = R * C * P ** 3 * 4, feat.size * 4
    ms = time_op(torch, lambda: roialign3d_forward(f, r, P, P, P, 0.125, 2), 20, flush)
    ops["roialign3d_fwd_f32"] = entry(ms, out_b + feat_b + 28 * R, {"rois_per_s": R / (ms * 1e-3)})
    g = torch.randn((R, C, P, P, P), device=dev)
    ms = time_op(torch, lambda: roialign3d_backward(g, r, feat.shape, 0.125, 2), 20, flush)
    ops["roialign3d_bwd_f32"] = entry(ms, out_b + feat_b + 28 * R, {"rois_per_s": R / (ms * 1e-3)})
    fb = f.bfloat16()
    ms = time_op(torch, lambda: roialign3d_forward(fb, r, P, P, P, 0.125, 2), 20, flush)
    ops["roialign3d_fwd_bf16"] = entry(ms, (out_b + feat_b) // 2 + 28 * R, {"rois_per_s": R / (ms * 1e-3)})
    gb = g.bfloat16()
    ms = time_op(torch, lambda: roialign3d_backward(gb, r, feat.shape, 0.125, 2), 20, flush)
    ops["roialign3d_bwd_bf16"] = entry(ms, (out_b + feat_b) // 2 + 28 * R, {"rois_per_s": R / (ms * 1e-3)})
    # the reference's own CUDA kernels on the same inputs (launchers roi_align_kernel_3d.cu:153-172, :340-360; the caller
    # zero-fills both outputs first, functions/roi_align_3d.py:26-28, :41-42 -- included, it is part of the reference's op)
    RL = _ref_roialign_lib()
    if RL is not None:
        vp, cf = ctypes.c_void_p, ctypes.c_float
        B_, _, S_, H_, W_ = feat.shape
        out_ref = torch.empty((R, C, P, P, P), device=dev)
        gin_ref = torch.empty(feat.shape, device=dev)

        def ref_fwd():
            out_ref.zero_()
            RL.ROIAlignForwardLaucher_3d(vp(f.data_ptr()), cf(0.125), R, S_, H_, W_, C, P, P, P, 2, vp(r.data_ptr()), vp(out_ref.data_ptr()),
                                         vp(torch.cuda.current_stream().cuda_stream))

        def ref_bwd():
            gin_ref.zero_()
            RL.ROIAlignBackwardLaucher_3d(vp(g.data_ptr()), cf(0.125), B_, R, S_, H_, W_, C, P, P, P, 2, vp(r.data_ptr()), vp(gin_ref.data_ptr()),
                                          vp(torch.cuda.current_stream().cuda_stream))
        for nm, fn, ours in (("roialign3d_fwd_f32_reference_kernel", ref_fwd, "roialign3d_fwd_f32"), ("roialign3d_bwd_f32_reference_kernel", ref_bwd, "roialign3d_bwd_f32")):
            ms = time_op(torch, fn, 20, flush)
            ops[nm] = entry(ms, out_b + feat_b + 28 * R, {"rois_per_s": R / (ms * 1e-3),
                                                          "note": "reference roi_align_kernel_3d.cu compiled unmodified for sm_100a (oracle/_ref), incl. the caller's zero fill"})
            ops[ours]["vs_ref_kernel"] = ms / ops[ours]["ms"]
    # peak stimulation, config 3: fp32 (1,14,32,128,128), win 3, median filter
    x = torch.from_numpy(synth.response_map(np.random.default_rng(1003), (32, 128, 128), n_peaks=60, channels=14)).to(dev)
    pplan = PeaksPlan(x.shape, dev, 3, 1)
    ms = time_op(torch, lambda: pplan.run(x), 10, flush)
    ops["peaks3d_f32_14x32x128x128"] = entry(ms, 8 * x.numel(), {"gvox_per_s": x.numel() / (ms * 1e-3) / 1e9,
                                                                  "note": "device op (memset + 4 launches, count stays on the device)"})
    ms2 = time_op(torch, lambda: peaks_forward(x, 3, 1), 10, flush)
    ops["peaks3d_f32_14x32x128x128"]["ms_with_exact_size_list"] = ms2
    x8 = x.repeat(8, 1, 1, 1, 1).contiguous()
    pplan8 = PeaksPlan(x8.shape, dev, 3, 1)
    ms = time_op(torch, lambda: pplan8.run(x8), 10, flush)
    ops["peaks3d_f32_8x14x32x128x128"] = entry(ms, 8 * x8.numel(), {"gvox_per_s": x8.numel() / (ms * 1e-3) / 1e9, "note": "8 volumes' maps in one call"})
    xf = torch.from_numpy(synth.response_map(np.random.default_rng(1003), (128, 512, 512), n_peaks=200, channels=1)).to(dev)
    pplanf = PeaksPlan(xf.shape, dev, 3, 1)
    ms = time_op(torch, lambda: pplanf.run(xf), 10, flush)
    ops["peaks3d_f32_1x128x512x512"] = entry(ms, 8 * xf.numel(), {"gvox_per_s": xf.numel() / (ms * 1e-3) / 1e9, "note": "full-resolution variant of config 3"})
    del x8, xf, pplan8, pplanf
    if with_cpu:
        try:
            from oracle import refpy
            rp = refpy.load_peaks()
            xc = x.cpu()
            medf = lambda inp: torch.median(inp.view(inp.size(0), inp.size(1), -1), dim=2)[0].contiguous().view(inp.size(0), inp.size(1), 1, 1, 1)
            ops["peaks3d_f32_14x32x128x128"]["cpu_ref_ms"] = cpu_time(lambda: rp.peak_stimulation_3d(xc, win_size=3, peak_filter=medf), 2)
            ops["peaks3d_f32_14x32x128x128"]["cpu_ref"] = "lib/prm/peak_stimulation_3d.py on CPU torch (%d threads)" % torch.get_num_threads()
        except Exception as e:                                   # staged file missing
            ops["peaks3d_f32_14x32x128x128"]["cpu_ref"] = "unavailable: %s" % e
    # IoU matrix: anchors x gt (917504 x 50)
    rng = np.random.default_rng(10)
    bx_np = synth.random_dets(rng, 917504, extent=(256, 256, 64), side=(8, 64))[:, :6].copy()
    q_np = synth.random_dets(rng, 50, extent=(256, 256, 64), side=(10, 40))[:, :6].copy()
    bx, q = torch.from_numpy(bx_np).to(dev), torch.from_numpy(q_np).to(dev)
    ms = time_op(torch, lambda: b200seg.bbox_overlaps_3d(bx, q), 20, flush)
    ops["iou3d_917504x50"] = entry(ms, 24 * (917504 + 50) + 4 * 917504 * 50, {"pairs_per_s": 917504 * 50 / (ms * 1e-3)})
    if ref_bbox is not None:
        ops["iou3d_917504x50"]["cpu_ref_ms"] = cpu_time(lambda: ref_bbox.bbox_overlaps_3d(bx_np, q_np), 2)
        ops["iou3d_917504x50"]["cpu_ref"] = "reference cython_bbox_3d.bbox_overlaps_3d (oracle/_ref), 1 core"
    # RLE codec of a 128x512x512 label mask (~200 blobs): encode reads the mask, decode writes it
    from b200seg import mask_3d, _lib as L_
    mvol = torch.from_numpy((synth.postproc_case(2000, **CASE_KW)["volume"] > 60).astype(np.uint8)).to(dev)
    S_, H_, W_ = mvol.shape
    cap = 1 << 22
    ms = time_op(torch, lambda: mask_3d._encode_device(mvol, cap), 10, flush)
    counts, n_c = mask_3d._encode_device(mvol, cap)
    ops["rle3d_encode_128x512x512"] = entry(ms, mvol.numel() + 8 * n_c, {"gvox_per_s": mvol.numel() / (ms * 1e-3) / 1e9, "runs": n_c,
                                                                       "note": "two walks over the mask (count, emit) + scan; includes the D2H read of the run count"})
    lib_ = L_.lib()
    ws_b = lib_.b200seg_rle3d_workspace_bytes(S_, H_, W_, n_c)
    ws_t = torch.empty(ws_b, dtype=torch.uint8, device=dev)
    out_m = torch.empty_like(mvol)
    cc = counts[:n_c].contiguous()
    ms = time_op(torch, lambda: L_.check(lib_.b200seg_rle3d_decode_dev(L_.ptr(cc), n_c, L_.ptr(out_m), S_, H_, W_, None, L_.ptr(ws_t), ws_b,
                                                                       L_.current_stream()), "rle3d_decode"), 10, flush)
    assert bool((out_m == mvol).all())
    ops["rle3d_decode_128x512x512"] = entry(ms, mvol.numel() + 8 * n_c, {"gvox_per_s": mvol.numel() / (ms * 1e-3) / 1e9})
    # evaluation: pairwise mask overlaps of two 128x512x512 label volumes (~200 instances each), one joint-histogram pass
    from b200seg import evaluation
    rng_e = np.random.default_rng(21)
    lab_p = np.zeros(SHAPE, np.uint16); lab_g = np.zeros(SHAPE, np.uint16)
    ev_box_list = []
    for lab, shift in ((lab_p, 0), (lab_g, 3)):
        for i in range(1, 201):
            c = [int(rng_e.integers(8, s - 40)) for s in SHAPE]; e = [int(rng_e.integers(8, 18)), int(rng_e.integers(16, 40)), int(rng_e.integers(16, 40))]
            lab[c[0] + shift:c[0] + e[0], c[1]:c[1] + e[1] + shift, c[2]:c[2] + e[2]] = i
            if shift == 0:
                ev_box_list.append([c[2], c[1], c[0], c[2] + e[2] - 1, c[1] + e[1] - 1, c[0] + e[0] - 1])
    tp, tg = torch.from_numpy(lab_p).to(dev), torch.from_numpy(lab_g).to(dev)
    ids = np.arange(1, 201)
    ms = time_op(torch, lambda: evaluation.mask_overlaps_labels(tp, tg, ids, ids), 10, flush)
    ops["mask_overlaps_128x512x512_200x200"] = entry(ms, 4 * tp.numel() + 3 * 4 * 200 * 200, {
        "gvox_per_s": tp.numel() / (ms * 1e-3) / 1e9, "note": "includes the host-side lookup tables and their H2D copies"})
    # whole-volume prefilters (binarization_nuclei.py:43-44) on the BASELINE config-2 shape and on the config-3 shape
    from b200seg import prefilter
    for shape in ((59, 350, 640), SHAPE):
        vol = torch.from_numpy(rng_e.integers(0, 4096, shape).astype(np.uint16)).to(dev)
        tag = "x".join(str(s) for s in shape)
        for name, fn in (("gaussian_sigma1", lambda: prefilter.gaussian_filter(vol, 1)), ("median3", lambda: prefilter.median_filter(vol, 3))):
            ms = time_op(torch, fn, 10, flush)
            ops["%s_u16_%s" % (name, tag)] = entry(ms, 4 * vol.numel(), {"gvox_per_s": vol.numel() / (ms * 1e-3) / 1e9})
    ms = time_op(torch, lambda: prefilter.zscore_norm(vol), 10, flush)
    ops["zscore_norm_u16_128x512x512"] = entry(ms, (2 + 2 + 4) * vol.numel(), {"gvox_per_s": vol.numel() / (ms * 1e-3) / 1e9,
                                               "note": "two reads (moments, apply) + fp32 write; 5 launches"})
    prm_maps = torch.rand((14, 32, 128, 128), device=dev)
    ms = time_op(torch, lambda: prefilter.prm_to_uint8(prm_maps), 10, flush)
    ops["prm_to_uint8_14x32x128x128"] = entry(ms, (4 + 4 + 1) * prm_maps.numel(), {"note": "two reads (min/max, scale) + uint8 write; 3 launches"})
    # Mask R-CNN mask paste-back (config 2 tile: 64x200x200, 100 detections, 14^3 masks): resize + threshold + clip, one launch
    from b200seg import segm, _lib as _bl
    n_sg, M_sg, (S_sg, H_sg, W_sg) = 100, 14, (64, 200, 200)
    lo_sg = np.stack([rng_e.uniform(-5, W_sg - 30, n_sg), rng_e.uniform(-5, H_sg - 30, n_sg), rng_e.uniform(-5, S_sg - 15, n_sg)], axis=1)
    ext_sg = np.stack([rng_e.uniform(25, 60, n_sg), rng_e.uniform(25, 60, n_sg), rng_e.uniform(10, 40, n_sg)], axis=1)
    bx_sg = np.ascontiguousarray(segm.expand_boxes(np.concatenate([lo_sg, lo_sg + ext_sg], axis=1).astype(np.float32), (M_sg + 2.0) / M_sg).astype(np.int32))
    _, off_sg = segm.clipped_boxes(bx_sg, S_sg, H_sg, W_sg)
    mk_sg = torch.from_numpy(rng_e.random((n_sg, 2, M_sg, M_sg, M_sg), dtype=np.float32)).to(dev)
    ix_sg = torch.arange(n_sg, dtype=torch.int32, device=dev) * 2 + 1
    d_bx, d_off, d_tab = torch.from_numpy(bx_sg).to(dev), torch.from_numpy(off_sg).to(dev), torch.from_numpy(segm.gauss_table(M_sg)).to(dev)
    crops_sg = torch.empty(int(off_sg[-1]) + 16, dtype=torch.uint8, device=dev)
    L_sg = _bl.lib()

    def run_segm():
        _bl.check(L_sg.b200seg_segm_paste_dev(_bl.ptr(mk_sg), _bl.ptr(ix_sg), _bl.ptr(d_bx), n_sg, M_sg, _bl.ptr(d_tab), 0.5, S_sg, H_sg, W_sg,
                                              _bl.ptr(crops_sg), _bl.ptr(d_off), _bl.current_stream()), "segm_paste")
    ms = time_op(torch, run_segm, 20, flush)
    ops["segm_paste_100x14^3_64x200x200"] = entry(ms, n_sg * M_sg ** 3 * 4 + int(off_sg[-1]) + 36 * n_sg, {
        "mask_voxels_per_s": int(off_sg[-1]) / (ms * 1e-3), "dets_per_s": n_sg / (ms * 1e-3),
        "note": "fp64-issue bound by construction: 8 taps x (3 DMUL + DADD) per output voxel in scipy's order; HBM bytes are the 14^3 blocks in and one byte per voxel out"})
    # nuclei per-instance chain (config 2 shape: 59x350x640 uint16, ~40 large blobs + 10 duplicate boxes), device resident
    from b200seg import binarization_nuclei as bn_
    from b200seg.binarization import crop_offsets as crop_offsets_
    nshape = (59, 350, 640)
    ncase = synth.postproc_case(1002, shape=nshape, n_blobs=40, n_dup=10, n_false=0, sigma_xy=(10, 15), sigma_z=(4, 7))
    nboxes = bn_.nuclei_boxes(ncase["dets"], np.zeros((len(ncase["dets"]), 3), np.int64), max(nshape[1], nshape[2]), nshape[0])
    nboxes[:, 3] = np.minimum(nboxes[:, 3], nshape[2] - 1); nboxes[:, 4] = np.minimum(nboxes[:, 4], nshape[1] - 1)
    assert np.array_equal(nboxes, ncase["boxes"])
    nvol = torch.from_numpy(ncase["volume"].astype(np.uint16) * 7 + 11).to(dev)
    d_nb, d_np, d_no = torch.from_numpy(nboxes).to(dev), torch.from_numpy(ncase["prm"]).to(dev), torch.from_numpy(ncase["crop_off"]).to(dev)
    ms = time_op(torch, lambda: bn_.binarize_nuclei(nvol, d_nb, d_np, d_no), 10, flush)
    nvox = int(ncase["crop_off"][-1])
    ops["binarize_nuclei_59x350x640_u16_50inst"] = entry(ms, 3 * nvox + nvox + 2 * nvol.numel(), {
        "gvox_per_s": nvol.numel() / (ms * 1e-3) / 1e9, "crop_voxels": nvox,
        "note": "14 launches (min/max, normalise, Otsu, 2 x largest component, 2 complements, dilate, erode, paste) + output allocation; latency bound on 50 crops"})
    # evaluation helpers on a 128x512x512 label volume pair (np.unique replacement; nuclei voxel counts with 200 matched boxes)
    L_ev = _bl.lib()
    present = torch.empty(65536, dtype=torch.uint8, device=dev)
    ms = time_op(torch, lambda: _bl.check(L_ev.b200seg_label_presence_dev(_bl.ptr(tp), tp.numel(), _bl.ptr(present), _bl.current_stream()), "presence"), 20, flush)
    ops["label_presence_128x512x512"] = entry(ms, 2 * tp.numel(), {"gvox_per_s": tp.numel() / (ms * 1e-3) / 1e9})
    ev_boxes = torch.from_numpy(np.ascontiguousarray(np.array(ev_box_list, np.int32))).to(dev)
    ev_counts = torch.zeros(3, dtype=torch.int64, device=dev)
    ev_wsb = L_ev.b200seg_eval_voxel_counts_workspace_bytes(tp.numel())
    ev_ws = torch.empty(ev_wsb, dtype=torch.uint8, device=dev)
    ms = time_op(torch, lambda: _bl.check(L_ev.b200seg_eval_voxel_counts_dev(_bl.ptr(tp), _bl.ptr(tg), SHAPE[0], SHAPE[1], SHAPE[2], _bl.ptr(ev_boxes),
                                                                             int(ev_boxes.shape[0]), _bl.ptr(ev_counts), _bl.ptr(ev_ws), ev_wsb,
                                                                             _bl.current_stream()), "voxel_counts"), 20, flush)
    ops["eval_voxel_counts_128x512x512_200boxes"] = entry(ms, 4 * tp.numel() + tp.numel() // 8, {"gvox_per_s": tp.numel() / (ms * 1e-3) / 1e9,
                                                                                                "note": "bit-volume memset + box rasterisation + count pass"})
    # RPN proposal generation on the soma test tile: 14 anchors x 16x40x40, pre/post NMS top-N 1000, thresh 0.23
    from b200seg.generate_proposals_3d import GenerateProposalsOp_3d
    A_, S_, H_, W_ = 14, 16, 40, 40
    anchors = np.concatenate([np.stack([-(s / 2 - 2) * np.ones(3), (s / 2 + 1) * np.ones(3)]).reshape(1, 6) * np.array([1, 1, r, 1, 1, r])
                              for s in (8, 12, 16, 20, 24, 30, 36) for r in (1.0, 0.5)]).astype(np.float32)
    n_ = A_ * S_ * H_ * W_
    sc = torch.from_numpy(rng_e.permutation(n_).astype(np.float32).reshape(1, A_, S_, H_, W_) / np.float32(n_)).to(dev)
    dl = torch.from_numpy((rng_e.standard_normal((1, 6 * A_, S_, H_, W_)) * 0.3).astype(np.float32)).to(dev)
    info = np.array([[S_ * 4, H_ * 4, W_ * 4, 1.0]], np.float32)
    gp = GenerateProposalsOp_3d(anchors, 0.25, pre_nms_topN=1000, post_nms_topN=1000, nms_thresh=0.23)
    ms = time_op(torch, lambda: gp.forward_device(sc, dl, info), 20, flush)
    ops["generate_proposals_14x16x40x40_top1000"] = {"us": ms * 1e3, "anchors_per_s": n_ / (ms * 1e-3),
                                                     "note": "latency bound: 5 select passes + collect + rank/decode + compact + 3 NMS launches + gather; "
                                                             "score and delta maps stay on the device"}
    # BASELINE configs[1]: device-resident Mask R-CNN tile flow (proposals -> RoIAlign3D -> stand-in box head -> box_results ->
    # mask RoIAlign3D -> stand-in mask head -> segm_results) on one 64x200x200 tile, eager and as ONE CUDA graph
    from b200seg.maskrcnn_flow import TileFlow
    flow = TileFlow(tile=(64, 200, 200), C=256, dets_per_im=300, seed=1)
    A_f, (S8, H8, W8) = 35, (8, 25, 25)
    nf = A_f * S8 * H8 * W8
    f_feat = torch.from_numpy(rng_e.standard_normal((1, 256, S8, H8, W8)).astype(np.float32)).to(dev)
    f_sc = torch.from_numpy((rng_e.permutation(nf).astype(np.float32) / np.float32(nf)).reshape(1, A_f, S8, H8, W8)).to(dev)
    f_dl = torch.from_numpy((rng_e.standard_normal((1, 6 * A_f, S8, H8, W8)) * 0.2).astype(np.float32)).to(dev)
    l_f0 = b200seg.launch_count()
    f_out = flow.run(f_feat, f_sc, f_dl)
    l_f = b200seg.launch_count() - l_f0
    ms_eager = time_op(torch, lambda: flow.run(f_feat, f_sc, f_dl), 10, flush)
    fg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(fg):
        f_out = flow.run(f_feat, f_sc, f_dl)
    ms_graph = time_op(torch, fg.replay, 20, flush)
    nd_f = int(f_out["n_dets"][0])
    ops["maskrcnn_tile_flow_64x200x200"] = {
        "ms_eager": ms_eager, "ms_cuda_graph": ms_graph, "tiles_per_s": 1e3 / ms_graph, "dets": nd_f, "rois": int(f_out["n_rois"][0]),
        "dets_per_s": nd_f / (ms_graph * 1e-3), "b200seg_kernel_launches": int(l_f), "volume_59x350x640_ms": 18 * ms_graph,
        "note": "synthetic features / RPN maps, random-init stand-in heads (the PyTorch model is not the product); 18 such tiles cover a "
                "59x350x640 volume at TEST.CROP_OVLP 100 (core/test.py:86-93); checked against box_results / segm_results in the GPU tests"}
    del flow, fg, f_out
    # NMS (latency bound: report microseconds)
    for n in (50, 1000):
        d_np = synth.random_dets(rng, n, extent=(256, 256, 64))
        d = torch.from_numpy(d_np).to(dev)
        off = torch.tensor([0, n], dtype=torch.int32, device=dev)
        ms = time_op(torch, lambda: b200seg.nms_3d_batched(d, off, n, NMS_THRESH), 20, flush)
        ops["nms3d_n%d" % n] = {"us": ms * 1e3, "pairs_per_s": n * (n - 1) / 2 / (ms * 1e-3),
                                "note": "latency bound (includes the wrapper's output allocations)"}
        if ref_nms is not None:
            ops["nms3d_n%d" % n]["cpu_ref_us"] = cpu_time(lambda: ref_nms.nms_3d(d_np, np.float32(NMS_THRESH)), 5) * 1e3
            ops["nms3d_n%d" % n]["cpu_ref"] = "reference cython_nms_3d.nms_3d (oracle/_ref), 1 core"
    return ops


def make_cases(seeds):
    """Synthetic volumes are independent: build them on a thread pool (numpy releases the GIL in the heavy parts)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(seeds), os.cpu_count() or 1, 16))) as ex:
        return list(ex.map(make_case, seeds))


def make_response_maps(ids):
    from concurrent.futures import ThreadPoolExecutor
    from b200seg import synth
    f = lambda i: synth.response_map(np.random.default_rng(5000 + i), (32, 128, 128), n_peaks=60, channels=14)[0]
    with ThreadPoolExecutor(max_workers=max(1, min(len(ids), os.cpu_count() or 1, 16))) as ex:
        return np.stack(list(ex.map(f, ids)))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import b200seg
    from b200seg.peak_stimulation_3d import PeaksPlan
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peak, peak_src = hbm_peak()
    V = int(np.prod(SHAPE))
    if args.scaling == "strong":
        vol_ids = volume_split(args.volumes, world, rank)
    else:
        vol_ids = list(range(rank * args.volumes_per_rank, (rank + 1) * args.volumes_per_rank))
    vpr = len(vol_ids)
    n_global = args.volumes if args.scaling == "strong" else args.volumes_per_rank * world
    assert vpr > 0, "fewer volumes than ranks"

    cases = make_cases([args.seed_base + i for i in vol_ids])
    maps_np = make_response_maps(vol_ids)                              # [vpr,14,32,128,128] fp32
    counts = [c["dets"].shape[0] for c in cases]
    prm_np = np.concatenate([c["prm"] for c in cases])
    offs, base = [], 0
    for c in cases:
        offs.append(c["crop_off"][:-1] + base); base += int(c["crop_off"][-1])
    crop_off_np = np.concatenate(offs + [np.array([base], np.int64)])
    vols = torch.from_numpy(np.stack([c["volume"] for c in cases])).to(dev)
    dets = torch.from_numpy(np.concatenate([c["dets"] for c in cases])).to(dev)
    boxes = torch.from_numpy(np.concatenate([c["boxes"] for c in cases])).to(dev)
    prm = torch.from_numpy(prm_np).to(dev)
    crop_off = torch.from_numpy(crop_off_np).to(dev)
    maps = torch.from_numpy(maps_np).to(dev)
    pp = b200seg.SomaPostproc(vpr, SHAPE, counts, prm_np.size, device=dev)
    plan = PeaksPlan(maps.shape, dev, 3, 1, want_agg=True, cap=maps.numel() // 16)
    # the detections live inside the chain's exchange buffer: the per-step gather needs no packing
    pp.dets_in.copy_(dets)
    dets = pp.dets_in
    # exchange buffers of all ranks have different sizes under strong scaling (array_split): gather into a padded buffer
    ex_words = torch.tensor([pp.exchange.numel()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ex_words, op=dist.ReduceOp.MAX)
    ex_max = int(ex_words.item())
    gather_out = torch.zeros((world, ex_max), dtype=torch.int32, device=dev) if world > 1 else None
    # The exchange of step i runs on a side stream and overlaps the (latency-bound) NMS phase of step i+1: the
    # chain's exchange buffer is snapshotted (0.2 MB device copy) so the next step may overwrite it.
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    snap = torch.zeros(ex_max, dtype=torch.int32, device=dev) if world > 1 else None
    ev_chain, ev_gather = torch.cuda.Event(), torch.cuda.Event()
    with_peaks = [True]

    graphs = {}                                             # with_peaks -> captured CUDA graph of the rank's kernels of one step

    # The peak finder and the binarization chain of a step share no data (the peak finder reads the response maps, the
    # chain the volumes / detections / PRM crops), and neither fills the GPU on its own (occupancy 19 % in the scan,
    # barrier / latency stalls in the emit pass and in the per-instance kernels): they run as two branches of the step --
    # fork / join on events, captured into the same CUDA graph -- unless --no-overlap asks for the serial order.
    peak_stream = torch.cuda.Stream(device=dev)
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    def step_body():
        overlap = with_peaks[0] and not args.no_overlap
        if with_peaks[0]:
            if overlap:
                cur = torch.cuda.current_stream()
                ev_fork.record(cur)
                peak_stream.wait_event(ev_fork)
                with torch.cuda.stream(peak_stream):
                    plan.run(maps)                              # peak list, count, aggregation stay on the device
                    ev_join.record(peak_stream)
            else:
                plan.run(maps)
        pp.run(vols, dets, boxes, prm, crop_off, NMS_THRESH)
        if overlap:
            torch.cuda.current_stream().wait_event(ev_join)

    def step():
        g_ = graphs.get(with_peaks[0])
        if g_ is not None:
            g_.replay()                                         # one launch for the memsets + ~15 kernels of the step
        else:
            step_body()
        if world > 1:
            main = torch.cuda.current_stream()
            main.wait_event(ev_gather)                      # the previous gather has finished reading `snap`
            snap[:pp.exchange.numel()].copy_(pp.exchange)
            ev_chain.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev_chain)
                # the only exchange step: detections, visit order, survivor counts and flags of every volume
                # (one all-gather of ~0.2 MB per rank over NCCL / NVLink)
                if os.environ.get("B200SEG_BENCH_EXCHANGE", "overlap") != "none":        # diagnostic switch
                    dist.all_gather_into_tensor(gather_out.view(-1), snap)
                ev_gather.record(side)

    def finish_steps():
        if world > 1:
            torch.cuda.current_stream().wait_event(ev_gather)   # the last exchange belongs to the timed region

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_block(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(nsteps):
            step()
        finish_steps()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l_a = b200seg.launch_count()
    step_body()
    launches_per_step = b200seg.launch_count() - l_a            # kernels of this library per step (memsets not counted)
    with_peaks[0] = False
    step_body()
    with_peaks[0] = True
    torch.cuda.synchronize()
    if not args.no_graph:
        # the chain is allocation-free and never synchronises, so a step is capturable: the launch-bound part of a step
        # (15 small launches at 8 volumes per rank) collapses into one graph launch
        try:
            for wp in (True, False):
                with_peaks[0] = wp
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    step_body()
                graphs[wp] = g_
        except Exception as e:                                  # capture unsupported: stay eager, say so
            graphs.clear()
            sys.stderr.write("bench: CUDA graph capture failed (%s), running eager\n" % e)
        with_peaks[0] = True
        for _ in range(2):
            step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                     # before the barrier: spawning nvidia-smi must not skew rank 0
    # The timed region is K steps between barriers (max over ranks).  One block of K steps of this workload lasts only
    # milliseconds, so the block is repeated until >= 0.5 s have been timed; ms_per_step is the mean over all blocks.
    blocks = [timed_block(args.steps)]
    while sum(blocks) < args.min_timed_ms and len(blocks) < 200:
        blocks.append(timed_block(args.steps))
    ms_step = sum(blocks) / (len(blocks) * args.steps)
    ms_eager = None
    if graphs:                                              # the same K steps launched kernel by kernel, for the record
        saved = dict(graphs); graphs.clear()
        step()
        ms_eager = timed_block(args.steps) / args.steps
        graphs.update(saved)
    ms_serial = None
    if not args.no_overlap:                                 # the same steps with the two branches one after the other (eager)
        saved = dict(graphs); graphs.clear()
        args.no_overlap = True
        step()
        ms_serial = timed_block(args.steps) / args.steps
        args.no_overlap = False
        graphs.update(saved)
    lt = torch.tensor([launches_per_step * args.steps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    value = n_global * V / (ms_step * 1e-3) / 1e9
    # ---- NCCL exchange content check: every rank's slot of the gathered buffer must equal that rank's own exchange data
    exchange_ok = None
    if world > 1:
        torch.cuda.synchronize()
        mine = torch.zeros(ex_max, dtype=torch.int32, device=dev)
        mine[:pp.exchange.numel()].copy_(pp.exchange)
        cks = torch.stack([mine.to(torch.int64).sum(), (mine.to(torch.int64) * torch.arange(1, ex_max + 1, device=dev)).sum()])
        all_cks = [torch.zeros_like(cks) for _ in range(world)]
        dist.all_gather(all_cks, cks)
        got = gather_out.to(torch.int64)
        ar = torch.arange(1, ex_max + 1, device=dev)
        exchange_ok = all(int(got[r].sum()) == int(all_cks[r][0]) and int((got[r] * ar).sum()) == int(all_cks[r][1]) for r in range(world))
        assert bool(torch.equal(gather_out[rank], mine)), "all-gather slot of this rank differs from its own exchange buffer"
        assert exchange_ok, "all-gathered exchange buffer does not match the ranks' checksums"
    # the chain alone (without peak stimulation), same protocol, one block
    with_peaks[0] = False
    step()
    ms_nopeaks = timed_block(args.steps) / args.steps
    with_peaks[0] = True
    step()
    torch.cuda.synchronize()
    keep_counts = pp.keep_count.cpu().numpy()
    rank_order_np = pp.rank_order.cpu().numpy()
    kept_crop_bytes = 0
    for v in range(vpr):
        order = rank_order_np[pp.det_off_host[v]:pp.det_off_host[v] + keep_counts[v]]
        kept_crop_bytes += int(np.diff(cases[v]["crop_off"])[order].sum())

    # ---- per-kernel breakdown + roofline of the dominant kernel (same inputs, same launches) -------
    prof = {"nms": 0.0, "otsu": 0.0, "cc": 0.0, "paste": 0.0}
    reps = max(3, min(10, args.steps))
    pp.run_profiled(vols, dets, boxes, prm, crop_off, NMS_THRESH)
    for _ in range(reps):
        r = pp.run_profiled(vols, dets, boxes, prm, crop_off, NMS_THRESH)
        for k in prof:
            prof[k] += r[k] / reps
    # peak stimulation kernel by kernel: the C-ABI's profiling knob stops the op after its first k launches
    L = b200seg._lib.lib()
    stage_ms = []
    for k in (0, 1, 2, 3, 99):
        L.b200seg_set_option(b"peaks_stop_after", k)
        for _ in range(2):
            plan.run(maps)
        torch.cuda.synchronize()
        tt = []
        for _ in range(reps):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(); plan.run(maps); b_.record(); torch.cuda.synchronize()
            tt.append(a_.elapsed_time(b_))
        stage_ms.append(float(np.median(tt)))
    L.b200seg_set_option(b"peaks_stop_after", 99)
    prof["peaks_scan"] = max(stage_ms[1] - stage_ms[0], 0.0)
    prof["peaks_median"] = max(stage_ms[3] - stage_ms[1], 0.0)
    prof["peaks_emit"] = max(stage_ms[4] - stage_ms[3], 0.0)
    Vm = int(maps.numel())
    n_pk = int(plan.n.item())
    # one launch covers every volume of the rank's batch
    alg = {"paste": 2 * V * vpr + kept_crop_bytes,                   # label volumes written once + mask bytes read
           "otsu": 3 * kept_crop_bytes,                              # image + prm read, mask written (uint8)
           "cc": kept_crop_bytes,                                    # masks read once (cleared runs are a few % more)
           "nms": sum(28 * c + 16 * c * ((c + 63) // 64) + 8 * c for c in counts),
           "peaks_scan": 4 * Vm + Vm // 8,                           # the map read once + the 1 bit / voxel candidate mask
           "peaks_median": 4 * Vm,                                   # the second read of the map (8 B / voxel in total)
           "peaks_emit": Vm // 8 + 4 * n_pk + 40 * n_pk}             # mask read, candidate values, int64 rows written
    names = {"paste": "paste_labels_kernel", "otsu": "soma_binarize_kernel", "cc": "largest_cc_fill_kernel", "nms": "nms3d (3 kernels)",
             "peaks_scan": "peaks_scan3f_kernel", "peaks_median": "peaks_sample_kernel + peaks_interval_kernel", "peaks_emit": "peaks_finalize_kernel"}
    tot_prof = max(sum(prof.values()), 1e-9)
    kernels = {k: {"kernel": names[k], "ms_per_step": prof[k], "share": prof[k] / tot_prof,
                   "alg_bytes_per_launch": alg[k], "achieved_gbs": alg[k] / (prof[k] * 1e-3) / 1e9 if prof[k] > 0 else None}
               for k in prof}
    dom = max(prof, key=lambda k: prof[k])
    roof = {"bound": "hbm", "kernel": names[dom],
            "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": None, "peak_source": peak_src,
            "launch_ms": prof[dom], "alg_bytes_per_launch": alg[dom]}
    tr = os.path.join(ROOT, "profiles", "traffic.json")          # dram bytes per launch from the committed ncu capture
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr)).get(roof["kernel"])
        except Exception:
            pass

    if args.chain_only:
        if rank == 0:
            print(json.dumps({"metric": "postproc_gvox_per_s", "value": value, "ms_per_step": ms_step, "ms_per_step_without_peaks": ms_nopeaks,
                              "kernels": kernels, "gpu_launches": int(lt.item()), "note": "chain-only profiling run"}))
        return
    # ---- e2e: host-buffer C-ABI call, pinned memory, H2D + D2H inside the timed region --------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_in = [dict(volume=pin(c["volume"]), dets=pin(c["dets"]), boxes=pin(c["boxes"]), prm=pin(c["prm"]),
                 crop_off=pin(c["crop_off"])) for c in cases]
    h_seg = [torch.empty(SHAPE, dtype=torch.uint16).pin_memory() for _ in range(vpr)]

    e2e_cases = [dict(volume=h["volume"].numpy(), dets=h["dets"].numpy(), boxes=h["boxes"].numpy(), prm=h["prm"].numpy(),
                      crop_off=h["crop_off"].numpy()) for h in h_in]
    e2e_segs = [t.numpy() for t in h_seg]

    def e2e_step():
        # one C call for the rank's batch: uploads, kernels and downloads of consecutive volumes overlap
        return b200seg.postproc_soma_host_batch(e2e_cases, NMS_THRESH, seg_out=e2e_segs)
    from b200seg.binarization import set_host_batch_out
    e2e_steps = max(2, min(args.steps, int(np.ceil(640.0 / max(vpr, 1)))))         # about 10 passes over 64 volumes at most

    def e2e_timed(n_steps):
        barrier()
        t0_ = time.perf_counter()
        for _ in range(n_steps):
            e2e_step()
        barrier()
        dt_ = torch.tensor([time.perf_counter() - t0_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
        return n_global * V * n_steps / float(dt_.item()) / 1e9

    # (a) label buffers of unknown content: the library zero-fills every volume on the host (2 B/voxel of host memory traffic)
    set_host_batch_out(0)
    for _ in range(2):
        e2e_out = e2e_step()
    e2e_dense_fill = e2e_timed(max(2, e2e_steps // 2))
    # (b) the headline: the same pinned label buffers are reused step after step and the library is told so -- it clears
    # the voxels it wrote last time instead of whole volumes.  Same bytes in the buffers afterwards (parity-checked below
    # against the oracle on THESE buffers); the reference allocates a lazily-zero np.zeros per volume (binarization_soma.py:57).
    set_host_batch_out(2)
    for _ in range(2):
        e2e_out = e2e_step()
    l1 = b200seg.launch_count()
    e2e_val = e2e_timed(e2e_steps)
    set_host_batch_out(0)
    e2e_launches = (b200seg.launch_count() - l1) // e2e_steps
    from b200seg.binarization import host_batch_traffic
    h2d, d2h = host_batch_traffic()                  # bytes the library actually moved over the link in the last step
    h2d_dense = sum(c["volume"].nbytes + c["dets"].nbytes + c["boxes"].nbytes + c["prm"].nbytes + c["crop_off"].nbytes + 8 for c in cases)
    d2h_dense = sum(2 * V + 4 + c["dets"].shape[0] * 13 for c in cases)
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity of THIS run's workload against the oracle (and the CPU baselines, which are the same computation) ----
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_chain
    oracle = load_oracle_in_parent()
    n_check = min(vpr, args.parity_volumes)
    t_port, checked = 0.0, 0
    seg_dev = pp.seg
    surv_np = pp.survive.cpu().numpy()
    for v in range(n_check):
        t0 = time.perf_counter()
        o = oracle_chain(cases[v], NMS_THRESH)
        t_port += time.perf_counter() - t0
        lo = int(pp.det_off_host[v]); k = int(keep_counts[v])
        assert np.array_equal(rank_order_np[lo:lo + k], o["order"]), "bench parity: NMS / visit order of volume %d differs from the oracle" % v
        assert np.array_equal(surv_np[lo:lo + k].astype(bool), o["survive"]), "bench parity: survivor flags of volume %d differ" % v
        assert np.array_equal(seg_dev[v].cpu().numpy(), o["seg"]), "bench parity: label volume %d (device chain) differs from the oracle" % v
        assert np.array_equal(e2e_segs[v], o["seg"]), "bench parity: label volume %d (host-buffer call) differs from the oracle" % v
        assert np.array_equal(e2e_out[v]["rank_order"], o["order"]), "bench parity: e2e visit order of volume %d differs" % v
        checked += 1
    peaks_checked = 0
    if rank == 0 and not args.no_cpu_baseline:
        op, oagg, othr = oracle.peak_stimulation_3d(maps_np[:1], win_size=3, filter_mode="median")
        rows = plan.peaks[:n_pk].cpu().numpy()
        assert np.array_equal(rows[rows[:, 0] == 0], op), "bench parity: peak list of volume 0 differs from the oracle"
        assert np.array_equal(plan.thr[0].cpu().numpy(), othr[0]), "bench parity: median thresholds of volume 0 differ"
        peaks_checked = 1

    ops = bench_ops(torch, peak, with_cpu=(rank == 0 and not args.no_cpu_baseline)) if rank == 0 or world > 1 else {}
    if world > 1:                                           # replicas: aggregate RoIs/s over ranks
        for k in ("roialign3d_fwd_f32", "roialign3d_bwd_f32", "roialign3d_fwd_bf16", "roialign3d_bwd_bf16"):
            t = torch.tensor([ops[k]["ms"]], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ops[k]["rois_per_s_all_ranks"] = 512 * world / (float(t.item()) * 1e-3)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import refpy
        port = {"value": n_check * V / t_port / 1e9, "unit": "Gvox/s", "cores": 1,
                "note": "oracle C restatement of the chain (cython NMS, otsu_py_2d_fast, scipy label, numpy paste) on %d whole volumes of this "
                        "step, %.1f s -- the same runs that produced parity_checked" % (n_check, t_port)}
        if refpy.available():
            k_inst = 64
            r = refpy.run_soma_script(cases[0], NMS_THRESH, max_instances=k_inst)
            frac = r["n_visited"] / max(r["n_after_nms"], 1)
            tref = r["t_nms"] + r["t_loop"]
            cpu = {"value": frac * V / tref / 1e9, "unit": "Gvox/s", "cores": 1, "kind": "reference",
                   "sample": "tools/binarization_soma.py:57-104 executed as is (reference Cython nms_3d, tools/otsu.py otsu_py_2d_fast, the script's numpy "
                             "crop / paste / np.unique; skimage label -> scipy) on volume 0 of this step for the first %d of %d detections of its visit "
                             "order, %.1f s on 1 of %d host cores; value = fraction of the visit list x voxels / time; peak stimulation: see "
                             "ops.peaks3d.cpu_ref_ms" % (r["n_visited"], r["n_after_nms"], tref, os.cpu_count() or 1),
                   "port": port}
        else:
            cpu = dict(port, kind="port", sample=port["note"])
    if rank == 0:
        line = {"metric": "postproc_gvox_per_s", "value": value, "unit": "Gvox/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, world, vpr),
                "timed_blocks": len(blocks), "timed_region_s": sum(blocks) / 1e3,
                "cuda_graph": bool(graphs), "peaks_and_chain_overlap": not args.no_overlap, "ms_per_step_eager": ms_eager, "ms_per_step_serial_eager": ms_serial, "kernel_launches_per_step": int(launches_per_step),
                "value_without_peaks": n_global * V / (ms_nopeaks * 1e-3) / 1e9, "ms_per_step_without_peaks": ms_nopeaks,
                "parity_checked": checked, "peaks_parity_checked": peaks_checked, "exchange_checked": exchange_ok,
                "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
                "e2e": {"value": e2e_val, "unit": "Gvox/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "gpu_launches": int(e2e_launches), "steps": e2e_steps,
                        "h2d_bytes_if_everything_were_copied": h2d_dense, "d2h_bytes_if_dense": d2h_dense,
                        "value_dense_fill": e2e_dense_fill,
                        "label_buffers": "value: the pinned label buffers are reused every step and the call is told that they still hold "
                                         "its previous result (b200seg_set_option host_batch_out=2: only the voxels written last time are "
                                         "cleared); value_dense_fill: buffers of unknown content, every volume zero-filled on the host "
                                         "(host_batch_out=0).  Identical bytes either way; both parity-checked against the oracle",
                        "transfer": "image crops of the NMS survivors with a positive PRM voxel packed by host threads into a pinned buffer "
                                    "(DMA); their PRM crops by zero-copy gather from the caller's pinned buffer; label volume as compacted "
                                    "non-zero 64-byte lines, written by host threads with full-line non-temporal stores",
                        "note": "binarization chain through b200seg_postproc_soma_host_batch; the peak finder's input is the network's "
                                "response map, which never exists on the host in the reference flow (peak_response_mapping_3d.py:150)"},
                "gpu_launches": int(lt.item()), "kernels": kernels, "ops": ops,
                "kept_instances_per_volume": float(np.mean(keep_counts)), "peaks_per_volume": n_pk / vpr,
                "chain_roofline": {"alg_bytes_per_volume": (4 * kept_crop_bytes / vpr) + 2 * V + alg["nms"] / vpr + 8 * Vm / vpr,
                                   "frac": ((4 * kept_crop_bytes / vpr) + 2 * V + 8 * Vm / vpr) / (ms_step / vpr * 1e-3) / 1e9 / peak}}
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--volumes", type=int, default=64, help="volumes per step over ALL ranks (strong scaling, BASELINE configs[4])")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--volumes-per-rank", type=int, default=8, help="weak scaling only")
    ap.add_argument("--parity-volumes", type=int, default=8, help="volumes of this run checked against the oracle chain")
    ap.add_argument("--min-timed-ms", type=float, default=500.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the peak finder and the chain of a step one after the other instead of as two concurrent branches")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a captured CUDA graph")
    ap.add_argument("--seed-base", type=int, default=2000, help="seed of the first synthetic volume (volume i of rank r: base + r*vpr + i)")
    ap.add_argument("--chain-only", action="store_true", help="profiling aid: time only the device-resident chain")
    ap.add_argument("--ops-only", action="store_true", help="profiling aid: only the per-operator timings (RoIAlign3D, peaks, IoU, NMS)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.ops_only:
        import torch
        torch.cuda.set_device(local_rank)
        print(json.dumps({"ops": bench_ops(torch, hbm_peak()[0]), "note": "ops-only profiling run"}))
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
