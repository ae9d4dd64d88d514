"""Per-volume post-processing of tools/binarization_soma.py:57-104 on the GPU.

`soma_binarize`, `paste_labels` expose the two steps; `SomaPostproc` runs the whole chain
(3D NMS -> visit order -> per-instance crop/normalise/2D-Otsu -> label paste-back) for a batch of
volumes without leaving the device, and `postproc_soma_host` is the host-buffer (numpy) call a
drop-in binarization script makes."""
import ctypes as C

import numpy as np

from . import _lib


def dets_to_boxes(dets, shape):
    """int()-truncated, volume-clipped boxes (binarization_soma.py:78; numpy slicing clips the end)."""
    S, H, W = shape
    b = np.asarray(dets)[:, :6].astype(np.int64)          # astype(int): truncation toward zero
    b[:, 0] = np.clip(b[:, 0], 0, W - 1); b[:, 3] = np.clip(b[:, 3], 0, W - 1)
    b[:, 1] = np.clip(b[:, 1], 0, H - 1); b[:, 4] = np.clip(b[:, 4], 0, H - 1)
    b[:, 2] = np.clip(b[:, 2], 0, S - 1); b[:, 5] = np.clip(b[:, 5], 0, S - 1)
    b[:, 3] = np.maximum(b[:, 3], b[:, 0]); b[:, 4] = np.maximum(b[:, 4], b[:, 1]); b[:, 5] = np.maximum(b[:, 5], b[:, 2])
    return b.astype(np.int32)


def crop_offsets(boxes):
    b = np.asarray(boxes, dtype=np.int64)
    sizes = (b[:, 3] - b[:, 0] + 1) * (b[:, 4] - b[:, 1] + 1) * (b[:, 5] - b[:, 2] + 1)
    off = np.zeros(b.shape[0] + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    return off


def soma_binarize(volume, boxes, prm, crop_off, order=None, n_valid=None):
    """Device op (torch cuda tensors): volume uint8 [S,H,W], boxes int32 [n,6], prm uint8 packed crops,
    crop_off int64 [n+1].  Returns (mask uint8 packed, b_max int32 [n], status int32 [n])."""
    import torch
    S, H, W = volume.shape
    n = boxes.shape[0]
    dev = volume.device
    mask = torch.zeros(max(prm.numel(), 1), dtype=torch.uint8, device=dev)
    b_max = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    status = torch.full((max(n, 1),), -1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().b200seg_soma_binarize_dev(
        _lib.ptr(volume), 1, S, H, W, None, n, _lib.ptr(boxes), _lib.ptr(prm), _lib.ptr(crop_off), _lib.ptr(order),
        _lib.ptr(n_valid), _lib.ptr(mask), _lib.ptr(b_max), _lib.ptr(status), _lib.current_stream()), "soma_binarize")
    return mask, b_max[:n], status[:n]


def largest_cc(masks, crop_off, boxes, status=None, order=None, n_valid=None):
    """Device op: keep the largest 26-connected component of every packed instance mask, in place
    (binarization_soma.py:97-99).  masks uint8 packed cuda, crop_off int64 [n+1], boxes int32 [n,6].
    Returns status int32 [n] (0 ok, 5 = no foreground, 6 = degenerate crop)."""
    import torch
    n = boxes.shape[0]
    L = _lib.lib()
    if status is None:
        status = torch.zeros(max(n, 1), dtype=torch.int32, device=masks.device)
    total = int(masks.numel())
    ws_bytes = L.b200seg_largest_cc_workspace_bytes(total)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=masks.device)
    _lib.check(L.b200seg_largest_cc_dev(_lib.ptr(masks), _lib.ptr(crop_off), total, 1, None, n, _lib.ptr(boxes),
                                        _lib.ptr(order), _lib.ptr(n_valid), _lib.ptr(status), _lib.ptr(ws), ws_bytes,
                                        _lib.current_stream()), "largest_cc")
    return status[:n]


def paste_labels(seg, boxes, ids, masks, mask_off, order=None, n_valid=None):
    """Device op: seg uint16 [S,H,W] is (over)written once; returns survive uint8 [n] by visit rank."""
    import torch
    S, H, W = seg.shape
    n = boxes.shape[0]
    survive = torch.zeros(max(n, 1), dtype=torch.uint8, device=seg.device)
    L = _lib.lib()
    ws_bytes = L.b200seg_paste_labels_workspace_bytes(1, S, H, W, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=seg.device)
    _lib.check(L.b200seg_paste_labels_dev(
        _lib.ptr(seg), 1, S, H, W, None, n, _lib.ptr(boxes), _lib.ptr(ids), _lib.ptr(masks), _lib.ptr(mask_off),
        _lib.ptr(order), _lib.ptr(n_valid), _lib.ptr(survive), _lib.ptr(ws), ws_bytes, _lib.current_stream()), "paste_labels")
    return survive[:n]


class SomaPostproc(object):
    """Device-resident chain for a batch of equally shaped volumes (buffers allocated once)."""

    def __init__(self, n_volumes, shape, det_counts, prm_bytes, device="cuda", keep_largest_cc=True):
        import torch
        self.torch = torch
        self.nv = int(n_volumes)
        self.keep_largest_cc = bool(keep_largest_cc)        # binarization_soma.py:97-99 (reference semantics)
        self.prm_bytes = int(prm_bytes)
        self.S, self.H, self.W = [int(v) for v in shape]
        self.det_off_host = np.zeros(self.nv + 1, dtype=np.int32)
        self.det_off_host[1:] = np.cumsum(det_counts)
        total = int(self.det_off_host[-1])
        self.total = total
        self.n_max = int(max(det_counts)) if len(det_counts) else 0
        dev = torch.device(device)
        self.det_off_dev = torch.from_numpy(self.det_off_host).to(dev)
        t = max(total, 1)
        self.seg = torch.empty((self.nv, self.S, self.H, self.W), dtype=torch.uint16, device=dev)
        self.keep = torch.empty(t, dtype=torch.int64, device=dev)
        self.masks = torch.empty(max(int(prm_bytes), 1) + 16, dtype=torch.uint8, device=dev)
        self.b_max = torch.zeros(t, dtype=torch.int32, device=dev)
        self.status = torch.zeros(t, dtype=torch.int32, device=dev)
        # Everything another rank needs to know about this rank's volumes lives in ONE flat buffer, so the
        # exchange step of a multi-GPU run is a single all_gather with no packing kernels:
        #   [dets total x 7 f32 | rank_order total i32 | keep_count nv i32 | survive total u8 (padded to 4)]
        # `dets_in` is where the caller should write (or keep) the detections to avoid a copy.
        words = 7 * t + t + max(self.nv, 1) + (t + 3) // 4
        self.exchange = torch.zeros(words, dtype=torch.int32, device=dev)
        self.dets_in = self.exchange[:7 * t].view(torch.float32).view(t, 7)
        self.rank_order = self.exchange[7 * t:8 * t]
        self.keep_count = self.exchange[8 * t:8 * t + max(self.nv, 1)]
        self.survive = self.exchange[8 * t + max(self.nv, 1):].view(torch.uint8)[:t]
        self.ws_bytes = _lib.lib().b200seg_postproc_soma_workspace_bytes(
            self.nv, self.n_max, self.S, self.H, self.W, self.prm_bytes if self.keep_largest_cc else 0)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def launches_per_call(self):
        """Kernels launched by one run(): 3 NMS + iota + binarize (+ largest CC) + paste bin + paste (all volumes per launch)."""
        return (3 + 1 + 1 + 1 + int(self.keep_largest_cc) if self.n_max > 0 else 0) + 1

    def run(self, volumes, dets, boxes, prm, crop_off, nms_thresh):
        """All arguments are cuda tensors (volumes uint8 [nv,S,H,W], dets f32 [total,7], boxes int32
        [total,6], prm uint8 packed, crop_off int64 [total+1]).  Enqueues on the current stream."""
        _lib.check(_lib.lib().b200seg_postproc_soma_dev(
            _lib.ptr(volumes), self.nv, self.S, self.H, self.W, _lib.ptr(dets), _lib.ptr(self.det_off_dev),
            _lib.ptr(self.det_off_host), _lib.ptr(boxes), _lib.ptr(prm), _lib.ptr(crop_off), self.prm_bytes,
            float(np.float32(nms_thresh)), int(self.keep_largest_cc), _lib.ptr(self.seg), _lib.ptr(self.keep), _lib.ptr(self.keep_count),
            _lib.ptr(self.rank_order), _lib.ptr(self.masks), _lib.ptr(self.b_max), _lib.ptr(self.status),
            _lib.ptr(self.survive), _lib.ptr(self.ws), self.ws_bytes, _lib.current_stream()), "postproc_soma_dev")
        return self.seg


def _run_profiled(self, volumes, dets, boxes, prm, crop_off, nms_thresh):
    """Same launch sequence as run(), issued step by step from Python with CUDA events around each
    kernel class.  Returns {"nms": ms, "otsu": ms, "paste": ms, "n_paste": launches, ...} (bench.py)."""
    torch = self.torch
    L = _lib.lib()
    st = _lib.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    spans = {"nms": [], "otsu": [], "cc": [], "paste": []}
    if not hasattr(self, "_ids"):
        self._ids = torch.arange(1, max(self.n_max, 1) + 1, dtype=torch.int32, device=self.seg.device).to(torch.uint16)
        self._nms_ws = torch.empty(L.b200seg_nms3d_workspace_bytes(self.nv, self.n_max), dtype=torch.uint8, device=self.seg.device)
        self._paste_ws = torch.empty(L.b200seg_paste_labels_workspace_bytes(self.nv, self.S, self.H, self.W, self.n_max), dtype=torch.uint8,
                                     device=self.seg.device)
        self._cc_ws = torch.empty(L.b200seg_largest_cc_workspace_bytes(self.prm_bytes), dtype=torch.uint8, device=self.seg.device)
    a, b = ev(), ev()
    a.record()
    _lib.check(L.b200seg_nms3d_dev(_lib.ptr(dets), _lib.ptr(self.det_off_dev), self.nv, self.n_max, float(np.float32(nms_thresh)), 0,
                                   _lib.ptr(self.keep), _lib.ptr(self.keep_count), _lib.ptr(self.rank_order),
                                   _lib.ptr(self._nms_ws), self._nms_ws.numel(), st), "nms3d_dev")
    b.record()
    spans["nms"].append((a, b))
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    if self.n_max > 0:
        self.status.fill_(-1)
        self.survive.zero_()
        self.b_max.zero_()
        _lib.check(L.b200seg_soma_binarize_dev(
            _lib.ptr(volumes), self.nv, self.S, self.H, self.W, _lib.ptr(self.det_off_dev), self.n_max, _lib.ptr(boxes),
            _lib.ptr(prm), _lib.ptr(crop_off), _lib.ptr(self.rank_order), _lib.ptr(self.keep_count), _lib.ptr(self.masks),
            _lib.ptr(self.b_max), _lib.ptr(self.status), st), "soma_binarize")
    b.record()
    if self.n_max > 0 and self.keep_largest_cc:
        _lib.check(L.b200seg_largest_cc_dev(
            _lib.ptr(self.masks), _lib.ptr(crop_off), self.prm_bytes, self.nv, _lib.ptr(self.det_off_dev), self.n_max, _lib.ptr(boxes),
            _lib.ptr(self.rank_order), _lib.ptr(self.keep_count), _lib.ptr(self.status), _lib.ptr(self._cc_ws), self._cc_ws.numel(), st),
            "largest_cc")
    c.record()
    _lib.check(L.b200seg_paste_labels_dev(
        _lib.ptr(self.seg), self.nv, self.S, self.H, self.W, _lib.ptr(self.det_off_dev), self.n_max, _lib.ptr(boxes),
        _lib.ptr(self._ids), _lib.ptr(self.masks), _lib.ptr(crop_off), _lib.ptr(self.rank_order), _lib.ptr(self.keep_count),
        _lib.ptr(self.survive), _lib.ptr(self._paste_ws), self._paste_ws.numel(), st), "paste_labels")
    d.record()
    spans["otsu"].append((a, b)); spans["cc"].append((b, c)); spans["paste"].append((c, d))
    torch.cuda.synchronize()
    out = {k: sum(x.elapsed_time(y) for x, y in v) for k, v in spans.items()}
    out["n_paste"] = 1
    out["n_otsu"] = 1
    return out


SomaPostproc.run_profiled = _run_profiled


def postproc_soma_host(volume, dets, boxes, prm, crop_off, nms_thresh, seg_out=None, keep_largest_cc=True):
    """numpy in / numpy out for ONE volume (H2D and D2H happen inside the C call).
    Returns dict(seg uint16 [S,H,W], n_keep, rank_order, b_max, status, survive, scores [[id, score]])."""
    volume = np.ascontiguousarray(volume, dtype=np.uint8)
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    boxes = np.ascontiguousarray(boxes, dtype=np.int32)
    prm = np.ascontiguousarray(prm, dtype=np.uint8)
    crop_off = np.ascontiguousarray(crop_off, dtype=np.int64)
    S, H, W = volume.shape
    n = dets.shape[0]
    seg = seg_out if seg_out is not None else np.empty((S, H, W), dtype=np.uint16)
    nn = max(n, 1)
    rank = np.zeros(nn, dtype=np.int32)
    b_max = np.zeros(nn, dtype=np.int32)
    status = np.zeros(nn, dtype=np.int32)
    survive = np.zeros(nn, dtype=np.uint8)
    cnt = C.c_int(0)
    _lib.check(_lib.lib().b200seg_postproc_soma_host(
        _lib.ptr(volume), S, H, W, _lib.ptr(dets), n, _lib.ptr(boxes), _lib.ptr(prm), _lib.ptr(crop_off),
        float(np.float32(nms_thresh)), int(bool(keep_largest_cc)), _lib.ptr(seg), C.byref(cnt), _lib.ptr(rank), _lib.ptr(b_max),
        _lib.ptr(status), _lib.ptr(survive)), "postproc_soma_host")
    k = cnt.value
    order = rank[:k]
    alive = survive[:k].astype(bool)
    scores = np.stack([np.arange(1, k + 1, dtype=np.float32)[alive], dets[order, 6][alive]], axis=1) if k else \
        np.zeros((0, 2), dtype=np.float32)
    return dict(seg=seg, n_keep=k, rank_order=order.copy(), b_max=b_max[:n], status=status[:n],
                survive=alive, scores=scores.astype(np.float32))


def host_batch_traffic():
    """(h2d_bytes, d2h_bytes) the last successful postproc_soma_host_batch call moved over the link."""
    a, b = C.c_ulonglong(0), C.c_ulonglong(0)
    _lib.lib().b200seg_postproc_soma_host_batch_traffic(C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def set_host_batch_out(state):
    """What the `seg_out` buffers handed to postproc_soma_host_batch hold on entry (results are identical either way):
    0 = anything (the library zero-fills them, the default), 1 = zeros (fresh np.zeros buffers, as the reference script
    allocates per volume), 2 = what the same call wrote into the same buffers last time (only those voxels are cleared)."""
    _lib.check(_lib.lib().b200seg_set_option(b"host_batch_out", int(state)), "set_option")


def set_host_batch_mode(mode):
    """bit 0: compacted label download, bit 1: zero-copy gather of the surviving PRM crops, bit 2: the image
    crops of the NMS survivors travel packed by host threads, the raw volume is never copied (default 7); bit 3 (8): their PRM
    crops travel in the same packed buffer instead of the zero-copy gather; bits 4 / 5 (16 / 32): the chain is launched for
    groups of 2 / 4 volumes; bit 6 / bit 8 (64 / 256): pinned label buffers get their lines staged + scattered by host threads /
    written in place by the GPU (default: by the number of host cores per rank)."""
    _lib.check(_lib.lib().b200seg_set_option(b"host_batch_mode", int(mode)), "set_option")


def postproc_soma_host_batch(cases, nms_thresh, seg_out=None, keep_largest_cc=True):
    """numpy in / numpy out for a BATCH of equally shaped volumes (one C call; uploads, kernels and downloads of
    consecutive volumes overlap on three streams -- pass pinned arrays for real overlap).
    cases: list of dict(volume, dets, boxes, prm, crop_off).  Returns a list of dicts like postproc_soma_host."""
    nv = len(cases)
    if nv == 0:
        return []
    vols = [np.ascontiguousarray(c["volume"], dtype=np.uint8) for c in cases]
    dets = [np.ascontiguousarray(c["dets"], dtype=np.float32) for c in cases]
    boxes = [np.ascontiguousarray(c["boxes"], dtype=np.int32) for c in cases]
    prm = [np.ascontiguousarray(c["prm"], dtype=np.uint8) for c in cases]
    coff = [np.ascontiguousarray(c["crop_off"], dtype=np.int64) for c in cases]
    S, H, W = vols[0].shape
    for v in vols:
        if v.shape != (S, H, W):
            raise ValueError("all volumes of a batch must have the same shape")
    n = np.array([d.shape[0] for d in dets], dtype=np.int32)
    segs = list(seg_out) if seg_out is not None else [np.empty((S, H, W), dtype=np.uint16) for _ in range(nv)]
    rank = [np.zeros(max(int(k), 1), dtype=np.int32) for k in n]
    b_max = [np.zeros(max(int(k), 1), dtype=np.int32) for k in n]
    status = [np.zeros(max(int(k), 1), dtype=np.int32) for k in n]
    survive = [np.zeros(max(int(k), 1), dtype=np.uint8) for k in n]
    n_keep = np.zeros(nv, dtype=np.int32)

    def parr(arrs):
        return (C.c_void_p * nv)(*[a.ctypes.data for a in arrs])
    keepalive = [parr(a) for a in (vols, dets, boxes, prm, coff, segs, rank, b_max, status, survive)]
    pv, pd, pb, pp, pc, ps, pr, pm, pst, psv = keepalive
    _lib.check(_lib.lib().b200seg_postproc_soma_host_batch(
        nv, S, H, W, pv, pd, _lib.ptr(n), pb, pp, pc, float(np.float32(nms_thresh)), int(bool(keep_largest_cc)), ps,
        _lib.ptr(n_keep), pr, pm, pst, psv),
        "postproc_soma_host_batch")
    out = []
    for v in range(nv):
        k = int(n_keep[v]); nd = int(n[v])
        order = rank[v][:k]
        alive = survive[v][:k].astype(bool)
        scores = np.stack([np.arange(1, k + 1, dtype=np.float32)[alive], dets[v][order, 6][alive]], axis=1) if k else \
            np.zeros((0, 2), dtype=np.float32)
        out.append(dict(seg=segs[v], n_keep=k, rank_order=order.copy(), b_max=b_max[v][:nd], status=status[v][:nd],
                        survive=alive, scores=scores.astype(np.float32)))
    return out
