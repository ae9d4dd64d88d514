"""Whole-volume prefilters and normalisations behind the names the reference scripts call.

gaussian_filter(img, sigma=1)     scipy.ndimage.gaussian_filter on the raw integer volume (tools/binarization_nuclei.py:43)
median_filter(img, size=3)        scipy.ndimage.median_filter (tools/binarization_nuclei.py:44)
prefilter_nuclei(img)             the two chained, as the nuclei script applies them
zscore_norm(im)                   (im - mean(im[im>0])) / std(im[im>0])   (tools/infer_simple.py:180-183, lib/utils/blob.py:180-184)
prm_to_uint8(prm)                 per-channel min-max stretch to uint8     (tools/infer_simple.py:233-238)

numpy in -> numpy out, cuda tensor in -> cuda tensor out.  Every call runs the CUDA kernels of csrc/prefilter.cu."""
import ctypes as C

import numpy as np

from . import _lib

_ELEM = {"uint8": 1, "uint16": 2}


def gaussian_kernel1d(sigma, truncate=4.0):
    """First radius+1 taps (centre last) of the order-0 kernel, computed like scipy.ndimage._filters._gaussian_kernel1d
    (numpy exp, numpy sum, true division) so that the doubles are the ones scipy would use."""
    sd = float(sigma)
    radius = int(float(truncate) * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[:radius + 1], dtype=np.float64), radius


def _to_dev(img, what):
    import torch
    if hasattr(img, "data_ptr"):
        t = img if img.is_cuda else img.cuda()
        return t.contiguous(), True
    a = np.ascontiguousarray(img)
    return torch.from_numpy(a).cuda(), False


def _check_int_volume(t, what):
    name = str(t.dtype).replace("torch.", "")
    if name not in _ELEM:
        raise TypeError("%s: integer volumes only (uint8 / uint16), got %s" % (what, name))
    if t.dim() != 3:
        raise ValueError("%s: expected a [S,H,W] volume, got shape %s" % (what, tuple(t.shape)))
    return _ELEM[name]


def gaussian_filter(img, sigma=1, truncate=4.0):
    import torch
    t, was_tensor = _to_dev(img, "gaussian_filter")
    eb = _check_int_volume(t, "gaussian_filter")
    w, radius = gaussian_kernel1d(sigma, truncate)
    out = torch.empty_like(t)
    if radius < 1:
        out.copy_(t)
    else:
        S, H, W = t.shape
        _lib.check(_lib.lib().b200seg_gaussian3d_dev(_lib.ptr(t), _lib.ptr(out), eb, S, H, W, w.ctypes.data_as(C.c_void_p), radius,
                                                     _lib.current_stream()), "gaussian3d")
    return out if was_tensor else out.cpu().numpy()


def median_filter(img, size=3):
    import torch
    if size != 3:
        raise NotImplementedError("median_filter: only the 3x3x3 window the reference uses (size=3)")
    t, was_tensor = _to_dev(img, "median_filter")
    eb = _check_int_volume(t, "median_filter")
    out = torch.empty_like(t)
    S, H, W = t.shape
    _lib.check(_lib.lib().b200seg_median3d_dev(_lib.ptr(t), _lib.ptr(out), eb, S, H, W, _lib.current_stream()), "median3d")
    return out if was_tensor else out.cpu().numpy()


def prefilter_nuclei(img, sigma=1):
    """binarization_nuclei.py:43-44: gaussian_filter(sigma=1) then median_filter(size=3); one upload, one download."""
    import torch
    t, was_tensor = _to_dev(img, "prefilter_nuclei")
    out = median_filter(gaussian_filter(t, sigma), 3)
    return out if was_tensor else out.cpu().numpy()


def zscore_norm(im, return_stats=False):
    """float32 (im - mean)/std over the non-zero voxels; uint8 / uint16 / float32 input of any shape."""
    import torch
    t, was_tensor = _to_dev(im, "zscore_norm")
    name = str(t.dtype).replace("torch.", "")
    eb = {"uint8": 1, "uint16": 2, "float32": 4}.get(name)
    if eb is None:
        raise TypeError("zscore_norm: uint8 / uint16 / float32 input, got %s" % name)
    L = _lib.lib()
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    stats = torch.empty(3, dtype=torch.float64, device=t.device)
    ws_bytes = L.b200seg_zscore_workspace_bytes()
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
    _lib.check(L.b200seg_zscore_norm_dev(_lib.ptr(t), eb, _lib.ptr(out), t.numel(), _lib.ptr(stats), _lib.ptr(ws), ws_bytes,
                                         _lib.current_stream()), "zscore_norm")
    res = out if was_tensor else out.cpu().numpy()
    if return_stats:
        s = stats.cpu().numpy()
        return res, float(s[0]), float(s[1])
    return res


def prm_to_uint8(prm):
    """prm [ch, z, y, x] float32 -> uint8, each channel stretched by its own min / max (infer_simple.py:233-238)."""
    import torch
    t, was_tensor = _to_dev(prm, "prm_to_uint8")
    if t.dtype != torch.float32:
        raise TypeError("prm_to_uint8: float32 maps, got %s" % t.dtype)
    if t.dim() < 2:
        raise ValueError("prm_to_uint8: expected [ch, ...] maps")
    L = _lib.lib()
    n_maps = t.shape[0]
    per_map = t.numel() // max(n_maps, 1)
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    ws_bytes = L.b200seg_prm_to_u8_workspace_bytes(n_maps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
    _lib.check(L.b200seg_prm_to_u8_dev(_lib.ptr(t), _lib.ptr(out), n_maps, per_map, _lib.ptr(ws), ws_bytes, _lib.current_stream()), "prm_to_u8")
    return out if was_tensor else out.cpu().numpy()
