"""2D-Otsu binarization behind the reference name (tools/otsu.py:199-284).

otsu_py_2d_fast(image, prm, b_range=None) -> (uint8 mask {0,255}, k_max=-1, b_max)
numpy integer crops in, numpy out; the work runs in the batched CUDA kernel (one crop here)."""
import ctypes as C

import numpy as np

from . import _lib


def otsu_py_2d_fast(image, prm, b_range=None):
    if b_range is not None:
        raise NotImplementedError("b_range is never passed by the reference call sites "
                                  "(binarization_soma.py:94, binarization_nuclei.py:124)")
    image = np.asarray(image)
    prm = np.asarray(prm)
    if image.shape != prm.shape:
        raise ValueError("image and prm must have the same shape")
    for a in (image, prm):
        if a.size and (a.min() < 0 or a.max() > 65535):
            raise ValueError("otsu_py_2d_fast expects integer levels in [0, 65535]")
    img = np.ascontiguousarray(image, dtype=np.uint16).ravel()
    pr = np.ascontiguousarray(prm, dtype=np.uint16).ravel()
    mask = np.empty(img.size, dtype=np.uint8)
    b = C.c_int(0)
    code = _lib.lib().b200seg_otsu2d_host(_lib.ptr(img), _lib.ptr(pr), img.size, _lib.ptr(mask), C.byref(b))
    if code == 1:
        # the reference falls off the scan without ever assigning k_max (otsu.py:277)
        raise UnboundLocalError("local variable 'k_max' referenced before assignment")
    _lib.check(code, "otsu2d_host")
    return mask.reshape(image.shape), -1, int(b.value)


def otsu_py_2d(image, prm, nbins=256):
    """tools/otsu.py:118-197, the O(G^3) twin of otsu_py_2d_fast (same criterion, same line family);
    imported by binarization_nuclei.py:12 but never called.  Served by the same kernel."""
    return otsu_py_2d_fast(image, prm)


def _never_called(name, where):
    def f(*a, **k):
        raise NotImplementedError("%s (%s) is imported but never called by the reference scripts; "
                                  "it is outside the hot path and has no CUDA implementation" % (name, where))
    f.__name__ = name
    return f


otsu_py = _never_called("otsu_py", "tools/otsu.py:18-64")      # 1-D Otsu variants: import-compat names only
otsu_mat = _never_called("otsu_mat", "tools/otsu.py:66-116")


def otsu_2d_batch(image, prm, crop_off, want_hist=False):
    """Device batched entry: image/prm uint16 cuda [total], crop_off int64 cuda [n+1].
    Returns dict(mask uint8 [total], b_max, g_info [n,4], status[, hist list])."""
    import torch
    L = _lib.lib()
    n = crop_off.numel() - 1
    dev = image.device
    mask = torch.empty(image.numel(), dtype=torch.uint8, device=dev)
    b_max = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    g_info = torch.zeros((max(n, 1), 4), dtype=torch.int32, device=dev)
    status = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    hist = hist_off = None
    if want_hist:
        off = crop_off.cpu().numpy()
        imgs = image.cpu().numpy()
        G = [int(imgs[off[i]:off[i + 1]].max()) - int(imgs[off[i]:off[i + 1]].min()) + 1 if off[i + 1] > off[i] else 0
             for i in range(n)]
        ho = np.zeros(n + 1, dtype=np.int64)
        ho[1:] = np.cumsum([g * g for g in G])
        hist = torch.zeros(max(int(ho[-1]), 1), dtype=torch.int32, device=dev)
        hist_off = torch.from_numpy(ho).to(dev)
    _lib.check(L.b200seg_otsu2d_dev(_lib.ptr(image), _lib.ptr(prm), _lib.ptr(crop_off), n, _lib.ptr(mask),
                                    _lib.ptr(b_max), _lib.ptr(g_info), _lib.ptr(status), _lib.ptr(hist),
                                    _lib.ptr(hist_off), _lib.current_stream()), "otsu2d_dev")
    out = dict(mask=mask, b_max=b_max[:n], g_info=g_info[:n], status=status[:n])
    if want_hist:
        hs = hist.cpu().numpy()
        out["hist"] = [hs[ho[i]:ho[i + 1]].reshape(G[i], G[i]) for i in range(n)]
    return out
