"""Run-length codec of binary 3D masks behind the reference names
(lib/utils/mask_3d.py:15-71, lib/utils/cython_mask_3d.pyx:19-77).

binary_mask_to_rle(mask) -> {'counts': [int, ...], 'size': [S, H, W]}     (Fortran-order runs, zeros first)
rle_to_binary_mask(rle)  -> uint8 array of shape `size`
numpy arrays or torch CUDA tensors in; the work runs in the CUDA kernels of csrc/rle3d.cu."""
import numpy as np

from . import _lib


def _encode_device(mask_t, cap):
    import torch
    L = _lib.lib()
    S, H, W = mask_t.shape
    dev = mask_t.device
    counts = torch.empty(cap, dtype=torch.int64, device=dev)
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = L.b200seg_rle3d_workspace_bytes(S, H, W, cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.b200seg_rle3d_encode_dev(_lib.ptr(mask_t), S, H, W, _lib.ptr(counts), cap, _lib.ptr(n), _lib.ptr(ws),
                                          ws_bytes, _lib.current_stream()), "rle3d_encode")
    return counts, int(n.item())


def binary_mask_to_rle(binary_mask, cap=1 << 16):
    import torch
    # any nonzero value is foreground (the reference tests np.where(mask), mask_3d.py:34-40): label values that are
    # multiples of 256 or float masks in (0, 1) must not vanish in a uint8 cast
    if hasattr(binary_mask, "data_ptr"):
        m = (binary_mask != 0)
    else:
        m = torch.from_numpy(np.ascontiguousarray(np.asarray(binary_mask) != 0))
    if m.dim() not in (2, 3):
        raise ValueError("Buffer has wrong number of dimensions (expected 2 or 3, got %d)" % m.dim())
    size = [int(s) for s in m.shape]
    if m.dim() == 2:
        # mask_3d.py documents 2-D or 3-D masks: runs follow the Fortran order of the array, which for [H,W] is the
        # Fortran order of the 3-D array [H,W,1]
        m = m.unsqueeze(2)
    m = m.to(device="cuda", dtype=torch.uint8).contiguous()
    counts, n = _encode_device(m, int(cap))
    if n > cap:                                               # more runs than the first guess: exact second pass
        counts, n = _encode_device(m, n)
    return {"counts": [int(v) for v in counts[:n].cpu().numpy()], "size": size}


def rle_to_binary_mask(rle):
    import torch
    counts = np.asarray(rle["counts"], dtype=np.int64)
    size = [int(s) for s in rle["size"]]
    if len(size) not in (2, 3):
        raise ValueError("rle_to_binary_mask expects a 2D or 3D size")
    assert int(counts.sum()) == int(np.prod(size))            # mask_3d.py:53
    L = _lib.lib()
    S, H, W = size if len(size) == 3 else (size[0], size[1], 1)
    dev = torch.device("cuda", torch.cuda.current_device())
    c = torch.from_numpy(counts).to(dev)
    mask = torch.empty((S, H, W), dtype=torch.uint8, device=dev)
    ws_bytes = L.b200seg_rle3d_workspace_bytes(S, H, W, max(int(counts.size), 1))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.b200seg_rle3d_decode_dev(_lib.ptr(c), int(counts.size), _lib.ptr(mask), S, H, W, None, _lib.ptr(ws), ws_bytes,
                                          _lib.current_stream()), "rle3d_decode")
    return mask.cpu().numpy().reshape(size)
