"""PRM peak stimulation behind the reference name (lib/prm/peak_stimulation_3d.py:51-52).

peak_stimulation_3d(input, return_aggregation=True, win_size=3, peak_filter=None)
  -> (peak_list int64 [Npk,5], aggregation [B,A])  or  peak_list
`peak_filter` may be the string "median" (fused exact lower-median threshold, the reference's
default `_median_filter`, peak_response_mapping_3d.py:45-49), a number, or any callable returning a
broadcastable threshold ([B,A,1,1,1] or scalar) like the reference's filters: the callable is
evaluated with torch and only the resulting per-map thresholds enter the fused kernel."""
import torch
from torch.autograd import Function

from . import _lib


def median_filter(input):
    """Marker for the fused median path; calling it gives the reference result (torch.median)."""
    b, c, s, h, w = input.size()
    thr, _ = torch.median(input.view(b, c, s * h * w), dim=2)
    return thr.contiguous().view(b, c, 1, 1, 1)


def set_median_mode(mode):
    """Exact-median strategy of the fused "median" filter: 0 automatic (sampled interval for maps of 2^16..2^21 elements,
    level-1 histogram otherwise), 1 histogram path only, 2 sampled interval forced to miss (exercises the slow exact
    selection).  Results are identical in every mode."""
    _lib.check(_lib.lib().b200seg_set_option(b"peaks_median_mode", int(mode)), "set_option")


def median_fallback_count():
    """Maps whose sampled interval missed the median since the library was loaded (they took the slow exact selection)."""
    import ctypes as C
    n = C.c_ulonglong(0)
    _lib.check(_lib.lib().b200seg_peaks3d_fallback_count(C.byref(n)), "peaks3d_fallback_count")
    return int(n.value)


class PeaksPlan(object):
    """Pre-allocated, allocation-free and sync-free form of the op for a fixed input shape (what the batched chain and
    bench.py use): `run(input)` enqueues one memset + four kernels on the current stream and returns device buffers
    (peaks [cap,5] int64, n int32[1] = total number of peaks, agg [B,A] or None, thr [B,A]); nothing is read back."""

    def __init__(self, shape, device, win_size=3, filter_mode=0, want_agg=True, cap=None):
        assert win_size % 2 == 1, 'Window size for peak finding must be odd.'
        B, A, S, H, W = [int(v) for v in shape]
        self.shape, self.win, self.mode = (B, A, S, H, W), int(win_size), int(filter_mode)
        L = _lib.lib()
        V = S * H * W
        if cap is None:
            cap = min(B * A * V, max(1 << 16, (B * A * V) // 8))
        self.cap = int(cap)
        self.peaks = torch.empty((max(self.cap, 1), 5), dtype=torch.int64, device=device)
        self.n = torch.zeros(1, dtype=torch.int32, device=device)
        self.agg = torch.empty((B, A), dtype=torch.float32, device=device) if want_agg else None
        self.thr = torch.empty((B, A), dtype=torch.float32, device=device)
        self.ws_bytes = L.b200seg_peaks3d_workspace_bytes(B, A, S, H, W)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)

    def run(self, input, thr_in=None):
        B, A, S, H, W = self.shape
        assert tuple(input.shape) == self.shape and input.is_cuda and input.dtype == torch.float32 and input.is_contiguous()
        _lib.check(_lib.lib().b200seg_peaks3d_dev(_lib.ptr(input), B, A, S, H, W, self.win, self.mode, _lib.ptr(thr_in),
                                                  _lib.ptr(self.peaks), self.cap, _lib.ptr(self.n), _lib.ptr(self.agg),
                                                  _lib.ptr(self.thr), _lib.ptr(self.ws), self.ws_bytes, _lib.current_stream()),
                   "peaks3d")
        return self.peaks, self.n, self.agg, self.thr


def peaks_forward(input, win_size=3, filter_mode=0, thresholds=None, want_agg=True, cap=None):
    """Device op with the reference's result shape.  Returns (peaks int64 [Npk,5], agg [B,A] or None, thr [B,A]); the
    exact-size peak list needs the count on the host (one 4-byte read, like torch.nonzero in the reference)."""
    assert input.is_cuda and input.dtype == torch.float32 and input.dim() == 5
    input = input.contiguous()
    B, A = input.shape[:2]
    plan = PeaksPlan(input.shape, input.device, win_size, filter_mode, want_agg, cap)
    thr_in = None
    if filter_mode == 2:
        thr_in = torch.as_tensor(thresholds, dtype=torch.float32, device=input.device).reshape(-1)
        thr_in = thr_in.expand(B * A).contiguous() if thr_in.numel() == 1 else thr_in.contiguous()
        assert thr_in.numel() == B * A
    plan.run(input, thr_in)
    npk = int(plan.n.item())
    if npk > plan.cap:                  # rare: more peaks than the default capacity, rerun with room
        plan = PeaksPlan(input.shape, input.device, win_size, filter_mode, want_agg, npk)
        plan.run(input, thr_in)
    return plan.peaks[:npk], plan.agg, plan.thr


class PeakStimulation(Function):
    @staticmethod
    def forward(ctx, input, return_aggregation, win_size, peak_filter):
        ctx.num_flags = 4
        assert win_size % 2 == 1, 'Window size for peak finding must be odd.'
        if peak_filter is None or peak_filter is False:
            mode, thr = 0, None
        elif peak_filter == "median" or peak_filter is median_filter or getattr(peak_filter, "__name__", "") == "_median_filter":
            mode, thr = 1, None
        elif isinstance(peak_filter, (int, float)):
            mode, thr = 2, float(peak_filter)
        else:
            t = peak_filter(input)
            t = torch.as_tensor(t, dtype=torch.float32, device=input.device)
            B, A = input.shape[:2]
            mode, thr = 2, (t.expand(B, A, 1, 1, 1).reshape(-1) if t.dim() == 5 else t.reshape(-1))
        peak_list, agg, _ = peaks_forward(input, win_size, mode, thr, want_agg=bool(return_aggregation))
        ctx.mark_non_differentiable(peak_list)
        if return_aggregation:
            ctx.save_for_backward(peak_list)
            ctx.in_shape = tuple(input.shape)
            return peak_list, agg
        return peak_list

    @staticmethod
    def backward(ctx, grad_peak_list, grad_output):
        (peak_list,) = ctx.saved_tensors
        B, A, S, H, W = ctx.in_shape
        L = _lib.lib()
        dev = grad_output.device
        grad_in = torch.empty((B, A, S, H, W), dtype=torch.float32, device=dev)
        n = torch.tensor([peak_list.shape[0]], dtype=torch.int32, device=dev)
        g = grad_output.contiguous().float()
        pl = peak_list.contiguous()
        _lib.check(L.b200seg_peaks3d_bwd_dev(_lib.ptr(pl), _lib.ptr(n), int(pl.shape[0]), _lib.ptr(g),
                                             _lib.ptr(grad_in), B, A, S, H, W, _lib.current_stream()), "peaks3d_bwd")
        return (grad_in,) + (None,) * ctx.num_flags


def peak_stimulation_3d(input, return_aggregation=True, win_size=3, peak_filter=None):
    return PeakStimulation.apply(input, return_aggregation, win_size, peak_filter)
