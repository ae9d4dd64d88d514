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


def peaks_forward(input, win_size=3, filter_mode=0, thresholds=None, want_agg=True, cap=None):
    """Device op.  Returns (peaks int64 [Npk,5], agg [B,A] or None, thr [B,A])."""
    assert input.is_cuda and input.dtype == torch.float32 and input.dim() == 5
    assert win_size % 2 == 1, 'Window size for peak finding must be odd.'
    L = _lib.lib()
    input = input.contiguous()
    B, A, S, H, W = input.shape
    dev = input.device
    V = S * H * W
    if cap is None:
        cap = min(B * A * V, max(1 << 16, (B * A * V) // 8))
    peaks = torch.empty((max(cap, 1), 5), dtype=torch.int64, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    agg = torch.empty((B, A), dtype=torch.float32, device=dev) if want_agg else None
    thr = torch.empty((B, A), dtype=torch.float32, device=dev)
    ws_bytes = L.b200seg_peaks3d_workspace_bytes(B, A, S, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    thr_in = None
    if filter_mode == 2:
        thr_in = torch.as_tensor(thresholds, dtype=torch.float32, device=dev).reshape(-1)
        thr_in = thr_in.expand(B * A).contiguous() if thr_in.numel() == 1 else thr_in.contiguous()
        assert thr_in.numel() == B * A

    def run(cap_, peaks_):
        _lib.check(L.b200seg_peaks3d_dev(_lib.ptr(input), B, A, S, H, W, int(win_size), int(filter_mode),
                                         _lib.ptr(thr_in), _lib.ptr(peaks_), cap_, _lib.ptr(n), _lib.ptr(agg),
                                         _lib.ptr(thr), _lib.ptr(ws), ws_bytes, _lib.current_stream()), "peaks3d")
    run(cap, peaks)
    npk = int(n.item())
    if npk > cap:                       # rare: more peaks than the default capacity, rerun with room
        peaks = torch.empty((npk, 5), dtype=torch.int64, device=dev)
        run(npk, peaks)
    return peaks[:npk], agg, thr


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
