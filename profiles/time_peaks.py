"""Times the peak-stimulation op (device part only, CUDA events, L2 flushed) on the BASELINE config-3 maps."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import b200seg
from b200seg import synth
from b200seg.peak_stimulation_3d import PeaksPlan

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for name, shape, ch, nv in (("14x32x128x128", (32, 128, 128), 14, 1), ("8x14x32x128x128", (32, 128, 128), 14, 8), ("1x128x512x512", (128, 512, 512), 1, 1)):
    x = torch.from_numpy(synth.response_map(np.random.default_rng(1003), shape, n_peaks=60, channels=ch)).to(dev)
    if nv > 1:
        x = x.repeat(nv, 1, 1, 1, 1).contiguous()
    plan = PeaksPlan(x.shape, dev, 3, 1, cap=(0 if os.environ.get('PEAKS_CAP0') else None))
    for _ in range(3):
        plan.run(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.run(x); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    out[name] = {"ms_median": float(np.median(ts)), "ms_min": float(min(ts)), "n_peaks": int(plan.n.item()),
                 "gbs_8V": 8 * x.numel() / (np.median(ts) * 1e-3) / 1e9}
    # kernel by kernel: the op stops after its first k launches (1 scan, 2 median pass(es), 99 everything)
    L = b200seg._lib.lib()
    st = []
    for k in (0, 1, 2, 3, 99):
        L.b200seg_set_option(b"peaks_stop_after", k)
        tk = []
        for _ in range(8):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); plan.run(x); b.record(); torch.cuda.synchronize()
            tk.append(a.elapsed_time(b))
        st.append(float(np.median(tk)))
    L.b200seg_set_option(b"peaks_stop_after", 99)
    out[name]["stages_ms"] = {"memset": st[0], "scan": st[1] - st[0], "median_a": st[2] - st[1], "median_b": st[3] - st[2], "finalize": st[4] - st[3]}
print(json.dumps({"variant": os.environ.get("B200SEG_PEAKS_VARIANT", "0"), "peaks": out}))
