"""ncu target: the peak-stimulation op on 8 volumes' maps (8x14x32x128x128), two warm-up runs and one profiled run."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import b200seg
from b200seg import synth
from b200seg.peak_stimulation_3d import PeaksPlan
dev = torch.device("cuda")
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = torch.from_numpy(synth.response_map(np.random.default_rng(1003), (32, 128, 128), n_peaks=60, channels=14)).to(dev)
x = x.repeat(nv, 1, 1, 1, 1).contiguous()
plan = PeaksPlan(x.shape, dev, 3, 1)
for _ in range(3):
    plan.run(x)
torch.cuda.synchronize()
print("peaks", int(plan.n.item()))
