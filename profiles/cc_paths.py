"""Which path did the instances of the bench workload take through the largest-component step?"""
import ctypes, sys
import numpy as np
sys.path.insert(0, ".")
import subprocess, json, os
from b200seg import _lib
L = _lib.lib()
c = (ctypes.c_longlong * 8)()
L.b200seg_largest_cc_path_counts(c, 1)
import bench  # noqa
import torch
sys.argv = ["bench.py", "--steps", "1", "--warmup", "0", "--chain-only"]
try:
    bench.main()
except SystemExit:
    pass
torch.cuda.synchronize()
L.b200seg_largest_cc_path_counts(c, 0)
print("path counts", list(c))
