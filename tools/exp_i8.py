import sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bench_legs as BL
from qmps_b200 import batched as B, _lib as L
lib = L.require_device(); dev = torch.device("cuda", 0)
peaks = BL.load_peaks()
for D in (64, 128, 256):
    for flag in (1, 0):
        lib.qmps_set_option(b"i8_power", flag)
        r = BL.leg_power(torch, B, dev, D if D != 128 else 128, peaks, "c128") if D != 128 else None
        if r: print(json.dumps({"D": D, "i8": flag, "apps_per_s": r["value"], "ms": r["ms_per_step"], "algo_tflops": r["roofline"]["algorithmic_tflops"]}))
lib.qmps_set_option(b"i8_power", 1)
