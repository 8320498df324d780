"""complex128 power method: tcgen05 kind::i8 path against the FP64 tensor-pipe (DMMA) path, per bond dimension."""
import sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bench_legs as BL
from qmps_b200 import batched as B, _lib as L
lib = L.require_device(); dev = torch.device("cuda", 0)
peaks = BL.load_peaks()
for D, N in ((64, 512), (128, 128), (192, 64), (256, 32), (512, 8)):
    for flag in (2, 0):
        lib.qmps_set_option(b"i8_power", flag)
        r = BL.leg_power(torch, B, dev, D, peaks, "c128", nprob=N)
        print(json.dumps({"D": D, "N": N, "i8": flag, "apps_per_s": r["value"], "ms": r["ms_per_step"], "algo_tflops": r["roofline"]["algorithmic_tflops"]}), flush=True)
lib.qmps_set_option(b"i8_power", 1)
