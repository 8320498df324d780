"""complex64 power method on tcgen05: fp32 slab images split in shared memory (tc_presplit = 0) against images that
carry both TF32 planes (1), per bond dimension."""
import sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import bench_legs as BL
from qmps_b200 import batched as B, _lib as L
lib = L.require_device(); dev = torch.device("cuda", 0)
peaks = BL.load_peaks()
for D, N in ((64, 512), (128, 128), (192, 64), (256, 32), (512, 8)):
    for flag in (0, 1):
        lib.qmps_set_option(b"tc_presplit", flag)
        r = BL.leg_power(torch, B, dev, D, peaks, "c64", nprob=N, with_e2e=False)
        print(json.dumps({"D": D, "N": N, "presplit": flag, "apps_per_s": r["value"], "ms": r["ms_per_step"], "algo_tflops": r["roofline"]["algorithmic_tflops"]}), flush=True)
lib.qmps_set_option(b"tc_presplit", -1)
