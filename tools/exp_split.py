"""cfg 4: fused theta -> energy kernel against the phase-split pipeline (ansatz kernel -> A in HBM -> energy kernel)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmps_b200 import _lib as L, batched as B, represent as R
from qmps_b200.ground_state import Hamiltonian
lib = L.require_device(); dev = torch.device("cuda", 0)
N = 65536
rng = np.random.default_rng(3)
theta = torch.from_numpy(rng.normal(size=(N, 24))).to(dev)
prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
sh = torch.tensor(B.ROTO3_SHIFTS, dtype=torch.float64, device=dev)
th3 = theta[:, None, :].repeat(1, 3, 1); th3[:, :, 5] += sh[None, :]; th3 = th3.reshape(3 * N, 24).contiguous()

def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out
ms_f, ef = t(lambda: B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS))
ms_a, A = t(lambda: B.ansatz_tensors(prog, th3))
ms_e, es = t(lambda: B.energy_tensor(A, H))
ef = ef[0] if isinstance(ef, tuple) else ef; es = es[0] if isinstance(es, tuple) else es
print(json.dumps({"fused_ms": ms_f, "ansatz_ms": ms_a, "energy_tensor_ms": ms_e, "split_ms": ms_a + ms_e,
                  "max_abs_diff": float((ef.reshape(-1) - es.reshape(-1)).abs().max())}))
