"""A few complex128 power-method applications on the kind::i8 path (ncu target: launch lists / captures)."""
import sys; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from qmps_b200 import batched as B, _lib as L
L.require_device().qmps_set_option(b"i8_power", 2)
g = torch.Generator(device="cuda").manual_seed(0)
for D, N in ((64, 512), (256, 32)):
    A = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2 * D)
    Bt = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2 * D)
    B.tm_power(A, Bt, 3)
torch.cuda.synchronize()
