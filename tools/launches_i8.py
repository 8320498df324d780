import sys; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from qmps_b200 import batched as B
g = torch.Generator(device="cuda").manual_seed(0)
for D, N in ((64, 512), (256, 32)):
    A = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2 * D)
    Bt = torch.view_as_complex(torch.randn((N, 2, D, D, 2), dtype=torch.float64, device="cuda", generator=g)) / np.sqrt(2 * D)
    B.tm_power(A, Bt, 3)
torch.cuda.synchronize()
