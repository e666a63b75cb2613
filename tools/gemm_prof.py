"""A few launches of the mid-stage GEMMs (fwd conv M=1152, wgrad K=9216) for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
dev = "cuda"
Nm, Mp = 10000, 32 * 36
A = torch.randn(Mp, Nm, device=dev).bfloat16()
W = torch.randn(3, Nm, Nm, device=dev).bfloat16()
U = torch.empty(Mp, Nm, device=dev)
for _ in range(2):
    net._gemm(A, Mp, Nm, Nm, W, Nm, Nm, Nm, Nm * Nm, 3, U, Nm, None, 0, Mp, Nm, Nm, 3, (-1, 0, 1), (0, 0, 0), (0, 0, 0), (0, 1, 2))
Kc = 8 * Mp
dUT = torch.randn(Nm, Kc, device=dev).bfloat16()
AT3 = torch.randn(3, Nm, Kc, device=dev).bfloat16()
dW = torch.zeros(3, Nm, Nm, device=dev)
net._gemm(dUT, Nm, Kc, Kc, AT3, Nm, Kc, Kc, Nm * Kc, 3, dW, Nm, None, 1, Nm, Nm, Kc, 1, (0,), (0,), (0,), (0,), nz=3, z_b_tap_step=1, z_c_stride=Nm * Nm)
torch.cuda.synchronize()
print("ok")
