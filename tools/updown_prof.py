"""One Upsample and one Downsample backward at level-0 shapes (32 samples) for ncu: python tools/updown_prof.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
net._ensure_grads()
R = 32 * 34
x = torch.randn(R, 8, 20000, device="cuda"); du = torch.randn(R, 4, 40000, device="cuda")
xd = torch.randn(R, 4, 40000, device="cuda"); dud = torch.randn(R, 4, 20000, device="cuda")
for _ in range(2):
    net._upconv_bwd(du, x, "ups.5.3.1.weight", "ups.5.3.1.bias", True, None, 34)
    net._downconv_bwd(dud, xd, "downs.0.3.weight", "downs.0.3.bias", True, None, 34)
    net._conv_fwd(xd, None, "downs.0.3.weight", "downs.0.3.bias", 4, 2, 1, 1, 20000)
torch.cuda.synchronize()
print("ok")
