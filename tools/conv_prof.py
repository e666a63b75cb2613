"""A few launches of the fused conv backward at level-0 / level-2 shapes (8 samples) for ncu."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
net._ensure_grads()
b, rt = 8, 34
net._time_path_fwd(torch.zeros(b, dtype=torch.long, device="cuda"), b, False)
net._dSS = torch.zeros(b, net.ss_total, device="cuda")
R = b * rt
for pre, c1, c2, L in [("downs.0.0", 4, 0, 40000), ("ups.6.0", 4, 4, 40000), ("downs.2.0", 8, 0, 10000)]:
    x1 = torch.randn(R, c1, L, device="cuda"); x2 = torch.randn(R, c2, L, device="cuda") if c2 else None
    w, bn, gname = pre + ".block1.proj.weight", pre + ".block1.proj.bias", pre + ".block1.norm.g"
    y, u = net._conv_fwd(x1, x2, w, bn, 3, 1, 1, 1, L, g=gname, ss=net.ss_off[pre + ".mlp.1"], act=1, save_u=True, rps=rt)
    dy = torch.randn_like(y)
    for _ in range(2):
        net._conv_bwd_fused(dy, u, gname, net.ss_off[pre + ".mlp.1"], 1, x1, x2, w, bn, 3, rps=rt)
torch.cuda.synchronize()
print("ok")
