"""One forward + backward of a LinearAttention block for ncu: python tools/la_prof.py [C] [L] [samples]."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, P = make_net()
net._ensure_grads()
C, L, pre = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 40000, "downs.0.2"
pre = {4: "downs.0.2", 8: "downs.2.2", 12: "downs.4.2", 16: "downs.6.2"}[C]
R = (int(sys.argv[3]) if len(sys.argv) > 3 else 8) * 34
x = torch.randn(R, C, L, device="cuda"); dres = torch.randn_like(x)
for _ in range(2):
    out, saved = net._la_fwd(pre, x, True); dx = net._la_bwd(pre, saved, dres)
torch.cuda.synchronize()
print("ok")
