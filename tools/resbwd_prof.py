"""Kernel list of one ResnetBlock backward at a given shape: python tools/resbwd_prof.py pre c1 c2 L [samples]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
net._ensure_grads()
pre, c1, c2, L = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
b, rt = (int(sys.argv[5]) if len(sys.argv) > 5 else 32), 34
net._time_path_fwd(torch.zeros(b, dtype=torch.long, device="cuda"), b, False)
net._dSS = torch.zeros(b, net.ss_total, device="cuda")
R = b * rt
x1 = torch.randn(R, c1, L, device="cuda"); x2 = torch.randn(R, c2, L, device="cuda") if c2 else None
out, saved = net._resnet_fwd(pre, x1, x2, rt, True)
dout = torch.randn_like(out)
for _ in range(2): net._resnet_bwd(pre, saved, dout, rt)
torch.cuda.synchronize()
if os.environ.get("NOPROF"):
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net._resnet_bwd(pre, saved, dout, rt)
    torch.cuda.synchronize()
for e in prof.key_averages():
    print(f"{e.device_time_total:9.1f} us x{e.count}  {e.key[:150]}")
