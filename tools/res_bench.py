"""Timing of the fused ResnetBlock forward at full-size shapes (8 samples)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
b, rt = 8, 34
net._time_path_fwd(torch.zeros(b, dtype=torch.long, device="cuda"), b, False)
R = b * rt
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for pre, c1, c2, L in [("downs.0.0", 4, 0, 40000), ("ups.6.0", 4, 4, 40000), ("downs.2.0", 8, 0, 10000), ("ups.4.0", 8, 8, 10000), ("downs.4.0", 12, 0, 2500), ("ups.2.0", 12, 12, 2500), ("ups.1.0", 16, 12, 1250), ("downs.6.0", 16, 0, 625)]:
    x1 = torch.randn(R, c1, L, device="cuda"); x2 = torch.randn(R, c2, L, device="cuda") if c2 else None
    cout = net.specs[pre + ".block1.proj.weight"][0]
    for save in (True, False):
        t = tm(lambda: net._resnet_fwd(pre, x1, x2, rt, save))
        by = R * L * 4 * ((c1 + c2) + (3 * cout if save else 0) + cout) / 1e6
        print(f"{pre} cin={c1+c2} cout={cout} L={L} save={save}: {t:.0f} us (hbm {by/6.55e3*1000:.0f} us)")
