"""Timing of the Upsample / Downsample backward at full-size shapes (32 samples): one-pass kernels against the
re-indexing compositions (`_force_unfused_upconv` / `_force_unfused_downconv`)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
net._ensure_grads()
R = 32 * 34
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
n = len(net.in_out)
for j in range(n - 1):
    wname, bname = f"ups.{j}.3.1.weight", f"ups.{j}.3.1.bias"
    cout, cin, _ = net.specs[wname]
    Lh = 625 * 2 ** j
    x = torch.randn(R, cin, Lh, device="cuda"); du = torch.randn(R, cout, 2 * Lh, device="cuda")
    t = {}
    for unfused in (True, False):
        net._force_unfused_upconv = unfused
        t[unfused] = tm(lambda: net._upconv_bwd(du, x, wname, bname, True, None, 34))
    by = R * Lh * 4 * (2 * cout + 2 * cin) / 1e6
    print(f"upsample  {wname} {cin}->{cout} Lh={Lh}: composed {t[True]:.0f} us, one pass {t[False]:.0f} us "
          f"(hbm-min {by / 6.55e3 * 1000:.0f} us, {by / 6.55e3 * 1000 / t[False] * 100:.0f} %)")
    del x, du
for i in range(n - 1):
    wname, bname = f"downs.{i}.3.weight", f"downs.{i}.3.bias"
    if wname not in net.specs or net.specs[wname][2] != 4:
        continue
    cout, cin, _ = net.specs[wname]
    L = 40000 // 2 ** i
    x = torch.randn(R, cin, L, device="cuda"); du = torch.randn(R, cout, L // 2, device="cuda")
    t = {}
    for unfused in (True, False):
        net._force_unfused_downconv = unfused
        t[unfused] = tm(lambda: net._downconv_bwd(du, x, wname, bname, True, None, 34))
    by = R * L * 4 * (cout // 2 + 2 * cin) / 1e6 if True else 0
    by = R * 4 * (cout * (L // 2) + 2 * cin * L) / 1e6
    print(f"downsample {wname} {cin}->{cout} L={L}: composed {t[True]:.0f} us, one pass {t[False]:.0f} us "
          f"(hbm-min {by / 6.55e3 * 1000:.0f} us, {by / 6.55e3 * 1000 / t[False] * 100:.0f} %)")
    del x, du
