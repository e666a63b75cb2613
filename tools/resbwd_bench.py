"""Timing of the ResnetBlock backward at full-size shapes (32 samples), against its algorithmic HBM bytes.
DQ_B200_LIB / env switches select builds; pass `nores` to force the unfused (three-call) path."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
from dquartic import _native as N
net, _ = make_net()
net._ensure_grads()
b, rt = 32, 34
net._time_path_fwd(torch.zeros(b, dtype=torch.long, device="cuda"), b, False)
net._dSS = torch.zeros(b, net.ss_total, device="cuda")
R = b * rt
if "nores" in sys.argv:
    _call = N.call
    def call(name, *a, **k):
        if name == "dq_conv_bwd_fused_res":
            return 1
        return _call(name, *a, **k)
    N.call = call
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for pre, c1, c2, L in [("downs.0.0", 4, 0, 40000), ("ups.6.0", 4, 4, 40000), ("ups.5.0", 8, 4, 20000), ("downs.2.0", 8, 0, 10000),
                       ("ups.4.0", 8, 8, 10000), ("ups.3.0", 12, 8, 5000), ("downs.4.0", 12, 0, 2500), ("ups.2.0", 12, 12, 2500),
                       ("downs.5.0", 12, 0, 1250), ("ups.1.0", 16, 12, 1250), ("downs.6.0", 16, 0, 625), ("ups.0.0", 16, 16, 625)]:
    x1 = torch.randn(R, c1, L, device="cuda"); x2 = torch.randn(R, c2, L, device="cuda") if c2 else None
    cout = net.specs[pre + ".block1.proj.weight"][0]
    out, saved = net._resnet_fwd(pre, x1, x2, rt, True)
    dout = torch.randn_like(out)
    t = tm(lambda: net._resnet_bwd(pre, saved, dout, rt))
    cin = c1 + c2
    rows = 7 * cout + 2 * cin   # dout, u2, h1 -> dh1 | dh1, u1, x, dout -> dx
    by = R * L * 4 * rows / 1e6
    print(f"{pre} cin={cin} cout={cout} L={L}: bwd {t:.0f} us (hbm-min {by/6.55e3*1000:.0f} us, {by/6.55e3*1000/t*100:.0f} %)")
