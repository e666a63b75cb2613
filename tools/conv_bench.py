"""Per-kernel timing of the small-channel conv family at full-size shapes (8 samples), vs the HBM bound."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
from dquartic import _native as N
net, _ = make_net()
net._ensure_grads()
b, rt = 8, 34
net._time_path_fwd(torch.zeros(b, dtype=torch.long, device="cuda"), b, False)
net._dSS = torch.zeros(b, net.ss_total, device="cuda")
def tm(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
R = b * rt
for pre, c1, c2, L in [("downs.0.0", 4, 0, 40000), ("downs.2.0", 8, 0, 10000), ("ups.4.0", 8, 8, 10000), ("downs.4.0", 12, 0, 2500)]:
    x1 = torch.randn(R, c1, L, device="cuda"); x2 = torch.randn(R, c2, L, device="cuda") if c2 else None
    w, bn, gname = pre + ".block1.proj.weight", pre + ".block1.proj.bias", pre + ".block1.norm.g"
    cout = net.specs[w][0]
    y, u = net._conv_fwd(x1, x2, w, bn, 3, 1, 1, 1, L, g=gname, ss=net.ss_off[pre + ".mlp.1"], act=1, save_u=True, rps=rt)
    dy = torch.randn_like(y)
    elt = R * L * 4 / 1e6  # MB per channel-plane set
    t_f = tm(lambda: net._conv_fwd(x1, x2, w, bn, 3, 1, 1, 1, L, g=gname, ss=net.ss_off[pre + ".mlp.1"], act=1, save_u=True, rps=rt))
    t_b = tm(lambda: net._block_bwd(dy, u, gname, net.ss_off[pre + ".mlp.1"], 1, rt))
    du = net._block_bwd(dy, u, gname, net.ss_off[pre + ".mlp.1"], 1, rt)
    dx1 = torch.empty_like(x1); dx2 = torch.empty_like(x2) if c2 else None
    t_d = tm(lambda: N.call("dq_conv1d_bwd_data", du, net._w(w), dx1, c1, 0, dx2, c2, 0, cout, 3, 1, 1, 1, R, L, L))
    t_w = tm(lambda: N.call("dq_conv1d_bwd_weight", du, x1, c1, x2, c2, None, 0, net._gw(w), net._gw(bn), cout, 3, 1, 1, 1, R, L, L, rt))
    t_fu = tm(lambda: net._conv_bwd_fused(dy, u, gname, net.ss_off[pre + ".mlp.1"], 1, x1, x2, w, bn, 3, rps=rt))
    cin = c1 + c2
    hb = lambda mb: mb / 6.55e3 * 1000  # us at 6.55 TB/s
    print(f"{pre} cin={cin} cout={cout} L={L}: fwd {t_f:.0f} us (hbm {hb(elt*(cin+2*cout)):.0f}) | block_bwd {t_b:.0f} (hbm {hb(elt*3*cout):.0f}) | bwd_data {t_d:.0f} (hbm {hb(elt*(cin+cout)):.0f}) | bwd_weight {t_w:.0f} (hbm {hb(elt*(cin+cout)):.0f}) || FUSED bwd {t_fu:.0f} (hbm {hb(elt*(2*cout+2*cin)):.0f})")
