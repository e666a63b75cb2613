"""Block-level differential check at the shapes of the 3-level config (R=6, rt=3)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from _util import TINY, make_net, rel_err
import dquartic_oracle as O
cfg = dict(TINY, dim=4, dim_mults=[1, 2, 4], downsample_dim=1300); rt, b = 3, 2
net, P = make_net(cfg, seed=5); net.train(); net._ensure_grads()
R = b * rt
t = torch.tensor([40, 870])
temb = O.time_mlp(P, t, cfg["dim"]) if hasattr(O, "time_mlp") else None
tp = net._time_path_fwd(t.cuda(), b, True)
temb = tp[3].cpu()
net._dSS = torch.zeros(b, net.ss_total, device="cuda")
for pre, c1, c2, L in [("downs.1.0", 4, 0, 650), ("downs.1.1", 4, 0, 650), ("downs.2.0", 8, 0, 325), ("ups.0.0", 16, 8, 325), ("ups.1.1", 8, 4, 650), ("ups.2.0", 4, 4, 1300)]:
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(R, c1, L, generator=g); x2 = torch.randn(R, c2, L, generator=g) if c2 else None
    co = P[pre + ".block1.proj.weight"].shape[0]
    dout = torch.randn(R, co, L, generator=g)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre + ".")}
    x1r = x1.clone().requires_grad_(True); x2r = x2.clone().requires_grad_(True) if c2 else None
    tr = temb.clone().requires_grad_(True)
    ref = O.resnet_block(Pg, pre, torch.cat((x1r, x2r), 1) if c2 else x1r, tr, rt)
    ref.backward(dout)
    net._gflat.zero_(); net._dSS.zero_()
    out, saved = net._resnet_fwd(pre, x1.cuda(), x2.cuda() if c2 else None, rt, True)
    dx1, dx2 = net._resnet_bwd(pre, saved, dout.cuda(), rt)
    errs = {k.split(".", 2)[2]: rel_err(net._params[k].grad, v.grad) for k, v in Pg.items() if "mlp" not in k}
    print(pre, L, f"fwd {rel_err(out, ref):.1e} dx1 {rel_err(dx1, x1r.grad):.1e}", {k: f"{v:.1e}" for k, v in errs.items() if v > 1e-5})
