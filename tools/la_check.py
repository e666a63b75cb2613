"""LA parity + timing: TC kernels vs oracle (CPU) on several shapes, then timing at level-0 full size."""
import sys, time
sys.path.insert(0, "tests")
import torch
from _util import make_net, TINY
import dquartic_oracle as O
net, P = make_net()
net._ensure_grads()
def rel(a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
for pre, C, L in [("downs.0.2", 4, 320), ("downs.2.2", 8, 80), ("ups.0.2", 16, 5), ("ups.3.2", 12, 40), ("downs.1.2", 4, 4500), ("downs.5.2", 12, 2049)]:
    R = 6
    g = torch.Generator().manual_seed(2)
    x = torch.randn(R, C, L, generator=g) * 1.5
    dres = torch.randn(R, C, L, generator=g)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre + ".")}
    xr = x.clone().requires_grad_(True)
    ref = O.linear_attention(Pg, pre, xr)
    ref.backward(dres)
    net._gflat.zero_()
    out, saved = net._la_fwd(pre, x.cuda(), True)
    dx = net._la_bwd(pre, saved, dres.cuda())
    torch.cuda.synchronize()
    errs = {k.split(".", 2)[2]: rel(net._params[k].grad, v.grad) for k, v in Pg.items()}
    print(f"{pre} C={C} L={L}: fwd {rel(out, ref):.2e} (minus x: {rel(out - x.cuda(), ref - x):.2e}) dx {rel(dx, xr.grad):.2e}", {k: f"{v:.1e}" for k, v in errs.items()})
# timing at full size, 8 samples
from dquartic.model.unet1d import UNet1d
for C, L, pre in [(4, 40000, "downs.0.2"), (8, 10000, "downs.2.2"), (16, 625, "downs.6.2")]:
    R = 8 * 34
    x = torch.randn(R, C, L, device="cuda"); dres = torch.randn_like(x)
    for _ in range(2):
        out, saved = net._la_fwd(pre, x, True); dx = net._la_bwd(pre, saved, dres)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(5): out, saved = net._la_fwd(pre, x, True)
    e[1].record()
    for _ in range(5): dx = net._la_bwd(pre, saved, dres)
    e[2].record(); torch.cuda.synchronize()
    print(f"C={C} L={L} R={R}: fwd {e[0].elapsed_time(e[1])/5:.2f} ms  bwd {e[1].elapsed_time(e[2])/5:.2f} ms")
