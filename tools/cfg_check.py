"""Gradient errors of a non-default config vs the oracle (diagnostic)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import TINY, make_net, rel_err
import dquartic_oracle as O
from dquartic.model.model import DDIMDiffusionModel
cfg = dict(TINY, dim=4, dim_mults=[1, 2, 4], downsample_dim=1300); rt, mz = 3, 1300
net, P = make_net(cfg, seed=5); net.train()
d = DDIMDiffusionModel(net, device="cuda")
b = 2
g = torch.Generator().manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 17)
x0 = torch.rand(b, rt, mz, generator=g) * (torch.rand(b, rt, mz, generator=g) < 0.3)
c2 = 0.5 * x0 + 0.5 * torch.rand(b, rt, mz, generator=g) * (torch.rand(b, rt, mz, generator=g) < 0.3)
c1 = torch.rand(b, rt, generator=g)
noise = torch.randn(b, rt, mz, generator=g)
t = torch.tensor([40, 870])
Pg = {k: v.clone().requires_grad_(not k.endswith("freqs")) for k, v in P.items()}
_, _, ab = O.schedule_tables(1000, "cosine")
ref_loss, _ = O.train_loss(Pg, cfg, ab, x0, c2, c1, t, noise)
ref_loss.backward()
net.zero_grad()
loss = d.train_step(x0.cuda(), c2.cuda(), c1.cuda(), noise=((noise + 1) * 0.5).cuda(), t=t.cuda())
loss.mean().backward()
print("loss", float(loss.mean()), float(ref_loss))
errs = sorted(((rel_err(net._params[k].grad, v.grad), k, float(v.grad.abs().max())) for k, v in Pg.items() if not k.endswith("freqs")), reverse=True)
for e, k, m in errs[:3]: print(f"{e:.4f} {k} max|g|={m:.3e}")
