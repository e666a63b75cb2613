"""BASELINE.json configs[4] at full size on one GPU: widened U-Net (dim 8 -> C = 8..32, 20000-channel mid stage,
4.8 B parameters) on 136 x 40000 maps.  Runs a few optimizer steps on one fixed batch (loss must fall), checks the
micro-batched step against the single pass and reports memory and throughput."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-deconvolution-dia-msms-data_b200"))
import torch
from dquartic.model.unet1d import UNet1d
from dquartic.model.model import DDIMDiffusionModel

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rt, mz = 136, 40000
dev = torch.device("cuda")
torch.manual_seed(0)
torch.cuda.manual_seed_all(0)
t0 = time.time()
net = UNet1d(dim=8, channels=1, dim_mults=(1, 2, 2, 3, 3, 4, 4), conditional=True, init_cond_channels=1,
             attn_cond_channels=1, downsample_dim=mz, device=dev)
d = DDIMDiffusionModel(net, device=dev)
d._prepare_training(1e-5)
torch.cuda.synchronize()
print(f"init {time.time() - t0:.1f} s, params {net.n_flat:,}, mem {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
g = torch.Generator(device="cuda").manual_seed(3)
x0 = torch.rand(b, rt, mz, device=dev, generator=g) * (torch.rand(b, rt, mz, device=dev, generator=g) < 0.02)
cond = 0.5 * x0 + 0.5 * torch.rand(b, rt, mz, device=dev, generator=g) * (torch.rand(b, rt, mz, device=dev, generator=g) < 0.02)
m1 = torch.rand(b, rt, device=dev, generator=g)
noise = torch.rand(b, rt, mz, device=dev, generator=g)
t = torch.randint(0, 1000, (b,), device=dev, generator=g)
losses = []
for i in range(steps):
    torch.cuda.synchronize(); t1 = time.time()
    loss = d._train_one_batch(x0, cond, m1, noise=noise, t=t)
    torch.cuda.synchronize(); dt = time.time() - t1
    losses.append(float(loss))
    print(f"step {i}: loss {float(loss):.6f} grad-norm {float(d.optimizer.last_grad_norm):.4f} {dt * 1e3:.0f} ms "
          f"({b / dt:.2f} samples/s) max mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
assert all(l == l for l in losses) and losses[-1] < losses[0], losses
net.eval()
with torch.no_grad():
    xT = torch.randn(1, rt, mz, device=dev, generator=g)
    torch.cuda.synchronize(); t1 = time.time()
    x, pn = d.sample(xT, cond[:1], m1[:1], num_steps=3)
    torch.cuda.synchronize()
    print(f"3 DDIM steps: {(time.time() - t1) * 1e3:.0f} ms, finite {bool(torch.isfinite(x).all())}")
print("OK")
