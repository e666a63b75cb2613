"""Kernel-time breakdown of one training micro-batch (torch.profiler / CUPTI), full-size model."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-deconvolution-dia-msms-data_b200"))
import torch
from dquartic.model.unet1d import UNet1d
from dquartic.model.model import DDIMDiffusionModel
from dquartic import _native as N

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
t0 = time.time()
net = UNet1d(dim=4, channels=1, dim_mults=(1, 2, 2, 3, 3, 4, 4), conditional=True, init_cond_channels=1,
             attn_cond_channels=1, downsample_dim=40000, device=dev)
d = DDIMDiffusionModel(net, device=dev)
d._prepare_training(1e-5)
print("init s", time.time() - t0, "params", net.n_flat)
x0 = torch.rand(mb, 34, 40000, device=dev) * (torch.rand(mb, 34, 40000, device=dev) < 0.02)
cond = 0.5 * x0 + 0.5 * torch.rand_like(x0) * (torch.rand_like(x0) < 0.02)
m1 = torch.rand(mb, 34, device=dev)
for _ in range(2):
    d._train_one_batch(x0, cond, m1)
torch.cuda.synchronize()
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
t0 = time.time()
for _ in range(2):
    d._train_one_batch(x0, cond, m1)
torch.cuda.synchronize()
dt = (time.time() - t0) / 2
print(f"step (b={mb}) {dt*1000:.1f} ms -> {mb/dt:.1f} samples/s; launches/step {N.launches // 4}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    d._train_one_batch(x0, cond, m1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
