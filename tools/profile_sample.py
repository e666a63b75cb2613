"""Kernel-time breakdown of DDIM sampling (forward-only path), full-size model, 32 windows, 4 steps."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-deconvolution-dia-msms-data_b200"))
import torch
from dquartic.model.unet1d import UNet1d
from dquartic.model.model import DDIMDiffusionModel
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
net = UNet1d(dim=4, channels=1, dim_mults=(1, 2, 2, 3, 3, 4, 4), conditional=True, init_cond_channels=1,
             attn_cond_channels=1, downsample_dim=40000, device=dev)
d = DDIMDiffusionModel(net, device=dev)
net.eval()
x0 = torch.rand(nw, 34, 40000, device=dev) * (torch.rand(nw, 34, 40000, device=dev) < 0.02)
cond = 0.5 * x0 + 0.5 * torch.rand_like(x0) * (torch.rand_like(x0) < 0.02)
m1 = torch.rand(nw, 34, device=dev)
xT = torch.randn_like(x0)
with torch.no_grad():
    d.sample(xT, cond, m1, num_steps=2)
    torch.cuda.synchronize()
    t0 = time.time()
    d.sample(xT, cond, m1, num_steps=10)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 10
    print(f"{nw} windows: {dt*1000:.1f} ms per DDIM step -> {nw/(dt*50):.2f} maps/s at 50 steps; max mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        d.sample(xT, cond, m1, num_steps=2)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
