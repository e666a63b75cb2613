"""BASELINE.json configs[4] at FULL size, data-parallel: widened U-Net (dim 8 -> C = 8..32, 20000-channel mid stage,
4,810,842,888 parameters) on 136 x 40000 maps, one process per GPU (torchrun), sharded optimizer (ZeRO-1 moments,
reduce-scatter / all-gather of the mid ranges).  A few optimizer steps on fixed per-rank batches: the loss must fall, the
ranks must hold identical parameters afterwards, memory and throughput are reported by rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/cfg5_dp.py [per_gpu_batch] [steps] [small] [lr]

`small` = the default network (dim 4, RT 34) instead: run it once as is and once with DQ_SHARDED_OPT=0 and compare the printed
losses digit by digit (the sharded optimizer with its deferred, overlapped parameter all-gather against the replicated one).
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-deconvolution-dia-msms-data_b200"))
import torch
import torch.distributed as dist
from dquartic.model.unet1d import UNet1d
from dquartic.model.model import DDIMDiffusionModel

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
small = len(sys.argv) > 3 and sys.argv[3] == "small"
lr = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-5
rt, mz = (34, 40000) if small else (136, 40000)
torch.manual_seed(0)
t0 = time.time()
net = UNet1d(dim=4 if small else 8, channels=1, dim_mults=(1, 2, 2, 3, 3, 4, 4), conditional=True, init_cond_channels=1,
             attn_cond_channels=1, downsample_dim=mz, device=dev)
dist.broadcast(net.flat_params(), src=0)
net.mark_params_modified()
d = DDIMDiffusionModel(net, device=dev)
d.micro_batch = min(b, 4)
d._prepare_training(lr)
torch.cuda.synchronize()
if rank == 0:
    plan = d._shard_plan()
    seg = d.optimizer._build_segments()[-1] if plan else None
    print(f"init {time.time() - t0:.1f} s, params {net.n_flat:,}, sharded ranges {plan}, "
          f"moments per rank {(seg[2] + seg[1]) if seg else net.n_flat:,} floats, "
          f"mem {torch.cuda.memory_allocated() / 1e9:.1f} GB", flush=True)
g = torch.Generator(device=dev).manual_seed(3 + rank)
x0 = torch.rand(b, rt, mz, device=dev, generator=g) * (torch.rand(b, rt, mz, device=dev, generator=g) < 0.02)
cond = 0.5 * x0 + 0.5 * torch.rand(b, rt, mz, device=dev, generator=g) * (torch.rand(b, rt, mz, device=dev, generator=g) < 0.02)
m1 = torch.rand(b, rt, device=dev, generator=g)
noise = torch.rand(b, rt, mz, device=dev, generator=g)
t = torch.randint(0, 1000, (b,), device=dev, generator=g)
losses = []
for i in range(steps):
    dist.barrier(); torch.cuda.synchronize(); t1 = time.time()
    loss = d._train_one_batch(x0, cond, m1, noise=noise, t=t)
    torch.cuda.synchronize(); dist.barrier(); dt = time.time() - t1
    lt = torch.tensor([loss], device=dev); dist.all_reduce(lt); losses.append(float(lt) / world)
    if rank == 0:
        print(f"step {i}: mean loss {losses[-1]:.9f} grad-norm {float(d.optimizer.last_grad_norm):.4f} {dt * 1e3:.0f} ms "
              f"({b * world / dt:.2f} samples/s on {world} GPUs) max mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
assert all(l == l for l in losses) and losses[-1] < losses[0], losses
# every rank must hold the same parameters after the sharded update + in-place all-gather
chk = net.flat_params()[: net.n_trainable_flat].double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert float(lo) == float(hi), (float(lo), float(hi))
from dquartic import _native
assert _native.la_tc_last_error() is None
if rank == 0:
    print(f"parameters identical on all {world} ranks (checksum {float(lo):.6f}); OK")
dist.destroy_process_group()
