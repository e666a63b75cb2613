"""Timing of dq_adamw on the full flat buffer (1.2 B parameters), aligned and deliberately misaligned (scalar kernel)."""
import torch, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-deconvolution-dia-msms-data_b200"))
from dquartic import _native as N
n = 1_204_738_392
bufs = [torch.randn(n + 4, device="cuda") * 0.01 for _ in range(2)] + [torch.zeros(n + 4, device="cuda") for _ in range(2)]
bufs[3].fill_(1e-4)
coef = torch.tensor([1.0, 1.0], device="cuda")
for off, name in ((0, "aligned (vector kernel)"), (1, "misaligned (scalar kernel)")):
    p, g, m, v = (b[off:off + n] for b in bufs)
    f = lambda: N.call("dq_adamw", p, g, m, v, n, coef, 1e-5, 0.9, 0.999, 1e-8, 0.01, 1e-5 / 0.1, 0.0316)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5
    print(f"adamw {name}: {t:.2f} ms = {28 * n / t / 1e9:.2f} TB/s")
# the two kernels must agree bit for bit
ps = []
for off in (0, 1):
    torch.manual_seed(0)
    b = [torch.randn(1000003 + 4, device="cuda") * 0.01 for _ in range(2)] + [torch.rand(1000003 + 4, device="cuda") * 1e-3 for _ in range(2)]
    ref = [x[:1000003].clone() for x in b]
    q = [x[off:off + 1000003] for x in b]
    for a, r in zip(q, ref): a.copy_(r)
    N.call("dq_adamw", q[0], q[1], q[2], q[3], 1000003, coef, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1e-3 / 0.1, 0.0316)
    ps.append([a.clone() for a in (q[0], q[2], q[3])])
print("vector == scalar bit for bit:", all(torch.equal(a, b) for a, b in zip(*ps)))
