"""GEMM throughput at the mid-stage shapes (micro-batch 32): fwd/dgrad (M=1152, N=10000, K=3x10000) and wgrad
(M=N=10000, K=1152, 3 taps as grid.z)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, _ = make_net()
dev = "cuda"
def bench(name, fn, flops, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s")
Nm = 10000
for mb in (8, 32):
    Mp = mb * 36
    A = torch.randn(Mp, Nm, device=dev).bfloat16()
    W = torch.randn(3, Nm, Nm, device=dev).bfloat16()
    U = torch.empty(Mp, Nm, device=dev)
    bias = torch.zeros(Nm, device=dev)
    for bn in (128, 256, 208, 0):
        net.gemm_bn = bn
        bench(f"fwd conv mb={mb} bn={bn}", lambda: net._gemm(A, Mp, Nm, Nm, W, Nm, Nm, Nm, Nm * Nm, 3, U, Nm, bias, 0, Mp, Nm, Nm, 3, (-1, 0, 1), (0, 0, 0), (0, 0, 0), (0, 1, 2)), 2.0 * Mp * Nm * Nm * 3)
    ld = (Mp + 7) // 8 * 8
    dUT = torch.randn(Nm, ld, device=dev).bfloat16()
    AT3 = torch.randn(3, Nm, ld, device=dev).bfloat16()
    dW = torch.zeros(3, Nm, Nm, device=dev)
    for bn in (128, 256, 208, 0):
        net.gemm_bn = bn
        bench(f"wgrad mb={mb} bn={bn}", lambda: net._gemm(dUT, Nm, Mp, ld, AT3, Nm, Mp, ld, Nm * ld, 3, dW, Nm, None, 1, Nm, Nm, Mp, 1, (0,), (0,), (0,), (0,), nz=3, z_b_tap_step=1, z_c_stride=Nm * Nm), 2.0 * Mp * Nm * Nm * 3)
