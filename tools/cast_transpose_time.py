"""Timing of dq_cast_transpose (bf16 operand refresh) on one 10000 x 10000 tap: python tools/cast_transpose_time.py"""
import torch, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffusion-deconvolution-dia-msms-data_b200"))
from dquartic import _native as N
x = torch.randn(10000, 10000, device="cuda")
o = torch.empty(10000, 10000, dtype=torch.bfloat16, device="cuda"); ot = torch.empty_like(o)
for _ in range(3): N.call("dq_cast_transpose", x, o, ot, 10000, 10000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): N.call("dq_cast_transpose", x, o, ot, 10000, 10000)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
print(f"cast_transpose 10000x10000: {t*1000:.0f} us = {0.8e9/t/1e9*1e3/1e3:.2f} TB/s")
