"""Timing of LA fwd / bwd at one shape: python tools/la_time.py C L samples"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, P = make_net()
net._ensure_grads()
C, L, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pre = {4: "downs.0.2", 8: "downs.2.2", 12: "downs.4.2", 16: "downs.6.2"}[C]
R = S * 34
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
print(f"{os.environ.get('DQ_B200_LIB','main')[-12:]} C={C} L={L} R={R}: fwd {e[0].elapsed_time(e[1])/5:.3f} ms  bwd {e[1].elapsed_time(e[2])/5:.3f} ms")
import ctypes
from dquartic import _native
_o = (ctypes.c_uint * 6)()
_native.lib().dq_la_tc_last_error.argtypes = [ctypes.POINTER(ctypes.c_uint)]
if _native.lib().dq_la_tc_last_error(_o): print("  !! tc pipeline timeout", [hex(v) for v in _o])
