"""Debug: run LA fwd+bwd for a list of (C, L, R) with a sync after each call: python tools/la_dbg.py C:L:R ..."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _util import make_net
net, P = make_net()
net._ensure_grads()
pres = {4: "downs.0.2", 8: "downs.2.2", 12: "downs.4.2", 16: "downs.6.2"}
for spec in sys.argv[1:]:
    C, L, R = map(int, spec.split(":"))
    x = torch.randn(R, C, L, device="cuda"); dres = torch.randn_like(x)
    out, saved = net._la_fwd(pres[C], x, True)
    torch.cuda.synchronize()
    print(spec, "fwd ok", flush=True)
    for it in range(4):
        dx = net._la_bwd(pres[C], saved, dres)
        torch.cuda.synchronize()
        print(spec, "bwd iter", it, float(dx.abs().mean()), flush=True)
    print(spec, "bwd ok", float(dx.abs().mean()), flush=True)
import ctypes
from dquartic import _native
out = (ctypes.c_uint * 6)()
_native.lib().dq_la_tc_last_error.argtypes = [ctypes.POINTER(ctypes.c_uint)]
print("tc_err", _native.lib().dq_la_tc_last_error(out), [hex(v) for v in out])
