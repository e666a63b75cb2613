"""Bring-up diagnostic: runs GEMM variants in separate processes so a trap in one does not hide the others."""
import subprocess
import sys

CASES = {
    "fwd3": dict(M=72, N=80, K=80, taps=3, a_row=(-1, 0, 1), bk=(0, 0, 0), btap=(0, 1, 2), nz=1, zk=0),
    "dgrad3": dict(M=72, N=80, K=80, taps=3, a_row=(1, 0, -1), bk=(0, 0, 0), btap=(0, 1, 2), nz=1, zk=0),
    "wgrad_k0": dict(M=80, N=80, K=72, taps=1, a_row=(0,), bk=(0,), btap=(0,), nz=1, zk=0),
    "wgrad_km8": dict(M=80, N=80, K=72, taps=1, a_row=(0,), bk=(-8,), btap=(0,), nz=1, zk=0),
    "wgrad_k8": dict(M=80, N=80, K=72, taps=1, a_row=(0,), bk=(8,), btap=(0,), nz=1, zk=0),
}


def run_case(name):
    import torch
    sys.path.insert(0, "tests")
    from _util import make_net
    from dquartic import _native as N

    c = CASES[name]
    net, _ = make_net()
    M, Nn, K = c["M"], c["N"], c["K"]
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g).bfloat16().cuda()
    nt = max(c["btap"]) + 1
    B = torch.randn(nt, Nn, K, generator=g).bfloat16().cuda()
    C = torch.zeros(c["nz"], M, Nn, device="cuda")
    net._gemm(A, M, K, K, B, Nn, K, K, Nn * K, nt, C, Nn, None, 0, M, Nn, K, c["taps"], c["a_row"], (0,) * c["taps"],
              c["bk"], c["btap"], nz=c["nz"], z_b_koff_step=c["zk"], z_c_stride=M * Nn)
    torch.cuda.synchronize()
    # reference
    Af, Bf = A.float().cpu(), B.float().cpu()
    ok = True
    for z in range(c["nz"]):
        ref = torch.zeros(M, Nn)
        for t in range(c["taps"]):
            As = torch.zeros_like(Af)
            o = c["a_row"][t]
            if o == 0: As = Af
            elif o < 0: As[-o:] = Af[:o]
            else: As[:-o] = Af[o:]
            Bs = torch.zeros(Nn, K)
            ko = c["bk"][t] + z * c["zk"]
            if ko == 0: Bs = Bf[c["btap"][t]]
            elif ko < 0: Bs[:, :ko] = Bf[c["btap"][t]][:, -ko:]
            else: Bs[:, ko:] = Bf[c["btap"][t]][:, :-ko]
            # B'[n][k] = B[n][k + ko]  -> shift left by ko
            Bs = torch.zeros(Nn, K)
            src = Bf[c["btap"][t]]
            if ko == 0: Bs = src
            elif ko > 0: Bs[:, :-ko] = src[:, ko:]
            else: Bs[:, -ko:] = src[:, :ko]
            ref += As @ Bs.T
        err = float((C[z].cpu() - ref).abs().max() / ref.abs().max())
        ok = ok and err < 1e-4
        print(name, "z", z, "rel err", err)
    print(name, "OK" if ok else "MISMATCH", "gemm_err", N.gemm_last_error())


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for name in CASES:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
            tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
            print(f"== {name}: rc={r.returncode}")
            for l in tail:
                print("   ", l[:200])
