"""GPU: each kernel family through the C-ABI against the oracle (torch fp32 on CPU, autograd for backward).
fp32 paths: rel <= 1e-4 of the tensor's max; bf16 tensor-core paths: rel <= 2e-2 (BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import TINY, make_net, rel_err

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4
TF32_TOL = 5e-3   # linear attention: mma.sync TF32 operands (10-bit mantissa), fp32 accumulate
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def ctx():
    import dquartic_oracle as O
    from dquartic import _native as N

    net, P = make_net()
    return dict(net=net, P=P, O=O, N=N)


def _time_setup(ctx, b, times):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    t = torch.tensor(times, dtype=torch.long)
    temb = O.time_mlp(P, t, TINY["dim"])
    net._ensure_grads()
    net._gflat.zero_()
    net._refresh_bf16()
    tp = net._time_path_fwd(t.cuda(), b, True)
    net._dSS = torch.zeros(b, net.ss_total, device="cuda")
    return t, temb, tp


def test_time_path_and_scale_shift(ctx):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    b = 3
    t, temb, tp = _time_setup(ctx, b, [0, 500, 999])
    assert rel_err(tp[3], temb) < FP32_TOL
    pre = "downs.2.1"
    ss_ref = F.linear(F.silu(temb), P[pre + ".mlp.1.weight"], P[pre + ".mlp.1.bias"])
    o = net.ss_off[pre + ".mlp.1"]
    assert rel_err(net._SS[:, o:o + ss_ref.shape[1]], ss_ref) < FP32_TOL


@pytest.mark.parametrize("pre,c1,c2,L", [("downs.0.0", 4, 0, 320), ("downs.3.1", 8, 0, 40), ("ups.0.0", 16, 16, 5),
                                         ("ups.3.1", 12, 8, 40), ("final_res_block", 4, 4, 320), ("ups.5.0", 8, 4, 160),
                                         ("downs.0.0", 4, 0, 1300), ("ups.6.1", 4, 4, 1030), ("downs.2.0", 8, 0, 2052),
                                         ("ups.1.0", 16, 12, 515), ("ups.6.0", 4, 4, 2052), ("ups.4.0", 8, 8, 1300),
                                         ("ups.5.1", 8, 4, 1028), ("ups.2.0", 12, 12, 1300), ("downs.4.0", 12, 0, 640),
                                         ("downs.6.1", 16, 0, 625), ("downs.5.0", 12, 0, 1250), ("ups.3.0", 12, 8, 1300),
                                         ("ups.0.1", 16, 16, 1252)])
def test_resnet_block_fwd_bwd(ctx, pre, c1, c2, L):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    b, rt = 2, 5
    R = b * rt
    t, temb, tp = _time_setup(ctx, b, [3, 700])
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(R, c1, L, generator=g)
    x2 = torch.randn(R, c2, L, generator=g) if c2 else None
    dout_shape_c = P[pre + ".block1.proj.weight"].shape[0]
    dout = torch.randn(R, dout_shape_c, L, generator=g)
    # oracle with autograd
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre + ".")}
    x1r = x1.clone().requires_grad_(True)
    x2r = x2.clone().requires_grad_(True) if c2 else None
    tr = temb.clone().requires_grad_(True)
    xin = torch.cat((x1r, x2r), 1) if c2 else x1r
    ref = O.resnet_block(Pg, pre, xin, tr, rt)
    ref.backward(dout)
    out, saved = net._resnet_fwd(pre, x1.cuda(), x2.cuda() if c2 else None, rt, True)
    # the pipelined kernels at 8 / 12 / 16 channels contract on the tensor cores (TF32 operands, fp32 accumulate)
    tol = TF32_TOL if (dout_shape_c >= 8 and L >= 128) else FP32_TOL
    assert rel_err(out, ref) < tol
    for got, name in zip(saved[2:], ("u1", "h1", "u2")):   # tensors saved for backward: finite everywhere
        assert torch.isfinite(got).all(), name
    dx1, dx2 = net._resnet_bwd(pre, saved, dout.cuda(), rt)
    assert rel_err(dx1, x1r.grad) < tol
    if c2:
        assert rel_err(dx2, x2r.grad) < tol
    for k, v in Pg.items():
        if k.endswith("mlp.1.weight") or k.endswith("mlp.1.bias"):
            continue
        assert rel_err(net._params[k].grad, v.grad) < 2 * tol, k
    # d scale/shift -> compare through the Linear's bias gradient (= sum over samples of dSS)
    o = net.ss_off[pre + ".mlp.1"]
    n = Pg[pre + ".mlp.1.bias"].shape[0]
    assert rel_err(net._dSS[:, o:o + n].sum(0), Pg[pre + ".mlp.1.bias"].grad) < 2 * tol


@pytest.mark.parametrize("pre,C,L", [("downs.0.2", 4, 320), ("downs.2.2", 8, 80), ("ups.0.2", 16, 5),
                                     ("ups.3.2", 12, 40), ("downs.1.2", 4, 4500)])
def test_linear_attention_fwd_bwd(ctx, pre, C, L):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    R = 6
    net._ensure_grads()
    net._gflat.zero_()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(R, C, L, generator=g) * 1.5
    dres = torch.randn(R, C, L, generator=g)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre + ".")}
    xr = x.clone().requires_grad_(True)
    ref = O.linear_attention(Pg, pre, xr)
    ref.backward(dres)
    out, saved = net._la_fwd(pre, x.cuda(), True)
    assert rel_err(out, ref) < TF32_TOL
    assert rel_err(out - x.cuda(), ref - x) < TF32_TOL  # the attention branch itself, not hidden by the residual
    dx = net._la_bwd(pre, saved, dres.cuda())
    assert rel_err(dx, xr.grad) < TF32_TOL
    for k, v in Pg.items():
        assert rel_err(net._params[k].grad, v.grad) < TF32_TOL, k


@pytest.mark.parametrize("L", [80, 1040, 1250])
@pytest.mark.parametrize("mode", ["down", "up", "last"])
def test_resample_convs(ctx, mode, L):
    net, P = ctx["net"], ctx["P"]
    net._ensure_grads()
    net._gflat.zero_()
    R = 7
    g = torch.Generator().manual_seed(3)
    if mode == "down":
        w, bn, K, s, pad, up = "downs.2.3.weight", "downs.2.3.bias", 4, 2, 1, 1
    elif mode == "up":
        w, bn, K, s, pad, up = "ups.2.3.1.weight", "ups.2.3.1.bias", 3, 1, 1, 2
    else:
        w, bn, K, s, pad, up = "downs.6.3.weight", "downs.6.3.bias", 3, 1, 1, 1
    cin = P[w].shape[1]
    x = torch.randn(R, cin, L, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = P[w].clone().requires_grad_(True), P[bn].clone().requires_grad_(True)
    xi = F.interpolate(xr, scale_factor=2, mode="nearest") if up == 2 else xr
    ref = F.conv1d(xi, wr, br, stride=s, padding=pad)
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    y, _ = net._conv_fwd(x.cuda(), None, w, bn, K, s, pad, up, ref.shape[2], rps=1)
    assert rel_err(y, ref) < FP32_TOL
    dx, _ = net._conv_bwd(dy.cuda(), x.cuda(), None, w, bn, K, s, pad, up, rps=1)
    assert rel_err(dx, xr.grad) < FP32_TOL
    assert rel_err(net._params[w].grad, wr.grad) < 2 * FP32_TOL
    assert rel_err(net._params[bn].grad, br.grad) < 2 * FP32_TOL


@pytest.mark.parametrize("M,Nn,K,taps", [(128, 128, 64, 1), (200, 80, 80, 3), (70, 256, 10000, 1), (300, 1000, 520, 3)])
def test_tcgen05_gemm_against_torch(ctx, M, Nn, K, taps):
    """C = sum_t A[m + t - 1] . B_t^T on bf16-rounded operands; fp32 accumulate -> tight tolerance."""
    net, N = ctx["net"], ctx["N"]
    g = torch.Generator().manual_seed(4)
    A = torch.randn(M, K, generator=g).bfloat16()
    B = (torch.randn(taps, Nn, K, generator=g) / math.sqrt(K * taps)).bfloat16()
    bias = torch.randn(Nn, generator=g)
    Af, Bf = A.float(), B.float()
    ref = torch.zeros(M, Nn)
    offs = (-1, 0, 1) if taps == 3 else (0,)
    for t, o in enumerate(offs):
        sh = torch.zeros_like(Af)
        if o == 0:
            sh = Af
        elif o == -1:
            sh[1:] = Af[:-1]
        else:
            sh[:-1] = Af[1:]
        ref += sh @ Bf[t].T
    ref += bias
    C = torch.empty(M, Nn, device="cuda")
    for bn in (128, 256, 208, 0):
        net.gemm_bn = bn
        C.fill_(float("nan"))
        net._gemm(A.cuda(), M, K, K, B.cuda(), Nn, K, K, Nn * K, taps, C, Nn, bias.cuda(), 0, M, Nn, K, taps, offs,
                  (0,) * taps, (0,) * taps, tuple(range(taps)))
        torch.cuda.synchronize()
        assert N.gemm_last_error() == 0
        assert rel_err(C, ref) < 2e-5, bn
        # accumulate mode
        net._gemm(A.cuda(), M, K, K, B.cuda(), Nn, K, K, Nn * K, taps, C, Nn, None, 1, M, Nn, K, taps, offs,
                  (0,) * taps, (0,) * taps, tuple(range(taps)))
        assert rel_err(C, 2 * ref - bias) < 2e-5, bn
    net.gemm_bn = 128


def test_mid_block_fwd_bwd(ctx):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    b, rt = 2, 34
    Nm = net.mid_channels
    t, temb, tp = _time_setup(ctx, b, [10, 900])
    g = torch.Generator().manual_seed(5)
    X = torch.randn(b * rt, Nm, generator=g)
    dOut = torch.randn(b * rt, Nm, generator=g)
    pre = "mid_block1"
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items() if k.startswith(pre + ".")}
    Xr = X.clone().requires_grad_(True)
    xin = Xr.view(b, rt, Nm).permute(0, 2, 1)  # (b, N, rt)
    ref = O.resnet_block(Pg, pre, xin, temb, 1).permute(0, 2, 1).reshape(b * rt, Nm)
    ref.backward(dOut)
    out, saved = net._mid_block_fwd(pre, X.cuda(), b, rt, True)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < BF16_TOL
    dX = net._mid_block_bwd(pre, saved, dOut.cuda(), b, rt)
    assert rel_err(dX, Xr.grad) < BF16_TOL
    for k, v in Pg.items():
        if "mlp" in k:
            continue
        assert rel_err(net._params[k].grad, v.grad) < BF16_TOL, k


def test_mid_attention_fwd_bwd(ctx):
    net, O, P = ctx["net"], ctx["O"], ctx["P"]
    b, rt = 2, 34
    Nm = net.mid_channels
    net._ensure_grads()
    net._gflat.zero_()
    net._refresh_bf16()
    g = torch.Generator().manual_seed(6)
    X = torch.randn(b * rt, Nm, generator=g)
    cond = torch.randn(b, 8, rt, generator=g)
    dOut = torch.randn(b * rt, Nm, generator=g)
    Pg = {k: (v.clone().requires_grad_(True) if not k.endswith("freqs") else v.clone()) for k, v in P.items()
          if k.startswith("mid_attn.")}
    Xr = X.clone().requires_grad_(True)
    cr = cond.clone().requires_grad_(True)
    ref = O.mid_attention(Pg, Xr.view(b, rt, Nm).permute(0, 2, 1), cr).permute(0, 2, 1).reshape(b * rt, Nm)
    ref.backward(dOut)
    cond_nlc = cond.permute(0, 2, 1).reshape(b * rt, 8).contiguous().cuda()
    out, saved = net._mid_attn_fwd(X.cuda(), cond_nlc, b, rt, True)
    assert rel_err(out, ref) < BF16_TOL
    dX, dcond = net._mid_attn_bwd(saved, dOut.cuda(), b, rt)
    assert rel_err(dX, Xr.grad) < BF16_TOL
    assert rel_err(dcond.view(b, rt, 8).permute(0, 2, 1), cr.grad) < BF16_TOL
    for k, v in Pg.items():
        if k.endswith("freqs"):
            continue
        assert rel_err(net._params[k].grad, v.grad) < BF16_TOL, k


def test_scheduler_kernels_bit_exact(ctx):
    O, N = ctx["O"], ctx["N"]
    from dquartic.model.model import DDIMDiffusionModel

    net = ctx["net"]
    d = DDIMDiffusionModel(net, device="cuda")
    _, _, ab = O.schedule_tables(1000, "cosine")
    assert torch.equal(d.alpha_bars.cpu(), ab)
    g = torch.Generator().manual_seed(7)
    x0 = torch.rand(3, 5, 321, generator=g)
    noise = torch.randn(3, 5, 321, generator=g)
    t = torch.tensor([0, 499, 999])
    ref = O.q_sample(ab, O.normalize(x0), t, noise)
    got = d._q_sample_fused(x0.cuda(), t.cuda(), noise.cuda())
    assert torch.equal(got.cpu(), ref)
    assert torch.equal(d.q_sample(O.normalize(x0).cuda(), t.cuda(), noise.cuda()).cpu(), ref)
    eps = torch.randn(3, 5, 321, generator=g)
    for tt in (999, 978, 1, 0):
        refp = O.ddim_update(ab, noise, eps, tt)
        sa, s1m, sap, s1mp = d._step_coefs(tt)
        out = torch.empty_like(noise).cuda()
        N.call("dq_ddim_step", noise.cuda(), eps.cuda(), out, sa, s1m, sap, s1mp, 1 if tt == 0 else 0, noise.numel())
        assert torch.equal(out.cpu(), refp), tt
    xo, pn = torch.empty_like(x0).cuda(), torch.empty_like(x0).cuda()
    cn = O.normalize(x0)
    N.call("dq_sample_finalize", noise.cuda(), cn.cuda(), xo, pn, x0.numel())
    assert torch.equal(xo.cpu(), O.unnormalize(noise))
    assert torch.equal(pn.cpu(), O.unnormalize(cn) - O.unnormalize(noise))


def test_fused_adamw_and_clip(ctx):
    O = ctx["O"]
    from dquartic.model.model_interface import FusedAdamW

    net, P = make_net(seed=5)
    opt = FusedAdamW(net, lr=1e-3)
    g = torch.Generator().manual_seed(8)
    grads = {k: torch.randn(v.shape, generator=g) * 3 for k, v in P.items() if not k.endswith("freqs")}
    p, m, v = {k: P[k].clone() for k in grads}, {k: torch.zeros_like(P[k]) for k in grads}, {k: torch.zeros_like(P[k]) for k in grads}
    for step in (1, 2):
        net.zero_grad()
        net._ensure_grads()
        for k in grads:
            net._params[k].grad.copy_(grads[k] * step)
        clipped, total = O.clip_grad_norm([grads[k] * step for k in grads])
        opt.step(max_grad_norm=10.0)
        assert abs(float(opt.last_grad_norm) - float(total)) < 1e-5 * float(total)
        for k, gc in zip(grads, clipped):
            p[k], m[k], v[k] = O.adamw_step(p[k], gc, m[k], v[k], step, 1e-3)
    for k in ("init_conv.weight", "mid_block1.block1.proj.weight", "downs.3.2.fn.fn.to_qkv.weight", "final_conv.bias"):
        assert rel_err(net._params[k], p[k]) < 1e-6, k
    assert torch.equal(net._params["mid_attn.fn.fn.rotary_emb.freqs"].cpu(), P["mid_attn.fn.fn.rotary_emb.freqs"])


def test_multiplex_bit_exact_against_reference_golden(ctx):
    import random

    from _util import golden
    from dquartic.utils.data_loader import DeviceBatchLoader, DIAMSDataset

    g = golden("data.npz")
    import tempfile, os
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "ms2.npy"), g["ms2_pool"])
        np.save(os.path.join(td, "ms1.npy"), g["ms1_pool"])
        for pool in ("hbm", "pinned"):
            ds = DIAMSDataset(ms2_file=os.path.join(td, "ms2.npy"), ms1_file=os.path.join(td, "ms1.npy"), normalize="minmax")
            random.seed(1234)
            dl = DeviceBatchLoader(ds, 6, "cuda", pool=pool)
            x0, m1, other, m2, cond = dl.make_batch(dl.draw(6), want_cond=True)
            for j in range(6):
                assert np.array_equal(x0[j].cpu().numpy(), g[f"item{j}:ms2_1"]), (pool, j)
                assert np.array_equal(m1[j].cpu().numpy(), g[f"item{j}:ms1_1"])
                assert np.array_equal(other[j].cpu().numpy(), g[f"item{j}:ms2_2"])
                assert np.array_equal(m2[j].cpu().numpy(), g[f"item{j}:ms1_2"])
                assert np.array_equal(cond[j].cpu().numpy(), g[f"item{j}:mix"])
        # float32 pool keeps numpy's float32 arithmetic
        np.save(os.path.join(td, "ms2f.npy"), g["f32_pool_ms2"])
        np.save(os.path.join(td, "ms1f.npy"), g["f32_pool_ms1"])
        ds = DIAMSDataset(ms2_file=os.path.join(td, "ms2f.npy"), ms1_file=os.path.join(td, "ms1f.npy"), normalize="minmax")
        random.seed(99)
        dl = DeviceBatchLoader(ds, 1, "cuda")
        x0, m1, other, m2 = dl.make_batch(dl.draw(1))
        assert np.array_equal(x0[0].cpu().numpy(), g["f32item:ms2_1"])
        assert np.array_equal(other[0].cpu().numpy(), g["f32item:ms2_2"])
        assert np.array_equal(m1[0].cpu().numpy(), g["f32item:ms1_1"])
        assert np.array_equal(m2[0].cpu().numpy(), g["f32item:ms1_2"])
