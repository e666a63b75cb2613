"""B200-native `UNet1d` denoiser — drop-in for the reference's `dquartic.model.unet1d.UNet1d`
(/root/reference/dquartic/model/unet1d.py:874-1166, `simple=True, conditional=True` — the only live path,
SURVEY.md findings 2/3).

Same constructor signature, same `forward(x, time, init_cond, attn_cond)`, same 396 state-dict names and shapes
(so reference checkpoints load), but:

* all parameters are views into ONE flat fp32 buffer (`_flat`), gradients into one flat buffer (`_gflat`): the
  optimizer, the grad-norm and the data-parallel all-reduce are single passes over contiguous memory;
* forward AND backward are explicit sequences of hand-written sm_100a kernels (csrc/*.cu through the C-ABI in
  include/dquartic_b200.h); autograd only sees one node (`_UNetFn`);
* it is batched: row r of sample i uses sample i's time embedding (the vmap of the b=1 reference);
* the 10 000-channel mid stage runs as bf16 tcgen05 GEMMs with fp32 accumulation on a padded [b][RT+2][N] layout.

There is no CPU path: calling forward on CPU tensors raises.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .. import _native as N

HEADS = 4
DIM_HEAD = 32
HD = HEADS * DIM_HEAD
ACT_NONE, ACT_SILU, ACT_GELU, ACT_SOFTPLUS = 0, 1, 2, 3


def exists(x):
    return x is not None


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) else d


# ------------------------------------------------------------------------------------------ parameter inventory
def param_specs(dim, dim_mults, channels, init_cond_channels, attn_cond_channels, downsample_dim):
    """Ordered name -> shape of the reference's state_dict (unet1d.py:940-1084)."""
    td = dim * 4
    dims = [dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    S = OrderedDict()

    def conv(name, co, ci, k, bias=True):
        S[name + ".weight"] = (co, ci, k)
        if bias:
            S[name + ".bias"] = (co,)

    def lin(name, o, i):
        S[name + ".weight"] = (o, i)
        S[name + ".bias"] = (o,)

    def resblock(name, ci, co):
        lin(name + ".mlp.1", 2 * co, td)
        conv(name + ".block1.proj", co, ci, 3)
        S[name + ".block1.norm.g"] = (1, co, 1)
        conv(name + ".block2.proj", co, co, 3)
        S[name + ".block2.norm.g"] = (1, co, 1)
        if ci != co:
            conv(name + ".res_conv", co, ci, 1)

    def linattn(name, c):
        S[name + ".fn.fn.to_qkv.weight"] = (3 * HD, c, 1)
        conv(name + ".fn.fn.to_out.0", c, HD, 1)
        S[name + ".fn.fn.to_out.1.g"] = (1, c, 1)
        S[name + ".fn.norm.g"] = (1, c, 1)

    conv("init_conv", dim, channels + init_cond_channels, 7)
    lin("time_mlp.1", td, dim)
    lin("time_mlp.3", td, td)
    lin("init_cond_proj.to_scale_shift.1", 2 * init_cond_channels, td)
    acd = dim * 2
    conv("attn_cond_proj.1.0", acd, attn_cond_channels, 7)
    conv("attn_cond_proj.1.2", acd, acd, 1)
    n = len(in_out)
    for i, (di, do) in enumerate(in_out):
        resblock(f"downs.{i}.0", di, di)
        resblock(f"downs.{i}.1", di, di)
        linattn(f"downs.{i}.2", di)
        conv(f"downs.{i}.3", do, di, 4 if i < n - 1 else 3)
    for j, (di, do) in enumerate(reversed(in_out)):
        resblock(f"ups.{j}.0", do + di, do)
        resblock(f"ups.{j}.1", do + di, do)
        linattn(f"ups.{j}.2", do)
        conv(f"ups.{j}.3.1" if j < n - 1 else f"ups.{j}.3", di, do, 3)
    dn = downsample_dim // (2 ** (len(dim_mults) - 1))
    cm = dims[-1] * dn
    resblock("mid_block1", cm, cm)
    S["mid_attn.fn.fn.rotary_emb.freqs"] = (DIM_HEAD // 4,)
    S["mid_attn.fn.fn.to_qv.weight"] = (2 * HD, cm, 1)
    S["mid_attn.fn.fn.to_k.weight"] = (HD, acd, 1)
    conv("mid_attn.fn.fn.to_out", cm, HD, 1)
    S["mid_attn.fn.norm.g"] = (1, cm, 1)
    resblock("mid_block2", cm, cm)
    resblock("final_res_block", dim * 2, dim)
    conv("final_conv", channels, dim, 1)
    return S


class _NS(nn.Module):
    """Name-space container: only exists so that parameters get the reference's dotted state-dict names."""


class _Tape:
    """Saved activations of one forward pass."""

    def __init__(self):
        self.d = {}


def _ceil8(n):
    return (n + 7) // 8 * 8


class UNet1d(nn.Module):
    def __init__(
        self,
        dim,
        init_dim=None,
        out_dim=None,
        dim_mults=(1, 2, 4, 8),
        channels=3,
        dropout=0.0,
        conditional=True,
        init_cond_channels=None,
        attn_cond_channels=None,
        attn_cond_init_dim=None,
        learned_variance=False,
        sinusoidal_pos_emb_theta=10000,
        attn_heads=4,
        attn_dim_head=32,
        tfer_dim_mult=620,
        tfer_depth=4,
        downsample_dim=40000,
        simple=True,
        pos_output_only=False,
        device=None,
    ):
        super().__init__()
        self._init_device = device  # B200 addition: build the flat buffer directly on this device
        # The reference's other branches are dead code (SURVEY.md findings 2, §2 "OUT OF SCOPE" rows): refuse loudly.
        if not simple:
            raise NotImplementedError("simple=False crashes in the reference (unet1d.py:822); only simple=True is built")
        if not conditional:
            raise NotImplementedError("only the conditional denoiser (MS2 mixture + MS1 conditioning) is built")
        if attn_heads != HEADS or attn_dim_head != DIM_HEAD:
            raise NotImplementedError("kernels are specialised for 4 heads x 32 (the reference defaults)")
        if channels != 1 or default(init_cond_channels, 0) != 1 or default(attn_cond_channels, 0) != 1:
            raise NotImplementedError("channels / init_cond_channels / attn_cond_channels must be 1 (config schema)")
        if exists(init_dim) or exists(out_dim) or exists(attn_cond_init_dim) or learned_variance:
            raise NotImplementedError("init_dim/out_dim/attn_cond_init_dim/learned_variance: defaults only")
        if dropout != 0.0:
            raise NotImplementedError("dropout is 0.0 everywhere in the reference config; the kernels have no RNG")
        if dim % 4 != 0 or dim * max(dim_mults) > 32:
            raise NotImplementedError("dim must be a multiple of 4 and dim*max(dim_mults) <= 32")
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        self.channels = channels
        self.conditional = conditional
        self.theta = sinusoidal_pos_emb_theta
        self.downsample_dim = downsample_dim
        self.downsampled_n = downsample_dim // (2 ** (len(dim_mults) - 1))
        self.out_dim = channels
        # reference unet1d.py:1084: Softplus head when pos_output_only (one extra elementwise kernel fwd and bwd)
        self.pos_output_only = bool(pos_output_only)
        self.final_act = nn.Softplus() if pos_output_only else nn.Identity()
        self.time_dim = dim * 4
        self.dims = [dim] + [dim * m for m in dim_mults]
        self.in_out = list(zip(self.dims[:-1], self.dims[1:]))
        self.mid_channels = self.dims[-1] * self.downsampled_n
        if self.mid_channels % 8 != 0:
            raise NotImplementedError("mid channel count must be a multiple of 8 (TMA row-stride alignment)")
        self.specs = param_specs(dim, self.dim_mults, channels, 1, 1, downsample_dim)
        self._build_layout()
        self._register()
        self.reset_parameters()
        self._gflat = None
        self._bf16 = {}
        self._bf16_version = -1
        self._manual_version = 0
        self.last_tape = None
        # data-parallel hook: called from backward with [(flat offset, numel)] of gradients that are final
        # (the four 300 M mid-conv weights + mid attention projections) while the down-path backward still runs
        self.grad_ready_callback = None
        self._final_microbatch = True
        self._wgrad_defer = None   # (padded rows before, padded rows total, is last micro-batch) or None
        self._wgrad_stash = {}

    # ---------------------------------------------------------------------------------------------- layout
    def _ss_producers(self):
        n = len(self.in_out)
        names = ["init_cond_proj.to_scale_shift.1"]
        for i in range(n):
            names += [f"downs.{i}.0.mlp.1", f"downs.{i}.1.mlp.1"]
        names += ["mid_block1.mlp.1", "mid_block2.mlp.1"]
        for j in range(n):
            names += [f"ups.{j}.0.mlp.1", f"ups.{j}.1.mlp.1"]
        names += ["final_res_block.mlp.1"]
        return names

    def _build_layout(self):
        """Flat buffer = [all scale/shift Linear weights | their biases | everything else (16-byte aligned) | freqs]."""
        off = 0
        self.offsets = {}
        self.ss_off = {}  # producer -> column offset inside SS (b, Ntot)
        col = 0
        for p in self._ss_producers():
            o, i = self.specs[p + ".weight"]
            self.offsets[p + ".weight"] = off
            self.ss_off[p] = col
            off += o * i
            col += o
        self.ss_total = col
        self.ss_w_off = 0
        self.ss_b_off = off
        for p in self._ss_producers():
            self.offsets[p + ".bias"] = off
            off += self.specs[p + ".bias"][0]
        for name, shape in self.specs.items():
            if name in self.offsets or name.endswith("rotary_emb.freqs"):
                continue
            off = (off + 3) // 4 * 4
            self.offsets[name] = off
            off += int(np.prod(shape))
        off = (off + 3) // 4 * 4
        self.n_trainable_flat = off  # AdamW / grad-norm range (padding included, it stays zero)
        self.offsets["mid_attn.fn.fn.rotary_emb.freqs"] = off
        off += DIM_HEAD // 4
        self.n_flat = off

    def _is_mid_conv(self, name):
        return name.startswith("mid_block") and name.endswith(".proj.weight")

    def _view(self, flat, name):
        shape = self.specs[name]
        n = int(np.prod(shape))
        v = flat[self.offsets[name]: self.offsets[name] + n]
        if self._is_mid_conv(name):  # stored tap-major [3][co][ci], exposed as the reference's (co, ci, 3)
            co, ci, k = shape
            return v.view(k, co, ci).permute(1, 2, 0)
        return v.view(shape)

    def _register(self):
        self._flat = torch.zeros(self.n_flat, dtype=torch.float32, device=self._init_device)
        self._params = OrderedDict()
        for name in self.specs:
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, _NS())
                mod = mod._modules[p]
            par = nn.Parameter(self._view(self._flat, name), requires_grad=not name.endswith("rotary_emb.freqs"))
            mod.register_parameter(parts[-1], par)
            self._params[name] = par

    def _repoint(self, flat):
        self._flat = flat
        for name, par in self._params.items():
            par.data = self._view(flat, name)
        if self._gflat is not None and self._gflat.device != flat.device:
            self._gflat = None
            for par in self._params.values():
                par.grad = None
        self._bf16 = {}
        self._bf16_version = -1

    def _apply(self, fn, recurse=True):
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise NotImplementedError("master parameters stay fp32 (the GEMM operands are cast to bf16 internally)")
        self._repoint(new_flat.contiguous())
        return self

    def reset_parameters(self):
        """PyTorch default init for Conv1d / Linear (kaiming_uniform a=sqrt(5): U(+-1/sqrt(fan_in)) for weight and
        bias), g = 1, rotary freqs = 1/10000^(2j/16) — the reference has no custom init."""
        with torch.no_grad():
            for name, shape in self.specs.items():
                n = int(np.prod(shape))
                v = self._flat[self.offsets[name]: self.offsets[name] + n]
                if name.endswith("rotary_emb.freqs"):
                    d = DIM_HEAD // 2
                    v.copy_(1.0 / (10000 ** (torch.arange(0, d, 2)[: d // 2].float() / d)))
                elif name.endswith(".g"):
                    v.fill_(1.0)
                else:
                    wname = name[:-5] + ".weight" if name.endswith(".bias") else name
                    fan_in = int(np.prod(self.specs[wname][1:]))
                    bound = 1.0 / math.sqrt(fan_in)
                    v.uniform_(-bound, bound)
        self._manual_version = getattr(self, "_manual_version", 0) + 1

    def _load_from_state_dict(self, *a, **k):  # pragma: no cover - containers handle their own params
        super()._load_from_state_dict(*a, **k)

    def state_dict(self, *args, **kwargs):
        self.sync_params()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        self.sync_params()
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self._manual_version += 1
        return out

    def mark_params_modified(self):
        """Call after writing parameters through raw pointers (the fused optimizer does)."""
        self._manual_version += 1

    # ---------------------------------------------------------------------------------------------- flat accessors
    def flat_params(self):
        self.sync_params()
        return self._flat

    def sync_params(self):
        """Make the current stream wait for parameter all-gathers the sharded optimizer left in flight
        (`FusedAdamW.step(defer_gather=True)`): they overlap the next step's down path, whose parameters are replicated;
        only the mid-stage weights arrive through them.  Called before anything reads those weights."""
        works = self.__dict__.get("_pending_param_works")
        if works:
            for w in works:
                w.wait()
            works.clear()

    def flat_grads(self):
        self._ensure_grads()
        return self._gflat

    def _ensure_grads(self):
        dev = self._flat.device
        if self._gflat is None or self._gflat.device != dev:
            self._gflat = torch.zeros(self.n_flat, dtype=torch.float32, device=dev)
            for name, par in self._params.items():
                if par.requires_grad:
                    par.grad = self._view(self._gflat, name)
            return
        for name, par in self._params.items():
            if par.requires_grad and par.grad is None:  # e.g. optimizer.zero_grad(set_to_none=True)
                g = self._view(self._gflat, name)
                g.zero_()
                par.grad = g

    def zero_grad(self, set_to_none=False):
        if self._gflat is not None:
            self._gflat.zero_()
            for name, par in self._params.items():
                if par.requires_grad and par.grad is None:
                    par.grad = self._view(self._gflat, name)

    def early_grad_ranges(self):
        names = [f"mid_block{i}.block{j}.proj.weight" for i in (1, 2) for j in (1, 2)]
        names += ["mid_attn.fn.fn.to_qv.weight", "mid_attn.fn.fn.to_out.weight"]
        return [(self.offsets[n], int(np.prod(self.specs[n]))) for n in names]

    def _w(self, name):
        """Contiguous flat slice of a parameter (memory order; mid convs are [3][co][ci])."""
        n = int(np.prod(self.specs[name]))
        return self._flat[self.offsets[name]: self.offsets[name] + n]

    def _gw(self, name):
        n = int(np.prod(self.specs[name]))
        return self._gflat[self.offsets[name]: self.offsets[name] + n]

    # ---------------------------------------------------------------------------------------------- bf16 operands
    def _bf16_sources(self):
        """Names of the fp32 parameters that have bf16 GEMM-operand copies."""
        return [f"{blk}.{bl}.proj.weight" for blk in ("mid_block1", "mid_block2") for bl in ("block1", "block2")] + \
            ["mid_attn.fn.fn.to_qv.weight", "mid_attn.fn.fn.to_out.weight"]

    def __deepcopy__(self, memo):
        """copy.deepcopy(net): the copy's Parameters must alias ITS flat buffer (a member-wise deep copy would leave
        them as independent tensors, and the kernels read the flat buffer)."""
        import copy
        self.sync_params()
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        skip = {"_flat", "_gflat", "_bf16", "_params", "_modules", "_parameters"}
        nn.Module.__init__(new)
        for k, v in self.__dict__.items():
            if k in skip:
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._init_device = self._flat.device
        new._register()
        with torch.no_grad():
            new._flat.copy_(self._flat)
        new._gflat = None
        new._bf16 = {}
        new._bf16_version = -1
        for name, par in self._params.items():
            new._params[name].requires_grad_(par.requires_grad)
        new.training = self.training
        return new

    def _refresh_bf16(self):
        # The parameters' OWN version counters are part of the key: after .to(device) / _apply the Parameters are
        # re-pointed with `par.data = view`, which gives them counters of their own, so an in-place update that does
        # not go through FusedAdamW / load_state_dict (torch.optim.*, nn.init, EMA copy_, p.mul_) bumps only those.
        self.sync_params()
        ver = (self._flat._version, self._manual_version, self._flat.data_ptr(),
               tuple(self._params[k]._version for k in self._bf16_sources()))
        if self._bf16_version == ver:
            return
        Nm = self.mid_channels
        dev = self._flat.device
        B = self._bf16

        def buf(key, shape):
            if key not in B or B[key].shape != torch.Size(shape):
                B[key] = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            return B[key]

        for blk in ("mid_block1", "mid_block2"):
            for bl in ("block1", "block2"):
                name = f"{blk}.{bl}.proj.weight"
                w = self._w(name).view(3, Nm, Nm)
                wb = buf(name, (3, Nm, Nm))
                wt = buf(name + ".T", (3, Nm, Nm))
                for t in range(3):
                    N.call("dq_cast_transpose", w[t], wb[t], wt[t], Nm, Nm)
        wqv = self._w("mid_attn.fn.fn.to_qv.weight").view(2 * HD, Nm)
        N.call("dq_cast_transpose", wqv, buf("wqv", (2 * HD, Nm)), buf("wqv.T", (Nm, 2 * HD)), 2 * HD, Nm)
        wo = self._w("mid_attn.fn.fn.to_out.weight").view(Nm, HD)
        N.call("dq_cast_transpose", wo, buf("wout", (Nm, HD)), buf("wout.T", (HD, Nm)), Nm, HD)
        self._bf16_version = ver

    # ---------------------------------------------------------------------------------------------- primitive ops
    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self._flat.device)

    def _zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self._flat.device)

    def _conv_fwd(self, x1, x2, wname, bname, K, stride, pad, up, Lout, g=None, ss=None, act=ACT_NONE, res=None,
                  save_u=False, in_ss=None, rps=1):
        R, c1, Lin = x1.shape
        c2 = x2.shape[1] if x2 is not None else 0
        cout = self.specs[wname][0]
        y = self._empty(R, cout, Lout)
        u = self._empty(R, cout, Lout) if save_u else None
        ssp, sss = (None, 0) if ss is None else (self._SS[:, ss:], self.ss_total)
        issp, isss = (None, 0) if in_ss is None else (self._SS[:, in_ss:], self.ss_total)
        N.call("dq_conv1d_fwd", x1, c1, x2, c2, _off_ptr(issp), isss, self._w(wname),
               self._w(bname) if bname else None, cout, K, stride, pad, up, self._w(g) if g else None,
               _off_ptr(ssp), sss, act, res, u, y, R, Lin, Lout, rps)
        return y, u

    def _conv_bwd(self, du, x1, x2, wname, bname, K, stride, pad, up, need_dx1=True, need_dx2=True, dx1=None,
                  dx2=None, in_ss=None, rps=1):
        """Accumulates dW/db; returns (dx1, dx2).  If dx1/dx2 tensors are passed the result is ACCUMULATED into them."""
        if K in (1, 3) and stride == 1 and up == 1 and in_ss is None:
            return self._conv_bwd_fused(du, None, None, None, ACT_NONE, x1, x2, wname, bname, K, need_dx1, need_dx2,
                                        dx1, dx2, None, rps)
        if K == 3 and stride == 1 and up == 2 and x2 is None and in_ss is None:
            return self._upconv_bwd(du, x1, wname, bname, need_dx1, dx1, rps), None
        if K == 4 and stride == 2 and pad == 1 and up == 1 and x2 is None and in_ss is None and x1.shape[2] % 2 == 0:
            return self._downconv_bwd(du, x1, wname, bname, need_dx1, dx1, rps), None
        R, cout, Lout = du.shape
        c1, Lin = x1.shape[1], x1.shape[2]
        c2 = x2.shape[1] if x2 is not None else 0
        issp, isss = (None, 0) if in_ss is None else (self._SS[:, in_ss:], self.ss_total)
        N.call("dq_conv1d_bwd_weight", du, x1, c1, x2, c2, _off_ptr(issp), isss, self._gw(wname),
               self._gw(bname) if bname else None, cout, K, stride, pad, up, R, Lin, Lout, rps)
        acc1 = 1 if dx1 is not None else 0
        acc2 = 1 if dx2 is not None else 0
        if need_dx1 and dx1 is None:
            dx1 = self._empty(R, c1, Lin)
        if need_dx2 and c2 and dx2 is None:
            dx2 = self._empty(R, c2, Lin)
        if (need_dx1 or (need_dx2 and c2)):
            N.call("dq_conv1d_bwd_data", du, self._w(wname), dx1 if need_dx1 else None, c1, acc1,
                   dx2 if (need_dx2 and c2) else None, c2, acc2, cout, K, stride, pad, up, R, Lin, Lout)
        return dx1, dx2

    def _upconv_bwd(self, du, x, wname, bname, need_dx, dx, rps):
        """Backward of Upsample = nearest x2 + Conv1d(k3) (unet1d.py:93-96) through the fused stride-1 kernel: the
        upsampled input is rebuilt (never saved), d x_up is folded back pairwise."""
        R, c, Lh = x.shape
        if not getattr(self, "_force_unfused_upconv", False):   # (test switch: the three-pass composition below)
            acc = 1 if (need_dx and dx is not None) else 0
            out = (dx if dx is not None else self._empty(R, c, Lh)) if need_dx else None
            rc = N.call("dq_upconv_bwd_fused", du, x, self._w(wname), out, acc, self._gw(wname),
                        self._gw(bname) if bname else None, du.shape[1], c, R, Lh, rps, allow=(1,))
            if rc == 0:
                return out
        xup = self._empty(R, c, 2 * Lh)
        N.call("dq_upsample2x", x, xup, x.numel())
        dxup, _ = self._conv_bwd_fused(du, None, None, None, ACT_NONE, xup, None, wname, bname, 3, need_dx1=need_dx, rps=rps)
        if not need_dx:
            return None
        acc = 1 if dx is not None else 0
        if dx is None:
            dx = self._empty(R, c, Lh)
        N.call("dq_fold2x", dxup, dx, x.numel(), acc)
        return dx

    def _downconv_bwd(self, du, x, wname, bname, need_dx, dx, rps):
        """Backward of Downsample = Conv1d(k4, s2, p1) (unet1d.py:110) as a k3/s1 conv over the space-to-depth input
        [x_even; x_odd] with re-packed weights (csrc/small.cu)."""
        R, c, L = x.shape
        cout = du.shape[1]
        Lh = L // 2
        if not getattr(self, "_force_unfused_downconv", False):   # (test switch: the re-indexing composition below)
            acc = 1 if (need_dx and dx is not None) else 0
            out = (dx if dx is not None else self._empty(R, c, L)) if need_dx else None
            rc = N.call("dq_downconv_bwd_fused", du, x, self._w(wname), out, acc, self._gw(wname),
                        self._gw(bname) if bname else None, cout, c, R, L, rps, allow=(1,))
            if rc == 0:
                return out
        xs = self._empty(R, 2 * c, Lh)
        N.call("dq_s2d", x, xs, R, c, L)
        w3 = self._empty(cout, 2 * c, 3)
        dw3 = self._zeros(cout, 2 * c, 3)
        N.call("dq_down_w", self._w(wname), w3, cout, c, 0)
        dxs = self._empty(R, 2 * c, Lh) if need_dx else None
        N.call("dq_conv_bwd_fused", du, None, None, None, 0, ACT_NONE, xs, 2 * c, None, 0, w3, None, dxs, 0, None, 0,
               dw3, self._gw(bname) if bname else None, None, None, cout, 3, R, Lh, rps)
        N.call("dq_down_w", self._gw(wname), dw3, cout, c, 1)
        if not need_dx:
            return None
        acc = 1 if dx is not None else 0
        if dx is None:
            dx = self._empty(R, c, L)
        N.call("dq_d2s", dxs, dx, R, c, L, acc)
        return dx

    def _conv_bwd_fused(self, dy, u, g, ss, act, x1, x2, wname, bname, K, need_dx1=True, need_dx2=True, dx1=None,
                        dx2=None, dadd=None, rps=1):
        """One-pass backward of a stride-1 conv (K in {1, 3}) and, if `u` is given, of its RMSNorm / scale-shift /
        activation epilogue: dx1/dx2 (accumulated into the passed tensors, `dadd` added to dx1), dW, db, dg, dSS."""
        R, cout, L = dy.shape
        c1 = x1.shape[1]
        c2 = x2.shape[1] if x2 is not None else 0
        acc1 = 1 if dx1 is not None else 0
        acc2 = 1 if dx2 is not None else 0
        if need_dx1 and dx1 is None:
            dx1 = self._empty(R, c1, L)
        if need_dx2 and c2 and dx2 is None:
            dx2 = self._empty(R, c2, L)
        ssp, sss = (None, 0) if ss is None else (self._SS[:, ss:], self.ss_total)
        dssp = None if ss is None else self._dSS[:, ss:]
        N.call("dq_conv_bwd_fused", dy, u, self._w(g) if g else None, _off_ptr(ssp), sss, act, x1, c1, x2, c2,
               self._w(wname), dadd, dx1 if need_dx1 else None, acc1, dx2 if (need_dx2 and c2) else None, acc2,
               self._gw(wname), self._gw(bname) if bname else None, self._gw(g) if g else None, _off_ptr(dssp),
               cout, K, R, L, rps)
        return dx1, dx2

    def _block_bwd(self, dy, u, g, ss, act, rps):
        R, C, L = dy.shape
        du = self._empty(R, C, L)
        ssp, sss = (None, 0) if ss is None else (self._SS[:, ss:], self.ss_total)
        dssp = None if ss is None else self._dSS[:, ss:]
        N.call("dq_block_bwd", dy, u, self._w(g) if g else None, _off_ptr(ssp), sss, act, du,
               self._gw(g) if g else None, _off_ptr(dssp), C, R, L, rps)
        return du

    # ---------------------------------------------------------------------------------------------- composite ops
    def _resnet_fwd(self, pre, x1, x2, rps, save):
        R, c1, L = x1.shape
        c2 = x2.shape[1] if x2 is not None else 0
        cout = self.specs[pre + ".block1.proj.weight"][0]
        has_res = (pre + ".res_conv.weight") in self.specs
        # one-pass kernel: Block1 -> Block2 + skip with the h1 tile in shared memory (csrc/conv_fused.cu)
        out = self._empty(R, cout, L)
        u1 = self._empty(R, cout, L) if save else None
        h1 = self._empty(R, cout, L) if save else None
        u2 = self._empty(R, cout, L) if save else None
        sso = self.ss_off[pre + ".mlp.1"]
        rc = N.call("dq_resblock_fwd", x1, c1, x2, c2, self._w(pre + ".block1.proj.weight"),
                    self._w(pre + ".block1.proj.bias"), self._w(pre + ".block1.norm.g"), _off_ptr(self._SS[:, sso:]),
                    self.ss_total, self._w(pre + ".block2.proj.weight"), self._w(pre + ".block2.proj.bias"),
                    self._w(pre + ".block2.norm.g"), self._w(pre + ".res_conv.weight") if has_res else None,
                    self._w(pre + ".res_conv.bias") if has_res else None, u1, h1, u2, out, cout, R, L, rps, allow=(1,))
        if rc == 0:
            return out, (x1, x2, u1, h1, u2)
        del out, u1, h1, u2
        h1, u1 = self._conv_fwd(x1, x2, pre + ".block1.proj.weight", pre + ".block1.proj.bias", 3, 1, 1, 1, L,
                                g=pre + ".block1.norm.g", ss=self.ss_off[pre + ".mlp.1"], act=ACT_SILU, save_u=save,
                                rps=rps)
        if (pre + ".res_conv.weight") in self.specs:
            res, _ = self._conv_fwd(x1, x2, pre + ".res_conv.weight", pre + ".res_conv.bias", 1, 1, 0, 1, L, rps=rps)
        else:
            res = x1
        out, u2 = self._conv_fwd(h1, None, pre + ".block2.proj.weight", pre + ".block2.proj.bias", 3, 1, 1, 1, L,
                                 g=pre + ".block2.norm.g", act=ACT_SILU, res=res, save_u=save, rps=rps)
        return out, (x1, x2, u1, h1, u2)

    def _resnet_bwd(self, pre, saved, dout, rps, need_dx=True, dx_acc=None):
        """`dx_acc` (identity-skip blocks only): a tensor of dx1's shape the input gradient is ACCUMULATED into (the
        skip-connection gradient that would otherwise be added by a separate pass); it is returned as dx1."""
        x1, x2, u1, h1, u2 = saved
        dh1, _ = self._conv_bwd_fused(dout, u2, pre + ".block2.norm.g", None, ACT_SILU, h1, None,
                                      pre + ".block2.proj.weight", pre + ".block2.proj.bias", 3, rps=rps)
        has_res = (pre + ".res_conv.weight") in self.specs
        if has_res and need_dx:
            # Block1 backward and the res_conv (1x1 over the same input) backward in one pass over x
            R, c1, L = x1.shape
            c2 = x2.shape[1] if x2 is not None else 0
            dx1 = self._empty(R, c1, L)
            dx2 = self._empty(R, c2, L) if c2 else None
            ss = self.ss_off[pre + ".mlp.1"]
            rc = N.call("dq_conv_bwd_fused_res", dh1, u1, self._w(pre + ".block1.norm.g"), _off_ptr(self._SS[:, ss:]),
                        self.ss_total, ACT_SILU, x1, c1, x2, c2, self._w(pre + ".block1.proj.weight"), dx1, dx2,
                        self._gw(pre + ".block1.proj.weight"), self._gw(pre + ".block1.proj.bias"),
                        self._gw(pre + ".block1.norm.g"), _off_ptr(self._dSS[:, ss:]), dout,
                        self._w(pre + ".res_conv.weight"), self._gw(pre + ".res_conv.weight"),
                        self._gw(pre + ".res_conv.bias"), dh1.shape[1], R, L, rps, allow=(1,))
            if rc == 0:
                return dx1, dx2
        dx1, dx2 = self._conv_bwd_fused(dh1, u1, pre + ".block1.norm.g", self.ss_off[pre + ".mlp.1"], ACT_SILU, x1, x2,
                                        pre + ".block1.proj.weight", pre + ".block1.proj.bias", 3, need_dx1=need_dx,
                                        need_dx2=need_dx, dx1=None if has_res else dx_acc,
                                        dadd=None if (has_res or not need_dx) else dout, rps=rps)
        if has_res:
            dx1, dx2 = self._conv_bwd_fused(dout, None, None, None, ACT_NONE, x1, x2, pre + ".res_conv.weight",
                                            pre + ".res_conv.bias", 1, need_dx1=need_dx, need_dx2=need_dx, dx1=dx1,
                                            dx2=dx2, rps=rps)
        return dx1, dx2

    def _la_fwd(self, pre, x, save):
        R, C, L = x.shape
        nch = N.la_nchunk(L)
        cp = _ceil8(C)
        part = self._empty(R, nch, HD, 2 + cp)
        msm = self._empty(R, HD, 2 + cp)   # per (row, head*32+d): max, sum, Ms[d][c] = sum_n softmax_L(k)[d,n] xn[c,n]
        gmat = self._empty(R, C, HD)       # G[c'][head*32+d] = sum_e Wout[c'][e] ctx[d][e]
        ypre = self._empty(R, C, L) if save else None
        out = self._empty(R, C, L)
        N.call("dq_linattn_fwd", x, self._w(pre + ".fn.norm.g"), self._w(pre + ".fn.fn.to_qkv.weight"),
               self._w(pre + ".fn.fn.to_out.0.weight"), self._w(pre + ".fn.fn.to_out.0.bias"),
               self._w(pre + ".fn.fn.to_out.1.g"), part, msm, gmat, ypre, out, C, R, L)
        return out, (x, ypre, msm, gmat)

    def _la_bwd(self, pre, saved, dres):
        x, ypre, msm, gmat = saved
        R, C, L = x.shape
        nch = N.la_nchunk(L)
        cp = _ceil8(C)
        dxnq = self._empty(R, C, L)
        dpart = self._empty(R, nch, HD, cp)
        hmat = self._empty(R, HD, cp)
        sd = self._empty(R, HD)
        dx = self._empty(R, C, L)
        N.call("dq_linattn_bwd", x, dres, ypre, msm, gmat, self._w(pre + ".fn.norm.g"),
               self._w(pre + ".fn.fn.to_qkv.weight"), self._w(pre + ".fn.fn.to_out.0.weight"),
               self._w(pre + ".fn.fn.to_out.1.g"), dxnq, dpart, hmat, sd, dx,
               self._gw(pre + ".fn.fn.to_qkv.weight"), self._gw(pre + ".fn.fn.to_out.0.weight"),
               self._gw(pre + ".fn.fn.to_out.0.bias"), self._gw(pre + ".fn.fn.to_out.1.g"),
               self._gw(pre + ".fn.norm.g"), C, R, L)
        return dx

    # ---- GEMM wrappers -------------------------------------------------------------------------------------
    def _gemm(self, A, a_rows, a_cols, a_ld, B, b_rows, b_cols, b_ld, b_tap_stride, b_ntaps, C, ldc, bias, acc, M, Nn,
              K, taps, a_row_off, a_k_off, b_k_off, b_tap, nz=1, z_b_koff_step=0, z_c_stride=0, z_b_tap_step=0):
        offs = list(a_row_off) + [0] * (4 - len(a_row_off)) + list(a_k_off) + [0] * (4 - len(a_k_off)) \
            + list(b_k_off) + [0] * (4 - len(b_k_off)) + list(b_tap) + [0] * (4 - len(b_tap))
        bn = self.gemm_bn   # 0: the library picks the tile width (multiple of 16) that leaves no ragged last wave
        N.call("dq_gemm_bf16_tn", A, a_rows, a_cols, a_ld, B, b_rows, b_cols, b_ld, b_tap_stride, b_ntaps, C, ldc,
               bias, acc, M, Nn, K, taps, offs, nz, z_b_koff_step, z_b_tap_step, z_c_stride, bn)

    gemm_bn = 0

    def _mid_conv_fwd(self, Ap, wname, bname, b, rt):
        """Ap: bf16 padded [Mp][N] -> fp32 padded [Mp][N] = conv3 over RT (+bias)."""
        Nm = self.mid_channels
        Mp = b * (rt + 2)
        U = self._empty(Mp, Nm)
        W = self._bf16[wname]
        self._gemm(Ap, Mp, Nm, Nm, W, Nm, Nm, Nm, Nm * Nm, 3, U, Nm, self._w(bname), 0, Mp, Nm, Nm, 3,
                   (-1, 0, 1), (0, 0, 0), (0, 0, 0), (0, 1, 2))
        return U

    def _mid_conv_dgrad(self, dUp, wname, b, rt):
        Nm = self.mid_channels
        Mp = b * (rt + 2)
        dX = self._empty(Mp, Nm)
        WT = self._bf16[wname + ".T"]
        self._gemm(dUp, Mp, Nm, Nm, WT, Nm, Nm, Nm, Nm * Nm, 3, dX, Nm, None, 0, Mp, Nm, Nm, 3,
                   (1, 0, -1), (0, 0, 0), (0, 0, 0), (0, 1, 2))
        return dX

    def _mid_conv_wgrad(self, dUp, Ap, wname, b, rt):
        """dW[t][co][ci] += sum_m' dU[m'][co] * A[m' + t - 1][ci]   (both padded bf16 [Mp][N]).

        With gradient accumulation over micro-batches (`_wgrad_defer = (rows_before, rows_total, is_last)`, set by
        ModelInterface._train_one_batch) the transposed operands of every micro-batch are appended along K and ONE
        GEMM per weight runs in the last micro-batch: the 1.2 GB fp32 read-modify-write of dW happens once per
        optimizer step instead of once per micro-batch."""
        Nm = self.mid_channels
        Mp = b * (rt + 2)
        plan = self._wgrad_defer
        if plan is None:
            off, total, last = 0, Mp, True
            ld = _ceil8(Mp)
            dUT = self._empty(Nm, ld, dtype=torch.bfloat16)
            AT3 = self._empty(3, Nm, ld, dtype=torch.bfloat16)  # AT3[t][ci][m'] = A[m' + t - 1][ci]
        else:
            off, total, last = plan
            ld = _ceil8(total)
            st = self._wgrad_stash.get(wname)
            if st is None or st[0].shape[1] != ld or st[0].device != dUp.device:
                st = (self._empty(Nm, ld, dtype=torch.bfloat16), self._empty(3, Nm, ld, dtype=torch.bfloat16))
                self._wgrad_stash[wname] = st
            dUT, AT3 = st
        N.call("dq_transpose_bf16", dUp, dUT.data_ptr() + 2 * off, Mp, Nm, ld, 0)
        for t in range(3):
            N.call("dq_transpose_bf16", Ap, AT3[t].data_ptr() + 2 * off, Mp, Nm, ld, t - 1)
        if last:
            self._gemm(dUT, Nm, total, ld, AT3, Nm, total, ld, Nm * ld, 3, self._gw(wname), Nm, None, 1, Nm, Nm, total, 1,
                       (0,), (0,), (0,), (0,), nz=3, z_b_tap_step=1, z_c_stride=Nm * Nm)

    def _mid_block_fwd(self, pre, X, b, rt, save):
        """X fp32 [M][N] -> fp32 [M][N]; ResnetBlock(N, N) with identity skip (unet1d.py:1029/1058)."""
        Nm = self.mid_channels
        M, Mp = b * rt, b * (rt + 2)
        Xp = self._empty(Mp, Nm, dtype=torch.bfloat16)
        N.call("dq_mid_pack", X, Xp, b, rt, Nm, 1)
        U1 = self._mid_conv_fwd(Xp, pre + ".block1.proj.weight", pre + ".block1.proj.bias", b, rt)
        H1p = self._zeros(Mp, Nm, dtype=torch.bfloat16)
        inv1 = self._empty(M)
        sso = self.ss_off[pre + ".mlp.1"]
        N.call("dq_rownorm_fwd", U1, 1, self._w(pre + ".block1.norm.g"), _off_ptr(self._SS[:, sso:]), self.ss_total,
               ACT_SILU, None, None, H1p, 1, inv1, b, rt, Nm)
        U2 = self._mid_conv_fwd(H1p, pre + ".block2.proj.weight", pre + ".block2.proj.bias", b, rt)
        out = self._empty(M, Nm)
        inv2 = self._empty(M)
        N.call("dq_rownorm_fwd", U2, 1, self._w(pre + ".block2.norm.g"), None, 0, ACT_SILU, X, out, None, 0, inv2,
               b, rt, Nm)
        return out, ((Xp, U1, inv1, H1p, U2, inv2) if save else None)

    def _mid_block_bwd(self, pre, saved, dOut, b, rt):
        Xp, U1, inv1, H1p, U2, inv2 = saved
        Nm = self.mid_channels
        M, Mp = b * rt, b * (rt + 2)
        dot = self._empty(M)
        dU2p = self._zeros(Mp, Nm, dtype=torch.bfloat16)
        N.call("dq_rownorm_bwd", dOut, 0, U2, 1, self._w(pre + ".block2.norm.g"), None, 0, ACT_SILU, inv2, dot, dU2p,
               1, None, 0, self._gw(pre + ".block2.norm.g"), None, self._gw(pre + ".block2.proj.bias"), b, rt, Nm)
        self._mid_conv_wgrad(dU2p, H1p, pre + ".block2.proj.weight", b, rt)
        dH1 = self._mid_conv_dgrad(dU2p, pre + ".block2.proj.weight", b, rt)
        dU1p = self._zeros(Mp, Nm, dtype=torch.bfloat16)
        sso = self.ss_off[pre + ".mlp.1"]
        N.call("dq_rownorm_bwd", dH1, 1, U1, 1, self._w(pre + ".block1.norm.g"), _off_ptr(self._SS[:, sso:]),
               self.ss_total, ACT_SILU, inv1, dot, dU1p, 1, None, 0, self._gw(pre + ".block1.norm.g"),
               _off_ptr(self._dSS[:, sso:]), self._gw(pre + ".block1.proj.bias"), b, rt, Nm)
        self._mid_conv_wgrad(dU1p, Xp, pre + ".block1.proj.weight", b, rt)
        dXp = self._mid_conv_dgrad(dU1p, pre + ".block1.proj.weight", b, rt)
        dX = self._empty(M, Nm)
        # un-pad and add the identity-skip gradient
        N.call("dq_rownorm_fwd", dXp, 1, None, None, 0, ACT_NONE, dOut, dX, None, 0, None, b, rt, Nm)
        return dX

    def _mid_attn_fwd(self, X, cond_nlc, b, rt, save):
        """Residual(PreNorm(Attention(use_xattn))) (unet1d.py:1030-1042, 541-567).  X fp32 [M][N]; cond (M, 8)."""
        Nm = self.mid_channels
        M = b * rt
        acd = self.dim * 2
        xn = self._empty(M, Nm, dtype=torch.bfloat16)
        inv = self._empty(M)
        N.call("dq_rownorm_fwd", X, 0, self._w("mid_attn.fn.norm.g"), None, 0, ACT_NONE, None, None, xn, 0, inv, b, rt, Nm)
        qv = self._empty(M, 2 * HD)
        self._gemm(xn, M, Nm, Nm, self._bf16["wqv"], 2 * HD, Nm, Nm, 0, 1, qv, 2 * HD, None, 0, M, 2 * HD, Nm, 1,
                   (0,), (0,), (0,), (0,))
        k = self._empty(M, HD)
        N.call("dq_linear_fwd", cond_nlc, self._w("mid_attn.fn.fn.to_k.weight"), None, k, M, acd, HD)
        P = self._empty(b, HEADS, rt, rt) if save else None
        O = self._empty(M, HD, dtype=torch.bfloat16)
        N.call("dq_attn_core_fwd", qv, k, self._w("mid_attn.fn.fn.rotary_emb.freqs"), P, None, O, b, rt)
        out = X.clone()
        self._gemm(O, M, HD, HD, self._bf16["wout"], Nm, HD, HD, 0, 1, out, Nm, self._w("mid_attn.fn.fn.to_out.bias"),
                   1, M, Nm, HD, 1, (0,), (0,), (0,), (0,))
        return out, ((X, inv, xn, qv, k, P, O, cond_nlc) if save else None)

    def _mid_attn_bwd(self, saved, dOut, b, rt):
        X, inv, xn, qv, k, P, O, cond_nlc = saved
        Nm = self.mid_channels
        M = b * rt
        acd = self.dim * 2
        ld = _ceil8(M)
        dOb = self._empty(M, Nm, dtype=torch.bfloat16)
        N.call("dq_mid_pack", dOut, dOb, b, rt, Nm, 0)
        N.call("dq_colsum", dOut, self._gw("mid_attn.fn.fn.to_out.bias"), M, Nm)
        # d to_out weight [N][128] += dOut^T . O
        dOT = self._empty(Nm, ld, dtype=torch.bfloat16)
        OT = self._empty(HD, ld, dtype=torch.bfloat16)
        N.call("dq_transpose_bf16", dOb, dOT, M, Nm, ld, 0)
        N.call("dq_transpose_bf16", O, OT, M, HD, ld, 0)
        self._gemm(dOT, Nm, M, ld, OT, HD, M, ld, 0, 1, self._gw("mid_attn.fn.fn.to_out.weight"), HD, None, 1, Nm, HD,
                   M, 1, (0,), (0,), (0,), (0,))
        # d O = dOut . Wout
        dO = self._empty(M, HD)
        self._gemm(dOb, M, Nm, Nm, self._bf16["wout.T"], HD, Nm, Nm, 0, 1, dO, HD, None, 0, M, HD, Nm, 1,
                   (0,), (0,), (0,), (0,))
        dqv = self._empty(M, 2 * HD)
        dqvb = self._empty(M, 2 * HD, dtype=torch.bfloat16)
        dk = self._empty(M, HD)
        N.call("dq_attn_core_bwd", qv, k, self._w("mid_attn.fn.fn.rotary_emb.freqs"), P, dO, dqv, dqvb, dk, b, rt)
        # to_k backward (tiny)
        dcond = self._empty(M, acd)
        N.call("dq_linear_bwd", cond_nlc, self._w("mid_attn.fn.fn.to_k.weight"), dk, dcond,
               self._gw("mid_attn.fn.fn.to_k.weight"), None, M, acd, HD)
        # d Wqv [256][N] += dqv^T . xn
        dqvT = self._empty(2 * HD, ld, dtype=torch.bfloat16)
        xnT = self._empty(Nm, ld, dtype=torch.bfloat16)
        N.call("dq_transpose_bf16", dqvb, dqvT, M, 2 * HD, ld, 0)
        N.call("dq_transpose_bf16", xn, xnT, M, Nm, ld, 0)
        self._gemm(dqvT, 2 * HD, M, ld, xnT, Nm, M, ld, 0, 1, self._gw("mid_attn.fn.fn.to_qv.weight"), Nm, None, 1,
                   2 * HD, Nm, M, 1, (0,), (0,), (0,), (0,))
        # d xn = dqv . Wqv
        dxn = self._empty(M, Nm)
        self._gemm(dqvb, M, 2 * HD, 2 * HD, self._bf16["wqv.T"], Nm, 2 * HD, 2 * HD, 0, 1, dxn, Nm, None, 0, M, Nm,
                   2 * HD, 1, (0,), (0,), (0,), (0,))
        dX = dOut.clone()
        dot = self._empty(M)
        N.call("dq_rownorm_bwd", dxn, 0, X, 0, self._w("mid_attn.fn.norm.g"), None, 0, ACT_NONE, inv, dot, None, 0, dX,
               1, self._gw("mid_attn.fn.norm.g"), None, None, b, rt, Nm)
        return dX, dcond

    # ---------------------------------------------------------------------------------------------- forward
    def _time_path_fwd(self, time, b, save):
        td, dim = self.time_dim, self.dim
        half = dim // 2
        neg_e = float(np.float32(-(math.log(self.theta) / (half - 1))))
        e0 = self._empty(b, dim)
        N.call("dq_time_embed", time, e0, b, dim, neg_e)
        h = self._empty(b, td)
        N.call("dq_linear_fwd", e0, self._w("time_mlp.1.weight"), self._w("time_mlp.1.bias"), h, b, dim, td)
        a = self._empty(b, td)
        N.call("dq_act_fwd", h, a, ACT_GELU, h.numel())
        t = self._empty(b, td)
        N.call("dq_linear_fwd", a, self._w("time_mlp.3.weight"), self._w("time_mlp.3.bias"), t, b, td, td)
        st = self._empty(b, td)
        N.call("dq_act_fwd", t, st, ACT_SILU, t.numel())
        SS = self._empty(b, self.ss_total)
        N.call("dq_linear_fwd", st, self._flat[self.ss_w_off:], self._flat[self.ss_b_off:], SS, b, td, self.ss_total)
        self._SS = SS
        return (e0, h, a, t, st) if save else None

    def _time_path_bwd(self, saved, b):
        e0, h, a, t, st = saved
        td, dim = self.time_dim, self.dim
        dst = self._empty(b, td)
        N.call("dq_linear_bwd", st, self._flat[self.ss_w_off:], self._dSS, dst, self._gflat[self.ss_w_off:],
               self._gflat[self.ss_b_off:], b, td, self.ss_total)
        dt = self._empty(b, td)
        N.call("dq_act_bwd", dst, t, dt, ACT_SILU, dt.numel())
        da = self._empty(b, td)
        N.call("dq_linear_bwd", a, self._w("time_mlp.3.weight"), dt, da, self._gw("time_mlp.3.weight"),
               self._gw("time_mlp.3.bias"), b, td, td)
        dh = self._empty(b, td)
        N.call("dq_act_bwd", da, h, dh, ACT_GELU, dh.numel())
        N.call("dq_linear_bwd", e0, self._w("time_mlp.1.weight"), dh, None, self._gw("time_mlp.1.weight"),
               self._gw("time_mlp.1.bias"), b, dim, td)

    def _forward_impl(self, x, time, init_cond, attn_cond, save):
        if not x.is_cuda:
            raise N.NativeError("UNet1d (B200) has no CPU path: move the model and inputs to a CUDA device")
        squeeze = x.dim() == 2
        if squeeze:
            x = x[None]
        b, rt, L = x.shape
        if L != self.downsample_dim:
            raise ValueError(f"m/z length {L} must equal downsample_dim {self.downsample_dim} (unet1d.py:1029, 1144)")
        R = b * rt
        n_lv = len(self.in_out)
        x = x.contiguous().float().view(R, 1, L)
        ic = torch.zeros_like(x) if init_cond is None else init_cond.contiguous().float().view(R, 1, L)
        if attn_cond is None:
            ac0 = self._zeros(b, 1, rt)
        else:
            if attn_cond.dim() != 2 and not (attn_cond.dim() == 3 and attn_cond.shape[-1] == 1):
                raise NotImplementedError("attn_cond must be the (b, rt) MS1 chromatogram")
            ac0 = attn_cond.contiguous().float().view(b, 1, rt)
        time = time.to(torch.long).contiguous().view(b)
        T = _Tape() if save else None
        tp = self._time_path_fwd(time, b, save)

        # init conv on cat(cond*(s+1)+sh, x)                                   unet1d.py:1107-1118
        ico = self.ss_off["init_cond_proj.to_scale_shift.1"]
        x0, _ = self._conv_fwd(ic, x, "init_conv.weight", "init_conv.bias", 7, 1, 3, 1, L, in_ss=ico, rps=rt)
        # MS1 chromatogram -> (b, 8, rt)                                        unet1d.py:1120-1130
        a1, a1u = self._conv_fwd(ac0, None, "attn_cond_proj.1.0.weight", "attn_cond_proj.1.0.bias", 7, 1, 3, 1, rt,
                                 act=ACT_GELU, save_u=save, rps=1)
        a2, _ = self._conv_fwd(a1, None, "attn_cond_proj.1.2.weight", "attn_cond_proj.1.2.bias", 1, 1, 0, 1, rt, rps=1)
        acd = self.dim * 2
        cond_nlc = self._empty(b, rt, acd)
        N.call("dq_ncl_nlc", a2, cond_nlc, b, acd, rt, 0)

        h = []
        saved_down = []
        cur = x0
        for i in range(n_lv):
            pre = f"downs.{i}"
            a, s0 = self._resnet_fwd(pre + ".0", cur, None, rt, save)
            h.append(a)
            bb, s1 = self._resnet_fwd(pre + ".1", a, None, rt, save)
            c, s2 = self._la_fwd(pre + ".2", bb, save)
            h.append(c)
            Lc = c.shape[2]
            if i < n_lv - 1:
                Ln = (Lc + 2 - 4) // 2 + 1
                nxt, _ = self._conv_fwd(c, None, pre + ".3.weight", pre + ".3.bias", 4, 2, 1, 1, Ln, rps=rt)
            else:
                nxt, _ = self._conv_fwd(c, None, pre + ".3.weight", pre + ".3.bias", 3, 1, 1, 1, Lc, rps=rt)
            saved_down.append((s0, s1, s2, c))
            cur = nxt

        d, mzd = cur.shape[1], cur.shape[2]
        if d * mzd != self.mid_channels:
            raise ValueError("input length is inconsistent with downsample_dim")
        Xm = cur.view(R, d * mzd)  # [(b rt)][(d mz)]: the reference's rearrange is a view here
        self._refresh_bf16()       # here, not at entry: a deferred parameter all-gather overlaps the down path above
        m1, sm1 = self._mid_block_fwd("mid_block1", Xm, b, rt, save)
        m2, sma = self._mid_attn_fwd(m1, cond_nlc.view(R, acd), b, rt, save)
        m3, sm2 = self._mid_block_fwd("mid_block2", m2, b, rt, save)
        cur = m3.view(R, d, mzd)

        saved_up = []
        for j in range(n_lv):
            pre = f"ups.{j}"
            skip_c = h.pop()
            y, s0 = self._resnet_fwd(pre + ".0", cur, skip_c, rt, save)
            skip_a = h.pop()
            z, s1 = self._resnet_fwd(pre + ".1", y, skip_a, rt, save)
            w, s2 = self._la_fwd(pre + ".2", z, save)
            Lw = w.shape[2]
            if j < n_lv - 1:
                nxt, _ = self._conv_fwd(w, None, pre + ".3.1.weight", pre + ".3.1.bias", 3, 1, 1, 2, Lw * 2, rps=rt)
            else:
                nxt, _ = self._conv_fwd(w, None, pre + ".3.weight", pre + ".3.bias", 3, 1, 1, 1, Lw, rps=rt)
            saved_up.append((s0, s1, s2, w))
            cur = nxt

        f, sf = self._resnet_fwd("final_res_block", cur, x0, rt, save)
        out, _ = self._conv_fwd(f, None, "final_conv.weight", "final_conv.bias", 1, 1, 0, 1, L, rps=rt)
        out_pre = None
        if self.pos_output_only:
            out_pre = out
            out = torch.empty_like(out_pre)
            N.call("dq_act_fwd", out_pre, out, ACT_SOFTPLUS, out.numel())
        out = out.view(b, rt, L)
        if squeeze:
            pass  # the reference also returns (1, rt, mz) for 2-D input ("(b rt) d mz -> b (rt d) mz" with b=1)
        if save:
            T.d = dict(b=b, rt=rt, L=L, tp=tp, SS=self._SS, ic=ic, x=x, x0=x0, ac0=ac0, a1=a1, a1u=a1u, cond_nlc=cond_nlc,
                       down=saved_down, up=saved_up, sm1=sm1, sma=sma, sm2=sm2, sf=sf, f=f, mid_shape=(d, mzd),
                       out_pre=out_pre)
        return out, T

    # ---------------------------------------------------------------------------------------------- backward
    def _backward_impl(self, T, d_out):
        S = T.d
        b, rt, L = S["b"], S["rt"], S["L"]
        R = b * rt
        n_lv = len(self.in_out)
        self._ensure_grads()
        self._SS = S["SS"]
        self._dSS = self._zeros(b, self.ss_total)
        acd = self.dim * 2
        d_out = d_out.contiguous().float().view(R, 1, L)
        if S.get("out_pre") is not None:   # Softplus head
            dz = torch.empty_like(d_out)
            N.call("dq_act_bwd", d_out, S["out_pre"], dz, ACT_SOFTPLUS, dz.numel())
            d_out = dz

        # final conv + final res block
        df, _ = self._conv_bwd(d_out, S["f"], None, "final_conv.weight", "final_conv.bias", 1, 1, 0, 1, rps=rt)
        dcur, dx0 = self._resnet_bwd("final_res_block", S["sf"], df, rt)

        skip_grads = []  # in pop order: c6, a6, c5, a5, ...
        for j in reversed(range(n_lv)):
            pre = f"ups.{j}"
            s0, s1, s2, w = S["up"][j]
            if j < n_lv - 1:
                dw, _ = self._conv_bwd(dcur, w, None, pre + ".3.1.weight", pre + ".3.1.bias", 3, 1, 1, 2, rps=rt)
            else:
                dw, _ = self._conv_bwd(dcur, w, None, pre + ".3.weight", pre + ".3.bias", 3, 1, 1, 1, rps=rt)
            dz = self._la_bwd(pre + ".2", s2, dw)
            dy, dskip_a = self._resnet_bwd(pre + ".1", s1, dz, rt)
            dcur, dskip_c = self._resnet_bwd(pre + ".0", s0, dy, rt)
            skip_grads.append((dskip_c, dskip_a))
        # skip_grads[k] belongs to up level j = n_lv-1-k, which consumed down level i = n_lv-1-j = k
        d, mzd = S["mid_shape"]
        dm = dcur.view(R, d * mzd)
        dm = self._mid_block_bwd("mid_block2", S["sm2"], dm, b, rt)
        dm, dcond = self._mid_attn_bwd(S["sma"], dm, b, rt)
        dm = self._mid_block_bwd("mid_block1", S["sm1"], dm, b, rt)
        dcur = dm.view(R, d, mzd)
        if self.grad_ready_callback is not None and self._final_microbatch:
            self.grad_ready_callback(self.early_grad_ranges())

        # up level j popped the skips of down level (n_lv-1-j); skip_grads was appended for j = n_lv-1 .. 0
        sg = {}
        for idx, j in enumerate(reversed(range(n_lv))):
            sg[n_lv - 1 - j] = skip_grads[idx]
        for i in reversed(range(n_lv)):
            pre = f"downs.{i}"
            s0, s1, s2, c = S["down"][i]
            dskip_c, dskip_a = sg[i]
            if i < n_lv - 1:
                dc, _ = self._conv_bwd(dcur, c, None, pre + ".3.weight", pre + ".3.bias", 4, 2, 1, 1, dx1=dskip_c, rps=rt)
            else:
                dc, _ = self._conv_bwd(dcur, c, None, pre + ".3.weight", pre + ".3.bias", 3, 1, 1, 1, dx1=dskip_c, rps=rt)
            db = self._la_bwd(pre + ".2", s2, dc)
            has_res = (pre + ".1.res_conv.weight") in self.specs
            da, _ = self._resnet_bwd(pre + ".1", s1, db, rt, dx_acc=None if has_res else dskip_a)   # += skip gradient
            if has_res:
                N.call("dq_add_inplace", da, dskip_a, da.numel())
            last = (i == 0) and (pre + ".0.res_conv.weight") not in self.specs
            dcur, _ = self._resnet_bwd(pre + ".0", s0, da, rt, dx_acc=dx0 if last else None)

        # init conv: gradient of its output = down-path gradient + final-res-block skip gradient
        if "downs.0.0.res_conv.weight" in self.specs:   # otherwise dx0 was accumulated by the block backward above
            N.call("dq_add_inplace", dcur, dx0, dcur.numel())
        ico = self.ss_off["init_cond_proj.to_scale_shift.1"]
        # one pass over (d, cond, x): dW, db and the per-sample d scale / d shift of the ConditionalScaleShift from raw
        # per-sample correlations (csrc/small.cu); the data gradient of the conditioning channel never exists
        rc = 1
        if S["ic"].shape[1] == 1 and S["x"].shape[1] == 1 and not getattr(self, "_force_generic_initconv", False):
            scratch = torch.zeros(b, max(1, self.dim // 4), 88, device=dcur.device, dtype=torch.float32)
            rc = N.call("dq_initconv_bwd", dcur, S["ic"], S["x"], _off_ptr(self._SS[:, ico:]), self.ss_total,
                        self._w("init_conv.weight"), self._gw("init_conv.weight"), self._gw("init_conv.bias"),
                        _off_ptr(self._dSS[:, ico:]), scratch, self.dim, R, L, rt, allow=(1,))
        if rc != 0:
            dxc = self._empty(R, 1, L)
            # weight/bias gradient with the ConditionalScaleShift applied to source 1; data gradient only for channel 0
            self._conv_bwd(dcur, S["ic"], S["x"], "init_conv.weight", "init_conv.bias", 7, 1, 3, 1, need_dx1=False,
                           need_dx2=False, in_ss=ico, rps=rt)
            N.call("dq_conv1d_bwd_data", dcur, self._w("init_conv.weight"), dxc, 1, 0, None, 1, 0, self.dim, 7, 1, 3, 1,
                   R, L, L)
            N.call("dq_sample_dot", dxc, S["ic"], _off_ptr(self._dSS[:, ico:]), _off_ptr(self._dSS[:, ico + 1:]),
                   self.ss_total, rt * L, b)

        # MS1 conditioning path
        da2 = self._empty(b, acd, rt)
        N.call("dq_ncl_nlc", dcond.view(b, rt, acd), da2, b, acd, rt, 1)
        da1, _ = self._conv_bwd(da2, S["a1"], None, "attn_cond_proj.1.2.weight", "attn_cond_proj.1.2.bias", 1, 1, 0, 1, rps=1)
        da1u = self._block_bwd(da1, S["a1u"], None, None, ACT_GELU, 1)
        self._conv_bwd(da1u, S["ac0"], None, "attn_cond_proj.1.0.weight", "attn_cond_proj.1.0.bias", 7, 1, 3, 1,
                       need_dx1=False, rps=1)

        self._time_path_bwd(S["tp"], b)
        self._dSS = None

    # ---------------------------------------------------------------------------------------------- public forward
    def forward(self, x, time, init_cond=None, attn_cond=None):
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._params.values())
        if need_grad:
            anchor = self._params["final_conv.bias"]
            return _UNetFn.apply(self, x, time, init_cond, attn_cond, anchor)
        out, _ = self._forward_impl(x, time, init_cond, attn_cond, save=False)
        return out


def _off_ptr(t):
    """Raw device address of a (possibly non-contiguous, column-offset) view; None passes through."""
    return None if t is None else t.data_ptr()


class _UNetFn(torch.autograd.Function):
    """One autograd node for the whole denoiser.  Parameter gradients are accumulated by the kernels straight
    into the flat gradient buffer (each `p.grad` is a view of it), so backward returns None for every input;
    the gradient w.r.t. the noisy input is not needed by the training step and is not computed."""

    @staticmethod
    def forward(ctx, net, x, time, init_cond, attn_cond, anchor):
        out, tape = net._forward_impl(x, time, init_cond, attn_cond, save=True)
        ctx.net = net
        ctx.tape = tape
        net.last_tape = tape
        return out

    @staticmethod
    def backward(ctx, d_out):
        ctx.net._backward_impl(ctx.tape, d_out)
        ctx.tape = None
        ctx.net.last_tape = None
        return None, None, None, None, None, None
