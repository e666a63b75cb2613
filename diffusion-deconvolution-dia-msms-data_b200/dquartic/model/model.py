"""`DDIMDiffusionModel` — drop-in for /root/reference/dquartic/model/model.py:151-406 on B200.

Same constructor, attributes (`betas, alphas, alpha_bars, loss_weight, normalize, unnormalize, pred_type,
ms1_loss_weight, model, num_timesteps, device`) and methods (`q_sample, p_sample, sample, train_step`).
The schedule tables are built exactly like the reference (fp64 -> fp32, fp32 cumprod) so they are bit-identical;
q_sample, the reverse step, the final un-normalise and the epsilon-MSE are single fused CUDA kernels.

Batched semantics (the reference only runs at b = 1): `train_step` returns a (b,) tensor whose
entries are the batch-mean MSE times `loss_weight[t_i]`; call `.mean().backward()` (the harness does).
"""
import math

import numpy as np
import torch

from .. import _native as N
from .model_interface import ModelInterface


def get_linear_beta_schedule(num_timesteps, beta_start=0.0001, beta_end=0.02):
    return torch.linspace(beta_start, beta_end, num_timesteps, dtype=torch.float64)


def get_cosine_beta_schedule(num_timesteps, s=0.008):
    t = torch.linspace(0, num_timesteps, num_timesteps + 1, dtype=torch.float64)
    f = torch.cos(((t / num_timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    f = f / f[0]
    return torch.clip(1 - (f[1:] / f[:-1]), 0, 0.999)


def get_alphas(betas):
    return 1.0 - betas


def get_alpha_bars(alpha):
    return torch.cumprod(alpha, dim=0)


def normalize_to_neg_one_to_one(img):
    if img.is_cuda and img.dtype == torch.float32:
        out = torch.empty_like(img, memory_format=torch.contiguous_format)
        N.call("dq_mix_affine", img.contiguous(), None, 1.0, 0.0, 2.0, -1.0, out, img.numel())
        return out
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):
    if t.is_cuda and t.dtype == torch.float32:
        out = torch.empty_like(t, memory_format=torch.contiguous_format)
        N.call("dq_add_mul", t.contiguous(), 1.0, 0.5, out, t.numel())
        return out
    return (t + 1) * 0.5


def identity(t, *args, **kwargs):
    return t


def extract(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


class _MSEFn(torch.autograd.Function):
    """mean((pred - target)^2) with the gradient produced in the same pass (F.mse_loss, model.py:361)."""

    @staticmethod
    def forward(ctx, pred, target):
        pred = pred.contiguous()
        target = target.contiguous()
        n = pred.numel()
        acc = torch.zeros(1, dtype=torch.float64, device=pred.device)
        d = torch.empty_like(pred)
        N.call("dq_mse", pred, target, acc, d, 2.0 / n, n)
        ctx.save_for_backward(d)
        return (acc / n).to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None


class DDIMDiffusionModel(ModelInterface):
    def __init__(
        self,
        model_class,
        num_timesteps=1000,
        beta_schedule_type="cosine",
        pred_type="eps",
        auto_normalize=True,
        ms1_loss_weight=0.0,
        device="cuda",
        **kwargs,
    ):
        super().__init__()
        self.model = None
        self.build(model_class, **kwargs)
        self.num_timesteps = num_timesteps
        self.device = device

        betas64 = get_linear_beta_schedule(num_timesteps) if beta_schedule_type == "linear" else get_cosine_beta_schedule(num_timesteps)
        self.betas = betas64.to(device).to(torch.float32)
        self.alphas = get_alphas(self.betas).to(torch.float32)
        # cumprod on the CPU in fp32, like the reference's CPU run, then moved: keeps the table bit-identical
        self.alpha_bars = get_alpha_bars(self.alphas.cpu()).to(torch.float32).to(device)
        self._ab_host = self.alpha_bars.detach().cpu().numpy().astype(np.float32)

        snr = self.alpha_bars / (1 - self.alpha_bars)
        if pred_type == "eps":
            self.loss_weight = torch.ones_like(snr)
        elif pred_type == "x0":
            self.loss_weight = snr
        else:
            raise ValueError(f"Unknown pred_type: {pred_type}")

        self.auto_normalize = auto_normalize
        self.normalize = normalize_to_neg_one_to_one if auto_normalize else identity
        self.unnormalize = unnormalize_to_zero_to_one if auto_normalize else identity
        self.pred_type = pred_type
        self.ms1_loss_weight = ms1_loss_weight

    # ------------------------------------------------------------------------------------------ forward process
    def q_sample(self, x_0, t, noise=None):
        """sqrt(ab[t]) * x_0 + sqrt(1 - ab[t]) * noise (model.py:225-242); x_0 is taken as given (already normalised)."""
        x_0 = x_0.float()   # the kernels compute in fp32 (a float64 tensor must not be reinterpreted)
        if noise is None:
            noise = torch.randn_like(x_0)
        noise = noise.float()
        b = x_0.shape[0]
        out = torch.empty_like(x_0, memory_format=torch.contiguous_format)
        N.call("dq_qsample", x_0.contiguous(), noise.contiguous(), t.to(torch.long).contiguous(), self.alpha_bars, out,
               b, x_0.numel() // b, 0)
        return out

    def _q_sample_fused(self, x_0_raw, t, noise):
        b = x_0_raw.shape[0]
        out = torch.empty_like(x_0_raw, memory_format=torch.contiguous_format)
        N.call("dq_qsample", x_0_raw.contiguous(), noise.contiguous(), t.contiguous(), self.alpha_bars, out, b,
               x_0_raw.numel() // b, 1 if self.auto_normalize else 0)
        return out

    # ------------------------------------------------------------------------------------------ reverse process
    def _step_coefs(self, t):
        ab = self._ab_host
        one = np.float32(1.0)
        sa = np.sqrt(ab[t])
        s1m = np.sqrt(one - ab[t])
        if t > 0:
            sap = np.sqrt(ab[t - 1])  # index t-1 regardless of the sampling stride (model.py:284)
            s1mp = np.sqrt(one - ab[t - 1])
        else:
            sap, s1mp = one, np.float32(0.0)
        return float(sa), float(s1m), float(sap), float(s1mp)

    def p_sample(self, x_t, t, init_cond=None, attn_cond=None):
        if self.pred_type not in ("eps", "x0"):
            raise ValueError(f"Unknown pred_type: {self.pred_type}")
        batch_size = x_t.size(0)
        t = int(t)
        x_t = x_t.float()
        t_tensor = torch.full((batch_size,), t, device=x_t.device, dtype=torch.long)
        sa, s1m, sap, s1mp = self._step_coefs(t)
        out = self.model(x_t, t_tensor, init_cond, attn_cond)
        if self.pred_type == "eps":
            eps_pred = out.contiguous()
            x_prev = torch.empty_like(eps_pred)
            N.call("dq_ddim_step", x_t.contiguous(), eps_pred, x_prev, sa, s1m, sap, s1mp, 1 if t == 0 else 0,
                   eps_pred.numel())
            return x_prev, eps_pred
        # pred_type == "x0": a secondary mode; composed from torch elementwise ops
        x0_pred = out
        eps_pred = (x_t - sa * x0_pred) / s1m
        x_prev = sap * x0_pred + s1mp * eps_pred if t > 0 else x0_pred
        return x_prev, eps_pred

    def sample(self, x_t, ms2_cond=None, ms1_cond=None, num_steps=1000):
        ms2_cond = self.normalize(ms2_cond) if ms2_cond is not None else None
        ms1_cond = self.normalize(ms1_cond) if ms1_cond is not None else None
        pred_noise = None
        time_steps = torch.linspace(self.num_timesteps - 1, 0, num_steps, dtype=torch.long).tolist()
        for t in time_steps:  # the step list lives on the host: no per-step device sync (reference: t.item())
            x_t, pred_noise = self.p_sample(x_t, t, ms2_cond, ms1_cond)
        if ms2_cond is not None and self.auto_normalize and x_t.is_cuda:
            xo = torch.empty_like(x_t)
            pn = torch.empty_like(x_t)
            N.call("dq_sample_finalize", x_t.contiguous(), ms2_cond.contiguous(), xo, pn, x_t.numel())
            return xo, pn
        x_t, pred_noise = self.unnormalize(x_t), self.unnormalize(pred_noise)
        if ms2_cond is not None:
            pred_noise = self.unnormalize(ms2_cond) - x_t
        return x_t, pred_noise

    # ------------------------------------------------------------------------------------------ training step
    def train_step(self, x_0, ms2_cond=None, ms1_cond=None, noise=None, ms1_loss_weight=0.0, t=None):
        """model.py:326-406.  `t` (optional, not in the reference) injects the timesteps for parity tests."""
        if ms1_loss_weight and ms1_loss_weight > 0.0:
            raise NotImplementedError(
                "ms1_loss_weight > 0 raises TypeError in the reference (torch.max(dim=-1) returns a tuple, "
                "model.py:366-368); the SIC loss is not defined (SURVEY.md §8f-3)")
        if self.pred_type not in ("eps", "x0"):
            raise ValueError(f"Unknown pred_type: {self.pred_type}")
        batch_size = x_0.size(0)
        dev = x_0.device
        if t is None:
            t = torch.randint(0, self.num_timesteps, (batch_size,), device=dev).long()
        else:
            t = t.to(dev).long()
        x_0 = x_0.float()
        noise = torch.randn_like(x_0) if noise is None else self.normalize(noise.float())
        ms2_n = self.normalize(ms2_cond) if ms2_cond is not None else None
        ms1_n = self.normalize(ms1_cond) if ms1_cond is not None else None
        x_t = self._q_sample_fused(x_0.float(), t, noise)
        pred = self.model(x_t, t, ms2_n, ms1_n)
        if self.pred_type == "eps":
            primary = _MSEFn.apply(pred, noise)
        else:
            primary = _MSEFn.apply(pred, self.normalize(x_0))
        return primary * extract(self.loss_weight, t, (batch_size,))
