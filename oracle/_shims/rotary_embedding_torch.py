"""Test-infrastructure shim: restatement of the third-party `rotary-embedding-torch` package
(pinned ^0.8.4 in the reference's pyproject.toml:23; NOT vendored under /root/reference and not
installed in this image).  Only the two entry points the reference calls are provided:

    RotaryEmbedding(dim=dim_head // 2)                 reference unet1d.py:529
    .rotate_queries_or_keys(t)  on (b, h, n, d)        reference unet1d.py:560-561

Published semantics restated here: freqs = 1/theta^(arange(0,dim,2)/dim) stored as a non-trainable
nn.Parameter called `freqs`; positions 0..n-1 along dim -2; angles = pos (x) freqs, each repeated
twice (interleaved); the first `dim` features are rotated with interleaved-pair rotate_half; the
rest pass through.  No xpos, no offset, no interpolation.  PARITY UNPINNED at this boundary: the
reference has no test for it; the only known-answer is the 8-element non-trainable parameter that
shows up in the torchinfo table of nbs/quantization_experiment.ipynb cell 14.

This file exists ONLY so that oracle/gen_golden.py can import the unmodified reference in the
authoring container.  Nothing in the product imports it.
"""
import torch
from torch import nn


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(dim=-1)
    x = torch.stack((-x2, x1), dim=-1)
    return x.reshape(*x.shape[:-2], -1)


def apply_rotary_emb(freqs, t, start_index=0, scale=1.0, seq_dim=-2):
    if t.ndim == 3:
        seq_len = t.shape[seq_dim]
        freqs = freqs[-seq_len:]
    rot_dim = freqs.shape[-1]
    end_index = start_index + rot_dim
    t_left, t_mid, t_right = t[..., :start_index], t[..., start_index:end_index], t[..., end_index:]
    t_mid = (t_mid * freqs.cos() * scale) + (rotate_half(t_mid) * freqs.sin() * scale)
    return torch.cat((t_left, t_mid, t_right), dim=-1).type(t.dtype)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000, learned_freq=False):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=learned_freq)

    def forward(self, t):
        freqs = torch.einsum("..., f -> ... f", t.type(self.freqs.dtype), self.freqs)
        return freqs.repeat_interleave(2, dim=-1)

    def rotate_queries_or_keys(self, t, seq_dim=-2, offset=0):
        seq_len = t.shape[seq_dim]
        seq = torch.arange(seq_len, device=t.device, dtype=t.dtype) + offset
        freqs = self.forward(seq)
        return apply_rotary_emb(freqs, t, seq_dim=seq_dim)
