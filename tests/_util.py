import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "diffusion-deconvolution-dia-msms-data_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

TINY = dict(dim=4, channels=1, dim_mults=[1, 2, 2, 3, 3, 4, 4], conditional=True, init_cond_channels=1,
            attn_cond_channels=1, tfer_dim_mult=620, downsample_dim=320, simple=True)


def make_net(cfg=TINY, seed=0, device="cuda"):
    """B200 UNet1d with the oracle's deterministic parameters loaded; returns (net, P_cpu)."""
    import dquartic_oracle as O
    from dquartic.model.unet1d import UNet1d

    net = UNet1d(dim=cfg["dim"], channels=1, dim_mults=tuple(cfg["dim_mults"]), conditional=True,
                 init_cond_channels=1, attn_cond_channels=1, downsample_dim=cfg["downsample_dim"], simple=True)
    P = O.det_params(cfg, seed)
    net.load_state_dict(P)
    return net.to(device), P


def rel_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))
