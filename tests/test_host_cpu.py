"""CPU (no GPU needed): the C-ABI library loads and exports every symbol of include/dquartic_b200.h, and the host
side mirrors the reference interface (names, shapes, schedules, data path, config rules, error behaviour)."""
import json
import os
import random
import re

import numpy as np
import pytest
import torch

from _util import ROOT, TINY, golden

import dquartic_oracle as O
from dquartic import _native
from dquartic.model.model import DDIMDiffusionModel
from dquartic.model.model_interface import FusedAdamW, WarmupLR_Scheduler
from dquartic.model.unet1d import UNet1d
from dquartic.utils.config_loader import generate_train_config, load_train_config
from dquartic.utils.data_loader import DIAMSDataset


def _net(cfg=TINY):
    return UNet1d(dim=cfg["dim"], channels=1, dim_mults=tuple(cfg["dim_mults"]), conditional=True,
                  init_cond_channels=1, attn_cond_channels=1, downsample_dim=cfg["downsample_dim"], simple=True)


def test_cabi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dquartic_b200.h")).read()
    declared = set(re.findall(r"^int (dq_\w+)\(", hdr, flags=re.M))
    assert len(declared) >= 35
    lib = _native.lib()  # raises if lib/libdquartic_b200.so is missing (no CPU fallback)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_native.exported_symbols())
    for name, (sig, _) in _native._SIGS.items():  # the ctypes arity equals the C declaration's parameter count
        m = re.search(r"int " + name + r"\((.*?)\);", hdr, flags=re.S)
        assert m, name
        assert len(m.group(1).split(",")) == len(sig), name


def test_state_dict_names_shapes_order_and_counts_match_reference():
    info = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    net = _net()
    sd = net.state_dict()
    assert [k for k, _ in info["tiny"]["keys"]] == list(sd.keys())
    assert all(list(sd[k].shape) == s for k, s in info["tiny"]["keys"])
    assert sum(p.numel() for p in net.parameters()) == info["tiny"]["total"]
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == info["tiny"]["trainable"]
    # default and notebook configurations: inventory only (no 4.8 GB allocation)
    from dquartic.model.unet1d import param_specs
    for name in ("default", "notebook"):
        cfg = info[name]["cfg"]
        specs = param_specs(cfg["dim"], tuple(cfg["dim_mults"]), 1, 1, 1, cfg["downsample_dim"])
        assert [k for k, _ in info[name]["keys"]] == list(specs.keys())
        assert sum(int(np.prod(s)) for s in specs.values()) == info[name]["total"]


def test_flat_buffer_views_and_state_dict_roundtrip():
    net = _net()
    P = O.det_params(TINY)
    net.load_state_dict(P)
    for k, v in net.state_dict().items():
        assert torch.equal(v, P[k]), k
    # every parameter aliases the flat buffer; the mid conv is stored tap-major
    w = net._params["mid_block1.block1.proj.weight"]
    assert w.untyped_storage().data_ptr() == net.flat_params().untyped_storage().data_ptr()
    assert w.stride()[2] == w.shape[0] * w.shape[1]
    assert torch.equal(net._w("mid_block1.block1.proj.weight").view(3, 80, 80)[1], P["mid_block1.block1.proj.weight"][:, :, 1])
    # scale/shift producers are contiguous so one Linear yields every (scale, shift) of the network
    n_ss = sum(net.specs[p + ".weight"][0] for p in net._ss_producers())
    assert net.ss_total == n_ss and net.ss_b_off == n_ss * net.time_dim
    # .double()/.half() are refused, .to(float32) keeps aliasing
    with pytest.raises(NotImplementedError):
        net.half()
    net.to(torch.float32)
    assert net._params["init_conv.weight"].untyped_storage().data_ptr() == net.flat_params().untyped_storage().data_ptr()


def test_no_cpu_fallback():
    net = _net()
    with pytest.raises(_native.NativeError):
        net(torch.zeros(1, 4, 320), torch.zeros(1, dtype=torch.long), torch.zeros(1, 4, 320), torch.zeros(1, 4))
    with pytest.raises(NotImplementedError):
        UNet1d(dim=4, channels=1, dim_mults=(1, 2), init_cond_channels=1, attn_cond_channels=1, simple=False)
    with pytest.raises(ValueError):
        DDIMDiffusionModel(net, pred_type="v", device="cpu")


def test_schedule_tables_bit_exact_against_reference():
    g = golden("schedule.npz")
    net = _net()
    for kind in ("cosine", "linear"):
        d = DDIMDiffusionModel(net, beta_schedule_type=kind, device="cpu")
        assert np.array_equal(d.betas.numpy(), g[f"{kind}_betas"])
        assert np.array_equal(d.alphas.numpy(), g[f"{kind}_alphas"])
        assert np.array_equal(d.alpha_bars.numpy(), g[f"{kind}_alpha_bars"])
    d = DDIMDiffusionModel(net, pred_type="x0", device="cpu")
    assert np.array_equal(d.loss_weight.numpy(), g["cosine_x0_loss_weight"])
    # reverse-step coefficients use alpha_bars[t-1] regardless of stride (model.py:284)
    sa, s1m, sap, s1mp = d._step_coefs(978)
    ab = g["cosine_alpha_bars"]
    assert sa == float(np.sqrt(ab[978])) and sap == float(np.sqrt(ab[977]))
    assert d._step_coefs(0)[2:] == (1.0, 0.0)


def test_dataset_pairs_and_minmax_bit_exact_against_reference(tmp_path):
    g = golden("data.npz")
    np.save(tmp_path / "ms2.npy", g["ms2_pool"])
    np.save(tmp_path / "ms1.npy", g["ms1_pool"])
    ds = DIAMSDataset(ms2_file=str(tmp_path / "ms2.npy"), ms1_file=str(tmp_path / "ms1.npy"), normalize="minmax")
    random.seed(1234)
    for j in range(6):
        item = ds[0]
        for nm, arr in zip(("ms2_1", "ms1_1", "ms2_2", "ms1_2"), item):
            assert arr.dtype == torch.float32
            assert np.array_equal(arr.numpy(), g[f"item{j}:{nm}"]), (j, nm)
    # pair sequence for N = 520 over two epochs, python `random` seeded like the golden run
    class Fake(DIAMSDataset):
        def __init__(self):
            self.ms2_data = np.zeros((520, 1, 1), np.int32)
            self.ms1_data = np.zeros((520, 1), np.int32)
            self.normalize = "minmax"
            self.used_pairs = set()
    f = Fake()
    random.seed(1234)
    seq = []
    for epoch in range(2):
        f.reset_epoch()
        seq += [f.draw_pair() for _ in range(64)]
    assert np.array_equal(np.array(seq), g["pairs520"])
    with pytest.raises(ValueError):
        DIAMSDataset()
    ds.normalize = None
    with pytest.raises(ValueError):
        ds[0]


def test_lr_schedule_and_optimizer_state_format():
    net = _net()
    opt = FusedAdamW(net, lr=1e-3)
    sched = WarmupLR_Scheduler(opt, num_warmup_steps=5, num_training_steps=20)
    lrs = []
    for e in range(20):
        lrs.append(sched.get_last_lr()[0])
        sched.step(e, 0.0)
    for e in range(20):
        assert lrs[e] == pytest.approx(1e-3 * O.lr_lambda(e, 5, 20), rel=1e-9)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(396))
    assert sd["param_groups"][0]["betas"] == (0.9, 0.999) and sd["param_groups"][0]["weight_decay"] == 1e-2


def test_config_loader_override_rules(tmp_path):
    p = str(tmp_path / "cfg.json")
    generate_train_config(p)
    ref = json.load(open("/root/reference/dquartic_train_config.json")) if os.path.exists("/root/reference") else None
    cfg = load_train_config(p)
    if ref is not None:
        assert cfg == ref
    cfg = load_train_config(p, batch_size=8, checkpoint_path=None, use_wandb=False, threads=2, ms2_data_path="a.npy")
    assert cfg["model"]["batch_size"] == 8 and cfg["model"]["checkpoint_path"] == "best_model.ckpt"
    assert cfg["wandb"]["use_wandb"] is False and cfg["threads"] == 2 and cfg["data"]["ms2_data_path"] == "a.npy"


def test_cli_surface():
    from click.testing import CliRunner
    from dquartic.cli import cli

    r = CliRunner().invoke(cli, ["train", "--help"])
    assert r.exit_code == 0
    for opt in ("--parquet_directory", "--ms2-data-path", "--ms1-data-path", "--batch-size", "--checkpoint-path",
                "--use-wandb", "--threads"):
        assert opt in r.output
    assert CliRunner().invoke(cli, ["--version"]).exit_code == 0
