"""CPU, world_size 2, gloo: the data-parallel gradient exchange of the harness (bucketed all-reduce of the flat
gradient, early all-reduce of the mid-stage ranges launched from backward, averaging) without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FlatModel(torch.nn.Module):
    """Stand-in exposing the flat-buffer interface of the B200 UNet1d."""

    def __init__(self, n):
        super().__init__()
        self._flat = torch.zeros(n)
        self._g = torch.zeros(n)
        self.n_trainable_flat = n - 8
        self.grad_ready_callback = None

    def flat_params(self):
        return self._flat

    def flat_grads(self):
        return self._g


def _worker(rank, world, port, early):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dquartic.model.model_interface import ModelInterface

    n = 10_000
    mi = ModelInterface(device="cpu")
    mi.model = _FlatModel(n)
    mi.grad_bucket_elems = 1024  # force several buckets
    g = mi.model.flat_grads()
    torch.manual_seed(100 + rank)
    g.copy_(torch.randn(n))
    mine = g.clone()
    if early:
        mi._early_allreduce([(2000, 3000), (7000, 500)])
    mi._allreduce_grads()
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ref = sum(gathered) / world
    ok_main = torch.allclose(g[: n - 8], ref[: n - 8], rtol=1e-6, atol=1e-7)
    ok_tail = torch.equal(g[n - 8:], mine[n - 8:])  # non-trainable tail (rotary freqs slot) is not exchanged
    assert ok_main and ok_tail, (rank, ok_main, ok_tail)
    assert mi._early_reduced == [] and mi._early_works == []
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("early", [False, True])
def test_bucketed_allreduce_world2_gloo(early):
    mp.spawn(_worker, args=(2, _free_port(), early), nprocs=2, join=True)
