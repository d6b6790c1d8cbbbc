"""Data-parallel scheme on CPU (gloo, world_size 2): every rank leaves its gradient
un-normalised, gradients and the four loss sums are all-reduced, and the division by the
global count of observed pixels happens afterwards -- this must equal one batch of 2*B on a
single worker (mean of means != global mean).  The arithmetic here is
the torch oracle; the test pins the host-side protocol the CUDA trainer follows, including
the rank-slicing of the shared index stream."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dincae_b200 import sampler
from oracle import model_torch as M


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    arch = M.Arch(16, 12, 10, [8, 12], [0.2], [1, 2])
    params = M.init_params(arch, 0)
    g = torch.Generator().manual_seed(1)
    B = 3
    xin = torch.randn(world * B, 16, 12, 10, generator=g)
    xtrue = torch.zeros(world * B, 16, 12, 2)
    xtrue[..., 1] = (torch.rand(world * B, 16, 12, generator=g) > 0.7 - 0.3 * (torch.arange(world * B) % 2)[:, None, None]).float()
    xtrue[..., 0] = torch.randn(world * B, 16, 12, generator=g) * xtrue[..., 1]
    # same global stream on every rank, rank r takes its slice
    s = sampler.TrainSampler(world * B, world * B, shuffle_buffer_size=4, seed=7)
    f, c, q = s.next_batch()
    mine = f[rank * B:(rank + 1) * B]
    ps = [p.clone().requires_grad_(True) for p in params]
    xrec = M.forward(arch, ps, xin[mine])
    cost, rms, n = M.cost_terms(xrec, xtrue[mine])
    grads = torch.autograd.grad(cost * n, ps)          # un-normalised: sum, not mean
    flat = torch.cat([gr.reshape(-1) for gr in grads])
    sums = torch.tensor([float(cost * n), float(n)], dtype=torch.float64)
    dist.all_reduce(flat)
    dist.all_reduce(sums)
    flat = flat / sums[1].float()
    if rank == 0:
        ps = [p.clone().requires_grad_(True) for p in params]
        xrec = M.forward(arch, ps, xin[f])
        cost_all, _, n_all = M.cost_terms(xrec, xtrue[f])
        want = torch.cat([gr.reshape(-1) for gr in torch.autograd.grad(cost_all, ps)])
        ret["grad_err"] = float((flat - want).abs().max() / want.abs().max())
        ret["cost_err"] = abs(float(sums[0] / sums[1]) - float(cost_all))
        ret["n_ok"] = float(sums[1]) == float(n_all)
    dist.destroy_process_group()


def test_two_ranks_equal_one_double_batch():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["n_ok"]
    assert ret["grad_err"] < 1e-5
    assert ret["cost_err"] < 1e-5
