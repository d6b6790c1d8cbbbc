"""Tensor-core dense GEMMs of the bf16 path (csrc/dense_tc.cu: tcgen05 kind::f16 over a tiled bf16
copy of the kernel) vs torch: tf.layers.dense forward / data gradient of DINCAE/__init__.py:479-483.
Operands are bf16-exact test data and accumulation is fp32, so only summation order differs:
tolerance 1e-4 of the output's maximum (stated; the bf16 rounding of REAL weights is covered by the
network-level bf16 tolerances)."""
import numpy as np
import pytest
import torch

from dincae_b200 import _lib
from dincae_b200._lib import check, ptr

pytestmark = pytest.mark.gpu


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,K,N", [(2, 128, 64), (16, 2648, 536), (50, 536, 2648), (7, 1000, 264), (130, 256, 1032)])
def test_dense_tc_forward_and_data_gradient(B, K, N):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + K)
    ldw = N + 8                                    # padded master rows
    w = torch.zeros(K, ldw)
    w[:, :N] = bf(torch.randn(K, N, generator=g) * 0.05)
    x = bf(torch.randn(B, K, generator=g))
    dz = bf(torch.randn(B, N, generator=g))
    bias = torch.randn(N, generator=g)
    act = torch.randn(B, K, generator=g)
    wd, xd, dzd, biasd, actd = w.cuda(), x.cuda(), dz.cuda(), bias.cuda(), act.cuda()   # keep the device copies alive
    w8 = torch.zeros(lib.dincae_dense_tc_weights_elems(K, N), dtype=torch.bfloat16, device="cuda")
    check(lib.dincae_dense_tc_tile_weights(ptr(wd), ldw, K, N, ptr(w8), stream()))
    # the tiled copy holds every element at [k/8][n/8][k%8][n%8]
    t = w8.float().cpu().view(K // 8, N // 8, 8, 8).permute(0, 2, 1, 3).reshape(K, N)
    assert torch.equal(t, w[:, :N])
    ws_bytes = lib.dincae_dense_tc_workspace(B, K, N)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    y = torch.full((B, N + 4), 9.0, device="cuda")
    check(lib.dincae_dense_tc_fwd(ptr(xd), K, ptr(w8), ptr(biasd), B, K, N, 1, ptr(y), N + 4, ptr(ws),
                                  ws_bytes, stream()))
    dx = torch.full((B, K), 9.0, device="cuda")
    check(lib.dincae_dense_tc_bwd_data(ptr(dzd), N, ptr(w8), ptr(actd), B, K, N, ptr(dx), K, ptr(ws),
                                       ws_bytes, stream()))
    torch.cuda.synchronize()
    want_y = torch.relu(x.double() @ w[:, :N].double() + bias.double())
    want_dx = (dz.double() @ w[:, :N].double().t()) * (act > 0)
    assert float((y[:, :N].cpu().double() - want_y).abs().max()) < 1e-4 * float(want_y.abs().max())
    assert float(y[:, N:].min()) == 9.0                       # the row padding is not touched
    assert float((dx.cpu().double() - want_dx).abs().max()) < 1e-4 * float(want_dx.abs().max())
