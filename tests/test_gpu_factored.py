"""Factored dense gradients (csrc/dense_factored.cu): the optimiser works from the factors x, dz
of dW = x^T dz instead of the materialised matrix (DINCAE/__init__.py:479-483 MatMul gradients,
:598-603 clip_by_global_norm + Adam).  Checked against the materialised path of this library and
against torch; the two paths differ only in summation order, so tolerances are fp32 rounding level."""
import numpy as np
import pytest
import torch

from dincae_b200 import _lib, layout
from dincae_b200._lib import check, ptr
from dincae_b200.engine import Plan, make_network
from oracle import model_torch as M

pytestmark = pytest.mark.gpu


def stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("world,rows,K,N", [(1, 5, 40, 24), (2, 8, 132, 68), (3, 50, 260, 52), (8, 50, 96, 32)])
def test_factor_ops_match_the_materialised_gradient(world, rows, K, N):
    lib = _lib.load()
    g = torch.Generator().manual_seed(world * 100 + rows)
    pad = 64                                            # per-rank blobs with unrelated data in between
    stride = rows * K + rows * N + 2 * pad
    # bf16-representable factors (what the bf16 network produces): the tensor-core variant of the
    # Adam kernel (more than 16 frames) multiplies TF32 operands, which hold bf16 values exactly
    blob = torch.randn(world, stride, generator=g).to(torch.bfloat16).float().cuda()
    xs = [blob[r, pad: pad + rows * K].view(rows, K) for r in range(world)]
    zs = [blob[r, 2 * pad + rows * K: 2 * pad + rows * K + rows * N].view(rows, N) for r in range(world)]
    x, dz = torch.cat(xs).double(), torch.cat(zs).double()
    dw = x.t() @ dz
    xb = blob.data_ptr() + 4 * pad
    zb = blob.data_ptr() + 4 * (2 * pad + rows * K)
    ws_bytes = lib.dincae_dense_factored_workspace(world * rows, K, N)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    ss = torch.full((1,), 3.0, dtype=torch.float64, device="cuda")
    check(lib.dincae_dense_factored_sumsq(xb, stride, K, zb, stride, N, world, rows, K, N, ptr(ss), ptr(ws),
                                          ws_bytes, stream()))
    got = torch.zeros(K, N, device="cuda")
    check(lib.dincae_dense_factored_expand(xb, stride, K, zb, stride, N, world, rows, K, N, ptr(got), N, stream()))
    torch.cuda.synchronize()
    want_ss = float((dw ** 2).sum())
    assert abs(ss.item() - 3.0 - want_ss) < 1e-5 * want_ss
    assert float((got.double().cpu() - dw.cpu()).abs().max()) < 1e-5 * float(dw.abs().max())

    # Adam from the factors == Adam on the materialised gradient (same scalars in `state`)
    w0 = torch.randn(K, N, generator=g).cuda()
    m0 = torch.randn(K, N, generator=g).cuda() * 0.01
    v0 = torch.rand(K, N, generator=g).cuda() * 0.01
    state = torch.tensor([0.9, 0.999, 0.0, 2e-3, 0.37, 0, 0, 0], dtype=torch.float32, device="cuda")
    denom = torch.tensor([7.0], dtype=torch.float64, device="cuda")
    w, m, v = w0.clone(), m0.clone(), v0.clone()
    check(lib.dincae_adam_update_factored(ptr(w), ptr(m), ptr(v), N, xb, stride, K, zb, stride, N, world, rows, K, N,
                                          1.5, ptr(denom), 0.9, 0.999, 1e-8, ptr(state), None, stream()))
    torch.cuda.synchronize()
    gk = dw.float().cuda() * (1.5 / 7.0) * 0.37
    m_ref = m0 + (gk - m0) * 0.1
    v_ref = v0 + (gk * gk - v0) * 0.001
    w_ref = w0 - (m_ref * 2e-3) / (v_ref.sqrt() + 1e-8)
    assert float((m - m_ref).abs().max()) < 5e-5 * float(m_ref.abs().max())     # fp32 sums of up to 400 terms
    assert float((v - v_ref).abs().max()) < 5e-5 * float(v_ref.abs().max())
    assert float((w - w_ref).abs().max()) < 2e-6

    # bias gradient of one rank's rows
    db = torch.zeros(N, device="cuda")
    check(lib.dincae_dense_bias_grad(zb, rows, N, N, ptr(db), stream()))
    torch.cuda.synchronize()
    assert torch.allclose(db.cpu(), zs[0].sum(0).cpu(), rtol=1e-5, atol=1e-5)


def test_adam_step_multi_sums_rank_copies():
    lib = _lib.load()
    g = torch.Generator().manual_seed(4)
    n, world, stride = 1003, 3, 1100
    grads = torch.randn(world, stride, generator=g).cuda()
    p0 = torch.randn(n, generator=g).cuda()
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.tensor([0.9, 0.999, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device="cuda")
    lr = torch.tensor([1e-3], device="cuda")
    extra = torch.tensor([11.0], dtype=torch.float64, device="cuda")
    denom = torch.tensor([2.0], dtype=torch.float64, device="cuda")
    ws_bytes = lib.dincae_adam_multi_workspace(n)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib.dincae_adam_step_multi(ptr(p), ptr(grads), stride, world, ptr(m), ptr(v), n, 5.0, ptr(lr), 0.9, 0.999,
                                     1e-8, 1.0, ptr(denom), ptr(extra), ptr(state), ptr(ws), ws_bytes, stream()))
    torch.cuda.synchronize()
    gsum = grads[:, :n].double().sum(0)
    norm = float(torch.sqrt((gsum ** 2).sum() + 11.0)) / 2.0
    assert abs(state[2].item() - norm) < 1e-5 * norm
    scale = 5.0 * min(1.0 / norm, 1.0 / 5.0)
    gk = (gsum / 2.0 * scale).float().cuda()
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    m_ref, v_ref = gk * 0.1, gk * gk * 0.001
    p_ref = p0 - (m_ref * lr_t) / (v_ref.sqrt() + 1e-8)
    assert float((p - p_ref).abs().max()) < 2e-6
    assert abs(state[0].item() - 0.81) < 1e-6 and abs(state[1].item() - 0.998001) < 1e-6


CASES = {
    "tiny": dict(B=2, H=16, W=16, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "odd": dict(B=3, H=30, W=22, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "dense3": dict(B=2, H=16, W=12, nvar=6, filt=[8, 12], frac=[0.3, 0.1], skip=[1]),
}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("lanes", [4, 8])
def test_network_factored_equals_materialised(case, lanes):
    c = CASES[case]
    arch = M.Arch(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    plan = Plan(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"], lanes=lanes)
    g = torch.Generator().manual_seed(7)
    params = M.init_params(arch, 3)
    B, H, W, nvar = c["B"], c["H"], c["W"], c["nvar"]
    xin = torch.randn(B, H, W, nvar, generator=g).to(torch.bfloat16).float()
    obs = torch.rand(B, H, W, generator=g) > 0.5
    xtrue = torch.zeros(B, H, W, 2)
    xtrue[..., 1] = obs.float()
    xtrue[..., 0] = torch.randn(B, H, W, generator=g) * obs
    nets = [make_network(plan, B + 1, train=True, factored=f) for f in (False, True)]
    assert nets[1].factored and not nets[0].factored
    for net in nets:
        net.set_params_tf([p.numpy() for p in params])
        net.lr.fill_(1e-3)
    for step in range(3):
        x = xin + 0.05 * step
        flats = []
        for net in nets:
            net.X[0][:B].copy_(torch.from_numpy(layout.nhwc_to_p4(x.numpy())))
            net.xtrue[:B].copy_(xtrue)
            net.forward(B)
            net.loss_backward(B, False, normalise=False)
            flats.append(net.flat_grads(B).cpu().double())
            net.adam_step(clip_norm=5.0, denom=net.loss_sums[3:4] if net.factored else net.loss_sums[3:4], B=B)
        torch.cuda.synchronize()
        err = float((flats[0] - flats[1]).abs().max()) / float(flats[0].abs().max())
        assert err < 2e-5, (step, err)
        assert abs(nets[0].adam_state[2].item() - nets[1].adam_state[2].item()) < 1e-5 * nets[0].adam_state[2].item()
        assert torch.equal(nets[0].loss_sums, nets[1].loss_sums)
    pa, pb = nets[0].params.cpu(), nets[1].params.cpu()
    # Adam normalises steps, so an entry whose gradient is at rounding level may differ by a step
    assert float((pa - pb).abs().max()) < 2.1e-3 and float((pa - pb).abs().mean()) < 1e-6


@pytest.mark.parametrize("rows,K,N", [(5, 136, 72), (16, 2648, 536), (50, 536, 2648)])
def test_factored_adam_keeps_the_tiled_bf16_copy_current(rows, K, N):
    """adam_update_factored(w8=...) leaves w8 == tile(bf16(updated w)), the operand of dense_tc.cu."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g).cuda()
    dz = torch.randn(rows, N, generator=g).cuda()
    w = torch.randn(K, N, generator=g).cuda()
    m = torch.zeros(K, N, device="cuda")
    v = torch.zeros(K, N, device="cuda")
    state = torch.tensor([0.9, 0.999, 0.0, 1e-2, 1.0, 0, 0, 0], dtype=torch.float32, device="cuda")
    w8 = torch.zeros(K * N, dtype=torch.bfloat16, device="cuda")
    check(lib.dincae_adam_update_factored(ptr(w), ptr(m), ptr(v), N, ptr(x), rows * K, K, ptr(dz), rows * N, N, 1, rows,
                                          K, N, 1.0, None, 0.9, 0.999, 1e-8, ptr(state), ptr(w8), stream()))
    ref8 = torch.zeros_like(w8)
    check(lib.dincae_dense_tc_tile_weights(ptr(w), N, K, N, ptr(ref8), stream()))
    torch.cuda.synchronize()
    assert torch.equal(w8, ref8)


@pytest.mark.parametrize("rows,K,N,with_w8", [(5, 40, 24, False), (16, 136, 72, True), (7, 2648, 536, True),
                                              (16, 536, 2648, True), (12, 8296, 1032, True), (3, 20, 300, False)])
def test_tma_pipelined_adam_equals_the_register_staged_kernel(rows, K, N, with_w8, monkeypatch):
    """adam_factored_tma_kernel (persistent CTAs, TMA loads / stores, 3 stages) against
    adam_factored_kernel on the same inputs: same summation order -> bit-identical m, v; weights
    within the approximate-division bound; bf16 tile copy consistent with the weights; ragged row / column tiles, tile ranges that cross column blocks,
    fewer tiles than CTAs; two consecutive steps so that the second starts from non-trivial m, v."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(rows * 7 + K)
    x = torch.randn(rows, K, generator=g).cuda()
    dz = torch.randn(rows, N, generator=g).cuda()
    w0 = torch.randn(K, N, generator=g).cuda()
    state = torch.tensor([0.9, 0.999, 0.0, 1e-2, 0.8, 0, 0, 0], dtype=torch.float32, device="cuda")
    denom = torch.tensor([3.0], dtype=torch.float64, device="cuda")
    out = []
    for mode in ("0", "1"):
        monkeypatch.setenv("DINCAE_ADAM_TMA", mode)
        w, m, v = w0.clone(), torch.zeros(K, N, device="cuda"), torch.zeros(K, N, device="cuda")
        w8 = torch.zeros(K * N, dtype=torch.bfloat16, device="cuda") if with_w8 else None
        for _ in range(2):
            check(lib.dincae_adam_update_factored(ptr(w), ptr(m), ptr(v), N, ptr(x), rows * K, K, ptr(dz), rows * N, N,
                                                  1, rows, K, N, 1.0, ptr(denom), 0.9, 0.999, 1e-8, ptr(state),
                                                  ptr(w8) if with_w8 else None, stream()))
        torch.cuda.synchronize()
        out.append((w, m, v, w8))
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])      # m, v
    # the weight step of the TMA kernel uses MUFU sqrt / reciprocal: < 4e-7 of a step of <= 2 lr
    # (i.e. a last-bit rounding difference of the weight at most)
    assert bool(((out[0][0] - out[1][0]).abs() <= 2.4e-7 * out[0][0].abs().clamp(min=1.0)).all())
    assert not torch.equal(out[0][0], w0)
    if with_w8:                                            # each copy is the bf16 tile form of its own weights
        for w, _, _, w8 in out:
            ref8 = torch.zeros_like(w8)
            check(lib.dincae_dense_tc_tile_weights(ptr(w), N, K, N, ptr(ref8), stream()))
            torch.cuda.synchronize()
            assert torch.equal(w8, ref8)


@pytest.mark.parametrize("world,rows,K,N", [(2, 16, 536, 2648), (4, 16, 2648, 536), (8, 16, 1040, 776), (3, 7, 136, 72),
                                            (1, 100, 520, 1032), (2, 9, 8296, 264)])
def test_tma_pipelined_tensor_core_adam_equals_the_staged_kernels(world, rows, K, N, monkeypatch):
    """adam_factored_tma_mma_kernel (16 < frames <= 128: TMA-swizzled p/m/v tiles, B fragments in
    registers, column-block walk) against adam_factored_mma_kernel / _persist_kernel: the same
    mma.sync sequence -> bit-identical m, v; weights within the approximate-division bound; bf16
    tile copy consistent; ragged tiles, several tasks per CTA and fewer tasks than CTAs."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(world * 31 + rows)
    stride_x, stride_z = rows * K + 32, rows * N + 64
    x = torch.randn(world, stride_x, generator=g).to(torch.bfloat16).float().cuda()
    dz = torch.randn(world, stride_z, generator=g).to(torch.bfloat16).float().cuda()
    w0 = torch.randn(K, N, generator=g).cuda()
    state = torch.tensor([0.9, 0.999, 0.0, 1e-2, 0.8, 0, 0, 0], dtype=torch.float32, device="cuda")
    denom = torch.tensor([3.0], dtype=torch.float64, device="cuda")
    out = []
    for mode in ("0", "1"):
        monkeypatch.setenv("DINCAE_ADAM_TMA", mode)
        w, m, v = w0.clone(), torch.zeros(K, N, device="cuda"), torch.zeros(K, N, device="cuda")
        w8 = torch.zeros(K * N, dtype=torch.bfloat16, device="cuda")
        for _ in range(2):
            check(lib.dincae_adam_update_factored(ptr(w), ptr(m), ptr(v), N, ptr(x), stride_x, K, ptr(dz), stride_z, N,
                                                  world, rows, K, N, 1.0, ptr(denom), 0.9, 0.999, 1e-8, ptr(state),
                                                  ptr(w8), stream()))
        torch.cuda.synchronize()
        out.append((w, m, v, w8))
    assert torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    assert bool(((out[0][0] - out[1][0]).abs() <= 2.4e-7 * out[0][0].abs().clamp(min=1.0)).all())
    assert not torch.equal(out[0][0], w0)
    for w, _, _, w8 in out:
        ref8 = torch.zeros_like(w8)
        check(lib.dincae_dense_tc_tile_weights(ptr(w), N, K, N, ptr(ref8), stream()))
        torch.cuda.synchronize()
        assert torch.equal(w8, ref8)
    # and against the exact product
    dw = torch.cat([x[r, :rows * K].view(rows, K) for r in range(world)]).double().t() @ \
        torch.cat([dz[r, :rows * N].view(rows, N) for r in range(world)]).double()
    gk = dw.float() * (0.8 / 3.0)
    assert float((out[1][1] - (0.1 * gk + 0.9 * 0.1 * gk)).abs().max()) < 1e-4 * float(gk.abs().max())
