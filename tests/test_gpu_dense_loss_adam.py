"""Dense bottleneck, output head + masked NLL, and clip + TF-Adam kernels vs the torch-CPU
restatement (DINCAE/__init__.py:465-485, 520-570, 598-603)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dincae_b200 import _lib
from dincae_b200._lib import check, ptr
from oracle import model_torch as M
from tests.util import dev, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fp32_backend():
    """These kernels are fp32; with the tensor-core conv back end selected the loss head rounds
    the gradient it hands to the last convolution to TF32 (checked in
    test_loss_head_rounds_gradient_for_tensor_core_backend)."""
    from dincae_b200 import _lib
    lib = _lib.load()
    lib.dincae_set_conv_backend(0)
    yield
    lib.dincae_set_conv_backend(1)


def stream():
    return torch.cuda.current_stream().cuda_stream


# batches <= 16 with K, N % 4 == 0 take the weight-streaming ("skinny") kernels: multi-split
# partials, single-split fused epilogues, 8- and 16-row register tiles, ragged edges
@pytest.mark.parametrize("B,K,N", [(50, 2744, 532), (3, 70, 33), (64, 128, 64), (130, 257, 19),
                                   (8, 2744, 532), (13, 1000, 4100), (4, 64, 520), (16, 300, 1028),
                                   (1, 128, 128), (9, 4, 4)])
def test_dense_forward_backward(B, K, N):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + K + N)
    x = torch.randn(B, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    dy = torch.randn(B, N, generator=g)
    gate = torch.randn(B, K, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = F.relu(xr @ wr + br)
    (y * dy).sum().backward()
    dz = (dy * (y > 0)).detach()

    ws_bytes = lib.dincae_dense_workspace(B, K, N)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    xd, wd, bd, dzd, gd = dev(x.numpy()), dev(w.numpy()), dev(b.numpy()), dev(dz.numpy()), dev(gate.numpy())
    yd = torch.zeros(B, N, device="cuda")
    check(lib.dincae_dense_fwd(ptr(xd), ptr(wd), ptr(bd), B, K, N, 1, ptr(yd), ptr(ws), ws_bytes, stream()))
    dxd = torch.zeros(B, K, device="cuda")
    check(lib.dincae_dense_bwd_data(ptr(dzd), ptr(wd), None, B, K, N, ptr(dxd), ptr(ws), ws_bytes, stream()))
    dxg = torch.zeros(B, K, device="cuda")
    check(lib.dincae_dense_bwd_data(ptr(dzd), ptr(wd), ptr(gd), B, K, N, ptr(dxg), ptr(ws), ws_bytes, stream()))
    dwd = torch.zeros(K, N, device="cuda")
    dbd = torch.zeros(N, device="cuda")
    check(lib.dincae_dense_bwd_weight(ptr(xd), ptr(dzd), B, K, N, ptr(dwd), ptr(dbd), stream()))
    torch.cuda.synchronize()
    assert rel_err(yd.cpu().numpy(), y.detach().numpy()) < 2e-5
    assert rel_err(dxd.cpu().numpy(), xr.grad.numpy()) < 2e-5
    assert rel_err(dxg.cpu().numpy(), (xr.grad * (gate > 0)).numpy()) < 2e-5
    assert rel_err(dwd.cpu().numpy(), wr.grad.numpy()) < 2e-5
    assert rel_err(dbd.cpu().numpy(), br.grad.numpy()) < 2e-5


@pytest.mark.parametrize("B,H,W,uncertain", [(2, 9, 13, False), (3, 16, 16, True), (50, 112, 112, False)])
def test_head_and_loss(B, H, W, uncertain):
    lib = _lib.load()
    g = torch.Generator().manual_seed(H + W)
    pre = torch.randn(B, H, W, 2, generator=g) * 3
    pre[0, 0, 0, 1] = 30.0        # exercises min(.,10)
    pre[0, 0, 1, 1] = -30.0       # exercises max(exp,1e-3) after leaky relu (-6)
    xtrue = torch.randn(B, H, W, 2, generator=g)
    obs = torch.rand(B, H, W, generator=g) > 0.45
    xtrue[..., 1] = obs.float()
    xtrue[..., 0] = xtrue[..., 0] * obs
    pr = pre.clone().requires_grad_(True)
    xrec = F.leaky_relu(pr, 0.2)
    cost, rms, n = M.cost_terms(xrec, xtrue, uncertain)
    cost.backward()
    m_rec, s2 = M.head(xrec.detach())

    xd = torch.zeros(B, 1, H, W, 4, device="cuda")
    xd[:, 0, :, :, :2] = xrec.detach().cuda()
    xt = dev(xtrue.numpy())
    nn = torch.tensor([int(n.item())], dtype=torch.int32, device="cuda")
    gout = torch.zeros(B, 1, H, W, 4, device="cuda")
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    lout = torch.zeros(2, device="cuda")
    ws_bytes = lib.dincae_head_loss_workspace(B, H, W)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    for rep in range(2):          # second call checks the workspace is left reusable
        check(lib.dincae_head_loss(ptr(xd), ptr(xt), B, H, W, int(uncertain), 0.2, ptr(nn), ptr(gout),
                                   ptr(sums), ptr(lout), ptr(ws), ws_bytes, stream()))
    mr = torch.zeros(B, H, W, device="cuda")
    sr = torch.zeros(B, H, W, device="cuda")
    check(lib.dincae_head_recon(ptr(xd), B, H, W, ptr(mr), ptr(sr), stream()))
    torch.cuda.synchronize()
    assert rel_err(mr.cpu().numpy(), m_rec.numpy()) < 1e-5
    assert rel_err(sr.cpu().numpy(), s2.numpy()) < 1e-5
    assert abs(lout[0].item() - cost.item()) < 2e-5 * max(1.0, abs(cost.item()))
    assert abs(lout[1].item() - rms.item()) < 2e-5 * max(1.0, abs(rms.item()))
    assert sums[3].item() == n.item()
    assert rel_err(gout[:, 0, :, :, :2].cpu().numpy(), pr.grad.numpy()) < 2e-5
    assert float(gout[..., 2:].abs().max()) == 0.0
    # tensor-core conv back end: the same gradient, rounded to TF32 on store
    from tests.util import tf32_round
    lib.dincae_set_conv_backend(1)
    g1 = torch.zeros_like(gout)
    check(lib.dincae_head_loss(ptr(xd), ptr(xt), B, H, W, int(uncertain), 0.2, ptr(nn), ptr(g1),
                               ptr(sums), ptr(lout), ptr(ws), ws_bytes, stream()))
    torch.cuda.synchronize()
    lib.dincae_set_conv_backend(0)
    assert np.array_equal(g1.cpu().numpy(), tf32_round(gout.cpu().numpy()))
    # un-normalised mode: gradient * 1/n equals the normalised one
    g2 = torch.zeros_like(gout)
    check(lib.dincae_head_loss(ptr(xd), ptr(xt), B, H, W, int(uncertain), 0.2, None, ptr(g2),
                               ptr(sums), ptr(lout), ptr(ws), ws_bytes, stream()))
    torch.cuda.synchronize()
    assert rel_err((g2 / n.item()).cpu().numpy(), gout.cpu().numpy()) < 1e-6
    assert abs(lout[0].item() - cost.item()) < 2e-5 * max(1.0, abs(cost.item()))


@pytest.mark.parametrize("n,n_decay,beta,clip", [(1000, 600, 0.0, 5.0), (4099, 4000, 0.01, 5.0),
                                                 (300000, 250000, 0.001, 0.5)])
def test_clip_and_tf_adam(n, n_decay, beta, clip):
    lib = _lib.load()
    g = torch.Generator().manual_seed(n)
    p = torch.randn(n, generator=g)
    params = [p[:n_decay].clone(), p[n_decay:].clone()]
    opt = M.TFAdam(params)
    pd = dev(p.numpy())
    md, vd = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.zeros(8, device="cuda")
    state[0], state[1] = 0.9, 0.999
    lr = torch.tensor([1e-3], device="cuda")
    ws_bytes = lib.dincae_adam_workspace(n)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    denom = torch.tensor([4.0], dtype=torch.float64, device="cuda")
    for step in range(3):
        grads = torch.randn(n, generator=g) * (3.0 if step == 1 else 0.01)
        gd = dev((grads * 4.0).numpy())
        gl = [grads[:n_decay] + beta * params[0], grads[n_decay:].clone()]
        clipped, norm = M.clip_by_global_norm(gl, clip)
        opt.step(params, clipped, 1e-3)
        check(lib.dincae_adam_step(ptr(pd), ptr(gd), ptr(md), ptr(vd), n, n_decay, beta, clip, ptr(lr),
                                   0.9, 0.999, 1e-8, 1.0, ptr(denom), ptr(state), ptr(ws), ws_bytes,
                                   stream()))
        torch.cuda.synchronize()
        assert abs(state[2].item() - norm.item()) < 1e-5 * norm.item()
        want = torch.cat(params).numpy()
        assert np.abs(pd.cpu().numpy() - want).max() < 2e-6
    assert abs(state[0].item() - 0.9 ** 4) < 1e-6
    out = torch.zeros(1, device="cuda")
    check(lib.dincae_l2_half_sumsq(ptr(pd), n_decay, ptr(out), ptr(ws), ws_bytes, stream()))
    torch.cuda.synchronize()
    want = 0.5 * float((torch.cat(params)[:n_decay].double() ** 2).sum())
    assert abs(out.item() - want) < 1e-5 * want


@pytest.mark.parametrize("mean_f32,big_endian", [(0, 1), (1, 1), (0, 0)])
def test_head_recon_records_equals_head_recon_plus_savesample(mean_f32, big_endian):
    """dincae_head_recon_records == dincae_head_recon followed by the numpy arithmetic of
    savesample (DINCAE/__init__.py:237-248, 288-291), BIT-EXACT, as NetCDF-3 record bytes."""
    from dincae_b200 import _lib
    from dincae_b200._lib import check, ptr
    lib = _lib.load()
    rs = np.random.RandomState(3 + mean_f32)
    B, H, W = 3, 9, 14
    xrec = torch.from_numpy(rs.standard_normal((B, 1, H, W, 4)).astype("f4") * 3).cuda()
    md = (rs.standard_normal((H, W)) * 5 + 15)
    md = md.astype("f4") if mean_f32 else md
    never = rs.uniform(size=(H, W)) < 0.2
    mr = torch.zeros(B, H, W, device="cuda")
    sr = torch.zeros(B, H, W, device="cuda")
    check(lib.dincae_head_recon(ptr(xrec), B, H, W, ptr(mr), ptr(sr), stream()))
    rec = torch.zeros(B, 2, H, W, dtype=torch.int32, device="cuda")
    md_dev = torch.from_numpy(md.astype("f8")).cuda()
    nv_dev = torch.from_numpy(never.view(np.uint8)).cuda()
    check(lib.dincae_head_recon_records(ptr(xrec), B, H, W, ptr(md_dev), ptr(nv_dev), -9999.0, mean_f32,
                                        big_endian, ptr(rec), stream()))
    torch.cuda.synchronize()
    got = rec.cpu().numpy().view(">f4" if big_endian else "<f4")
    want_mean = (mr.cpu().numpy() + md[None]).astype("f4")       # numpy result type: f8 or f4
    want_sigma = np.sqrt(sr.cpu().numpy())
    want_mean[:, never] = -9999.0
    want_sigma[:, never] = -9999.0
    assert np.array_equal(got[:, 0].astype("f4").view(np.uint32), want_mean.view(np.uint32))
    assert np.array_equal(got[:, 1].astype("f4").view(np.uint32), want_sigma.view(np.uint32))
    # no mask at all
    check(lib.dincae_head_recon_records(ptr(xrec), B, H, W, ptr(md_dev), None, -9999.0, mean_f32,
                                        big_endian, ptr(rec), stream()))
    torch.cuda.synchronize()
    got = rec.cpu().numpy().view(">f4" if big_endian else "<f4")
    assert np.array_equal(got[:, 1].astype("f4"), np.sqrt(sr.cpu().numpy()))
