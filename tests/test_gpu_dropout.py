"""Optional dropout of the dense bottleneck (csrc/dropout.cu; tf.layers.dropout of
DINCAE/__init__.py:483, inert in the reference and therefore off by default): the forward mask is a
deterministic function of (seed, layer, frame key, column), and the backward pass -- ReLU gate of
the existing kernels + one scale -- equals autograd through relu -> dropout with the same mask."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dincae_b200 import _lib, layout
from dincae_b200._lib import check, ptr
from dincae_b200.engine import Plan, make_network
from oracle import model_torch as M

pytestmark = pytest.mark.gpu


def stream():
    return torch.cuda.current_stream().cuda_stream


def test_dropout_forward_mask_properties():
    lib = _lib.load()
    B, N, ld, rate = 7, 1000, 1004, 0.3
    y0 = torch.rand(B, ld, device="cuda") + 0.5
    keys = torch.arange(100, 100 + B, dtype=torch.int64, device="cuda")
    outs = []
    for step, kk in ((0, keys), (0, keys), (1, keys), (0, keys + 1)):
        y = y0.clone()
        mask = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
        check(lib.dincae_dropout_fwd(ptr(y), B, N, ld, rate, 1234, step, 0, ptr(kk), ptr(mask), stream()))
        torch.cuda.synchronize()
        assert torch.equal(y[:, :N], torch.where(mask.bool(), y0[:, :N] / (1 - rate), torch.zeros(()).cuda()))
        assert torch.equal(y[:, N:], y0[:, N:])                         # row padding untouched
        outs.append(mask.cpu())
    assert torch.equal(outs[0], outs[1])                                  # deterministic
    assert not torch.equal(outs[0], outs[2]) and not torch.equal(outs[0], outs[3])
    assert torch.equal(outs[0][1:], outs[3][:-1])                         # the mask follows the frame key
    frac = 1.0 - outs[0].float().mean().item()
    assert abs(frac - rate) < 0.03


def test_network_gradients_with_dropout_match_autograd():
    c = dict(B=3, H=16, W=16, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4])
    lib = _lib.load()
    lib.dincae_set_conv_backend(0)
    try:
        arch = M.Arch(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
        plan = Plan(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
        g = torch.Generator().manual_seed(1)
        params = M.init_params(arch, 1)
        B, H, W, nvar = c["B"], c["H"], c["W"], c["nvar"]
        xin = torch.randn(B, H, W, nvar, generator=g)
        obs = torch.rand(B, H, W, generator=g) > 0.5
        xtrue = torch.zeros(B, H, W, 2)
        xtrue[..., 1] = obs.float()
        xtrue[..., 0] = torch.randn(B, H, W, generator=g) * obs
        rate = 0.3
        net = make_network(plan, B, train=True)
        net.set_params_tf([p.numpy() for p in params])
        net.dropout_rate = rate
        net.dropout_seed = 99
        net.dropout_keys = torch.arange(B, dtype=torch.int64, device="cuda")
        nd = len(plan.dense_units)
        net.dropout_masks = [torch.zeros(B, plan.dense_pad[i + 1], dtype=torch.uint8, device="cuda") for i in range(nd)]
        net.X[0][:B].copy_(torch.from_numpy(layout.nhwc_to_p4(xin.numpy())))
        net.xtrue[:B].copy_(xtrue)
        net.n_noncloud.fill_(int((xtrue[..., 1] != 0).sum()))
        net.forward(B)
        net.loss_backward(B, False, normalise=True)
        torch.cuda.synchronize()
        got = net.get_grads_tf()

        # oracle: the same graph with the SAME masks applied after each dense ReLU (float64)
        ps = [p.double().requires_grad_(True) for p in params]
        masks = []
        L = arch.nlayers
        hL, wL, cL = plan.sizes[L][0], plan.sizes[L][1], plan.enc_c[L]
        for i in range(nd):
            mk = net.dropout_masks[i].cpu().double()
            if i == nd - 1:                       # planar-flatten order -> NHWC flatten order
                mk = mk[:, torch.from_numpy(plan.flat_map)]
            else:
                mk = mk[:, :plan.dense_units[i]]
            masks.append(mk / (1 - rate))
        it = iter(ps)
        enc_pool = [xin.double()]
        enc_conv = [None]
        for l in range(1, L + 1):
            k, b = next(it), next(it)
            cc = F.leaky_relu(M._conv(enc_pool[l - 1], k, b), 0.2)
            enc_conv.append(cc)
            enc_pool.append(M._pool(cc))
        x = enc_pool[L]
        shp = x.shape
        x = x.reshape(shp[0], -1)
        for i in range(nd):
            k, b = next(it), next(it)
            x = F.relu(x @ k + b) * masks[i]
        dec = [x.reshape(shp)]
        for l in range(1, L + 1):
            l2 = L + 1 - l
            up = M._resize_nearest(dec[l - 1], enc_conv[l2].shape[1:3])
            if l in arch.skipconnections:
                up = torch.cat([up, enc_pool[l2 - 1]], dim=3)
            k, b = next(it), next(it)
            dec.append(F.leaky_relu(M._conv(up, k, b), 0.2))
        cost, rms, n = M.cost_terms(dec[L], xtrue.double(), False)
        cost.backward()
        assert abs(net.loss_out[0].item() - cost.item()) < 1e-4 * max(1.0, abs(cost.item()))
        for (nm, _), gg, p in zip(arch.param_shapes(), got, ps):
            want = p.grad.numpy()
            scale = max(np.abs(want).max(), 1e-9)
            assert np.abs(gg - want).max() / scale < 2e-4, nm
    finally:
        lib.dincae_set_conv_backend(1)
