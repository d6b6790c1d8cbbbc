"""bf16 3x3 conv kernels (tcgen05 kind::f16: forward / data gradient / weight gradient with the
fused pool, upsample, concat and activation gradients) vs PyTorch-CPU fp32 autograd of the TF
graph blocks DINCAE/__init__.py:442-452 and :496-513.

Tolerances (stated per north_star): inputs, weights and incoming gradients are bf16-exact
test data, products accumulate in fp32, so the only rounding is the bf16 store of an output
(2^-9 relative per element) and of an un-pooled gradient inside the kernels: 1e-2 of the
tensor's maximum for bf16 outputs, 5e-3 for the fp32 weight gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dincae_b200 import _lib, layout
from dincae_b200._lib import check, ptr
from tests.util import rel_err

pytestmark = pytest.mark.gpu
SLOPE = 0.2
TOL_ACT = 1e-2
TOL_W = 5e-3


def bf(t):
    """fp32 tensor -> nearest bf16-representable fp32 tensor."""
    return t.to(torch.bfloat16).to(torch.float32)


def p8(x_nhwc):
    return torch.from_numpy(layout.nhwc_to_p8(np.asarray(x_nhwc, dtype=np.float32))).cuda().to(torch.bfloat16)


def from_p8(t, C):
    return layout.p8_to_nhwc(t.detach().float().cpu().numpy(), C)


def nchw(x):
    return x.permute(0, 3, 1, 2)


def nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv_ref(u, k, b=None):
    return nhwc(F.conv2d(nchw(u), k.permute(3, 2, 0, 1), b, padding=1))


def pool_ref(a):
    return nhwc(F.avg_pool2d(nchw(a), 2, 2, ceil_mode=True, count_include_pad=False))


def up_ref(d, size):
    return nhwc(F.interpolate(nchw(d), size=size, mode="nearest"))


def stream():
    return torch.cuda.current_stream().cuda_stream


def pack_weights(lib, k, splits, cout_pad, nd, ci_first=0):
    """fp32 master in WP layout (sources padded to 8 channels) + the two bf16 operand copies
    (the data-gradient copy for the nd input channels from ci_first on)."""
    wp = torch.from_numpy(layout.pack_conv(k.numpy(), splits, cout_pad, lanes=8)).cuda()
    cinp4 = wp.shape[1]
    import ctypes
    nde = ctypes.c_size_t(0)
    nfe = lib.dincae_conv_weights_bf16_elems(cinp4, cout_pad, nd, ctypes.byref(nde))
    wf = torch.zeros(nfe, dtype=torch.bfloat16, device="cuda")
    wd = torch.zeros(nde.value, dtype=torch.bfloat16, device="cuda")
    check(lib.dincae_conv_weights_bf16(ptr(wp), cinp4, cout_pad, ci_first, nd, ptr(wf), ptr(wd), stream()))
    return wp, wf, wd, cinp4


ENC_CASES = [(2, 12, 20, 10, 16), (3, 15, 11, 16, 24), (2, 7, 7, 36, 54), (1, 112, 112, 10, 16),
             (2, 33, 70, 5, 3), (1, 20, 24, 72, 108), (1, 18, 16, 108, 162),
             (1, 20, 200, 28, 32), (2, 9, 131, 5, 3), (1, 40, 136, 32, 48),
             # input planes fill the weight-gradient M tile exactly: bias row as its own all-ones MMA
             (1, 16, 24, 64, 40), (1, 12, 16, 128, 24)]


@pytest.mark.parametrize("B,H,W,Cin,Cout", ENC_CASES)
def test_encoder_stage_bf16(B, H, W, Cin, Cout):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    x = bf(torch.randn(B, H, W, Cin, generator=g))
    k = bf(torch.randn(3, 3, Cin, Cout, generator=g) * 0.2)
    b = torch.randn(Cout, generator=g) * 0.1
    Hh, Wh = (H + 1) // 2, (W + 1) // 2
    gpool = bf(torch.randn(B, Hh, Wh, Cout, generator=g))
    dx_add = bf(torch.randn(B, H, W, Cin, generator=g))

    xr = x.clone().requires_grad_(True)
    kr = k.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    z = conv_ref(xr, kr, br)
    a = F.leaky_relu(z, SLOPE)
    pooled = pool_ref(a)
    (pooled * gpool).sum().backward()

    cin8, co8 = layout.planes8(Cin), layout.planes8(Cout)
    cpad = layout.round_up(Cout, 16)
    nd = layout.round_up(8 * cin8, 16)
    wp, wf, wd, cinp4 = pack_weights(lib, k, [Cin], cpad, nd)
    bias = torch.zeros(cpad, device="cuda")
    bias[:Cout] = b.cuda()
    xd = p8(x.numpy())
    out_full = torch.zeros(B, co8, H, W, 8, dtype=torch.bfloat16, device="cuda")
    out_pool = torch.zeros(B, co8, Hh, Wh, 8, dtype=torch.bfloat16, device="cuda")
    wq4 = 4 * ((W + 3) // 4)
    out_sign = torch.zeros(B, co8, H, wq4, dtype=torch.uint8, device="cuda")
    c4 = layout.planes(Cout)
    out_f32 = torch.zeros(B, c4, H, W, 4, device="cuda")
    check(lib.dincae_conv3x3_fwd_bf16(ptr(xd), cin8, 0, None, 0, ptr(wf), ptr(bias), cpad, B, H, W, co8,
                                      SLOPE, ptr(out_full), ptr(out_pool), ptr(out_sign),
                                      ptr(out_f32), c4, stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p8(out_full, Cout), a.detach().numpy()) < TOL_ACT
    assert rel_err(from_p8(out_pool, Cout), pooled.detach().numpy()) < TOL_ACT
    assert rel_err(layout.p4_to_nhwc(out_f32.cpu().numpy(), Cout), a.detach().numpy()) < 1e-4
    if Cout % 8:
        assert float(out_pool[:, -1, :, :, Cout % 8:].float().abs().max()) == 0.0
    want_sign = layout.sign_mask8(z.detach().numpy())
    got_sign = out_sign.cpu().numpy()
    diff = want_sign ^ got_sign
    zabs = np.abs(z.detach().numpy())
    nflip = int(np.unpackbits(diff).sum())
    assert nflip <= np.count_nonzero(zabs < 1e-5 * zabs.max())

    # backward through pool + leaky relu with the REFERENCE sign mask
    sign = torch.from_numpy(want_sign).cuda()
    gp = p8(gpool.numpy())
    dx = torch.zeros(B, cin8, H, W, 8, dtype=torch.bfloat16, device="cuda")
    add = p8(dx_add.numpy())
    check(lib.dincae_conv3x3_dgrad_bf16(None, ptr(gp), ptr(sign), SLOPE, co8, ptr(wd), nd, B, H, W,
                                        0, None, 1.0, None, cin8, ptr(dx), ptr(add), stream()))
    ws_bytes = lib.dincae_conv3x3_wgrad_bf16_workspace(cinp4, cpad)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((9, cinp4, cpad, 4), 7.0, device="cuda")
    db = torch.full((cpad,), 7.0, device="cuda")
    check(lib.dincae_conv3x3_wgrad_bf16(ptr(xd), cin8, 0, None, 0, None, ptr(gp), ptr(sign), SLOPE,
                                        co8, cinp4, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws), ws_bytes,
                                        stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p8(dx, Cin), (xr.grad + dx_add).numpy()) < TOL_ACT
    assert rel_err(layout.unpack_conv(dw.cpu().numpy(), [Cin], Cout, lanes=8), kr.grad.numpy()) < TOL_W
    assert rel_err(db.cpu().numpy()[:Cout], br.grad.numpy()) < TOL_W
    # padding of the packed gradient is exactly zero (keeps padded weights at zero forever)
    full = dw.cpu().numpy().copy()
    idx, _ = layout.cin_map([Cin], lanes=8)
    full[:, idx // 4, :Cout, idx % 4] = 0
    assert np.abs(full).max() == 0.0
    if cpad > Cout:
        assert np.abs(db.cpu().numpy()[Cout:]).max() == 0.0


DEC_CASES = [(2, 14, 14, 54, 36, 36, True), (3, 15, 11, 24, 16, 16, True), (2, 28, 28, 36, 24, 24, False),
             (1, 112, 112, 16, 10, 2, True), (2, 7, 9, 6, 0, 5, True),
             (1, 18, 16, 108, 72, 72, True), (1, 16, 24, 162, 108, 108, True),
             (1, 12, 160, 16, 10, 2, True), (1, 10, 258, 32, 28, 2, False), (1, 24, 130, 48, 32, 32, True)]


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,want_skip_grad", DEC_CASES)
def test_decoder_stage_bf16(B, H, W, C0, C1, Cout, want_skip_grad):
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + C0)
    Hh, Wh = (H + 1) // 2, (W + 1) // 2
    d_prev = bf(F.leaky_relu(torch.randn(B, Hh, Wh, C0, generator=g), SLOPE))   # stored activation
    skip = bf(torch.randn(B, H, W, C1, generator=g)) if C1 else None
    k = bf(torch.randn(3, 3, C0 + C1, Cout, generator=g) * 0.2)
    b = torch.randn(Cout, generator=g) * 0.1
    G = bf(torch.randn(B, H, W, Cout, generator=g))

    dr = d_prev.clone().requires_grad_(True)
    kr = k.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    sr = skip.clone().requires_grad_(True) if C1 else None
    u = up_ref(dr, (H, W))
    if C1:
        u = torch.cat([u, sr], dim=3)
    z = conv_ref(u, kr, br)
    act = F.leaky_relu(z, SLOPE)
    (z * G).sum().backward()
    # gradient w.r.t. the previous PRE-activation = grad(d_prev) * lrelu'(sign of d_prev)
    want_prev = dr.grad * torch.where(d_prev > 0, torch.ones(()), torch.full((), SLOPE))

    c0p, c1p, co8 = layout.planes8(C0), layout.planes8(C1), layout.planes8(Cout)
    cpad = layout.round_up(Cout, 16)
    splits = [C0] + ([C1] if C1 else [])
    use_hi = bool(C1) and want_skip_grad
    nd = layout.round_up(8 * (c0p + (c1p if use_hi else 0)), 16)
    wp, wf, wd, cinp4 = pack_weights(lib, k, splits, cpad, min(nd, 256))
    bias = torch.zeros(cpad, device="cuda")
    bias[:Cout] = b.cuda()
    d_prev_dev = p8(d_prev.numpy())
    skip_dev = p8(skip.numpy()) if C1 else None
    out_full = torch.zeros(B, co8, H, W, 8, dtype=torch.bfloat16, device="cuda")
    check(lib.dincae_conv3x3_fwd_bf16(ptr(d_prev_dev), c0p, 1, ptr(skip_dev), c1p, ptr(wf), ptr(bias),
                                      cpad, B, H, W, co8, SLOPE, ptr(out_full), None, None, None, 0,
                                      stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p8(out_full, Cout), act.detach().numpy()) < TOL_ACT

    Gd = p8(G.numpy())
    g_prev = torch.zeros(B, c0p, Hh, Wh, 8, dtype=torch.bfloat16, device="cuda")
    dx_hi = torch.zeros(B, max(c1p, 1), H, W, 8, dtype=torch.bfloat16, device="cuda")
    if nd <= 256:
        check(lib.dincae_conv3x3_dgrad_bf16(ptr(Gd), None, None, SLOPE, co8, ptr(wd), nd, B, H, W,
                                            c0p, ptr(d_prev_dev), SLOPE, ptr(g_prev),
                                            c1p if use_hi else 0, ptr(dx_hi) if use_hi else None, None,
                                            stream()))
    else:
        # more input channels than one MMA is wide (270 -> 108 of configs[3]): one call per range
        nd0, nd1 = layout.round_up(8 * c0p, 16), layout.round_up(8 * c1p, 16)
        _, _, wd0, _ = pack_weights(lib, k, splits, cpad, nd0, 0)
        _, _, wd1, _ = pack_weights(lib, k, splits, cpad, nd1, 8 * c0p)
        check(lib.dincae_conv3x3_dgrad_bf16(ptr(Gd), None, None, SLOPE, co8, ptr(wd0), nd0, B, H, W,
                                            c0p, ptr(d_prev_dev), SLOPE, ptr(g_prev), 0, None, None,
                                            stream()))
        check(lib.dincae_conv3x3_dgrad_bf16(ptr(Gd), None, None, SLOPE, co8, ptr(wd1), nd1, B, H, W,
                                            0, None, 1.0, None, c1p, ptr(dx_hi), None, stream()))
    ws_bytes = lib.dincae_conv3x3_wgrad_bf16_workspace(cinp4, cpad)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((9, cinp4, cpad, 4), 7.0, device="cuda")
    db = torch.full((cpad,), 7.0, device="cuda")
    check(lib.dincae_conv3x3_wgrad_bf16(ptr(d_prev_dev), c0p, 1, ptr(skip_dev), c1p, ptr(Gd), None, None,
                                        SLOPE, co8, cinp4, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws),
                                        ws_bytes, stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p8(g_prev, C0), want_prev.numpy()) < TOL_ACT
    if use_hi:
        assert rel_err(from_p8(dx_hi, C1), sr.grad.numpy()) < TOL_ACT
    assert rel_err(layout.unpack_conv(dw.cpu().numpy(), splits, Cout, lanes=8), kr.grad.numpy()) < TOL_W
    assert rel_err(db.cpu().numpy()[:Cout], br.grad.numpy()) < TOL_W


def test_layout_casts_round_trip():
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 5, 7, 11
    x = bf(torch.randn(B, H, W, C, generator=g))
    x4 = torch.from_numpy(layout.nhwc_to_p4(x.numpy())).cuda()
    c4, c8 = layout.planes(C), layout.planes8(C)
    y8 = torch.full((B, c8, H, W, 8), 3.0, dtype=torch.bfloat16, device="cuda")
    check(lib.dincae_p4_to_p8(ptr(x4), c4, c8, B, H, W, ptr(y8), stream()))
    back = torch.full((B, c4, H, W, 4), 5.0, device="cuda")
    check(lib.dincae_p8_to_p4(ptr(y8), c4, c8, B, H, W, ptr(back), stream()))
    torch.cuda.synchronize()
    assert np.array_equal(from_p8(y8, C), x.numpy())
    assert float(y8[:, -1, :, :, C % 8:].float().abs().max()) == 0.0
    assert torch.equal(back, x4)
