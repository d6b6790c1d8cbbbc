"""3x3 conv kernels (forward / data gradient / weight gradient, with the fused pool,
upsample, concat and activation gradients) vs PyTorch-CPU autograd of the TF graph blocks
DINCAE/__init__.py:442-452 and :496-513.  fp32 FFMA path: tolerance 2e-5 of the max."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dincae_b200 import _lib, layout
from dincae_b200._lib import check, ptr
from tests.util import dev, from_p4, p4, rel_err, tf32_round

pytestmark = pytest.mark.gpu
SLOPE = 0.2
# fp32 FFMA kernels: rounding-level agreement.  tcgen05 kind::tf32: operands rounded to 10
# mantissa bits (2^-11 relative each), fp32 accumulation -> 2e-3 of the tensor's max.
TOLS = {0: 2e-5, 1: 2e-3}


def feed(t, backend):
    """Tensors handed to the tensor-core back end must be TF32-representable (the library's
    own producers round on store, include/dincae_b200.h); the test data follows the contract
    so that the oracle sees the same values."""
    return tf32_round(t) if backend == 1 else t


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


# the last two are BASELINE configs[3] widths (encoder [32,48,72,108,162]): MMA N of 112 / 176
ENC_CASES = [(2, 12, 20, 10, 16), (3, 15, 11, 16, 24), (2, 7, 7, 36, 54), (1, 112, 112, 10, 16),
             (2, 33, 70, 5, 3), (1, 20, 24, 72, 108), (1, 18, 16, 108, 162),
             # wider than 128 columns: the tensor-core weight gradient cuts rows into column tiles
             (1, 20, 200, 10, 16), (2, 9, 131, 5, 3)]


@pytest.mark.parametrize("B,H,W,Cin,Cout", ENC_CASES)
def test_encoder_stage_forward_and_backward(B, H, W, Cin, Cout, backend):
    lib = _lib.load()
    TOL = TOLS[backend]
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + Cin)
    x = feed(torch.randn(B, H, W, Cin, generator=g), backend)
    k = torch.randn(3, 3, Cin, Cout, generator=g) * 0.2
    b = torch.randn(Cout, generator=g) * 0.1
    Hh, Wh = (H + 1) // 2, (W + 1) // 2
    gpool = torch.randn(B, Hh, Wh, Cout, generator=g)
    dx_add = torch.randn(B, H, W, Cin, generator=g)

    xr = x.clone().requires_grad_(True)
    kr = k.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    z = conv_ref(xr, kr, br)
    a = F.leaky_relu(z, SLOPE)
    pooled = pool_ref(a)
    (pooled * gpool).sum().backward()

    cinp, coutp = layout.planes(Cin), layout.planes(Cout)
    cpad = layout.round_up(Cout, 16)
    wp = dev(layout.pack_conv(k.numpy(), [Cin], cpad))
    bias = torch.zeros(cpad, device="cuda")
    bias[:Cout] = b.cuda()
    xd = p4(x.numpy())
    out_full = torch.zeros(B, coutp, H, W, 4, device="cuda")
    out_pool = torch.zeros(B, coutp, Hh, Wh, 4, device="cuda")
    out_sign = torch.zeros(B, coutp, H, (W + 3) // 4, dtype=torch.int16, device="cuda")
    check(lib.dincae_conv3x3_fwd(ptr(xd), cinp, 0, None, 0, ptr(wp), ptr(bias), cpad, B, H, W,
                                 coutp, SLOPE, ptr(out_full), ptr(out_pool), ptr(out_sign), stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p4(out_full, Cout), a.detach().numpy()) < TOL
    assert rel_err(from_p4(out_pool, Cout), pooled.detach().numpy()) < TOL
    # padded lanes stay exactly zero
    if Cout % 4:
        assert float(out_pool[:, -1, :, :, Cout % 4:].abs().max()) == 0.0
    # sign mask: exact except where |z| is at rounding level
    want_sign = layout.sign_mask(z.detach().numpy())
    got_sign = out_sign.cpu().numpy().view(np.uint16)
    diff = want_sign ^ got_sign
    zabs = np.abs(z.detach().numpy())
    nflip = sum(bin(int(v)).count("1") for v in diff.reshape(-1) if v)
    assert nflip <= np.count_nonzero(zabs < TOL * zabs.max())

    # backward through pool + leaky relu; the mask of the REFERENCE pre-activation is used so
    # that a rounding-level sign flip of the forward does not leak into this comparison
    out_sign = dev(want_sign.view(np.int16))
    gp = p4(gpool.numpy())
    dx = torch.zeros(B, cinp, H, W, 4, device="cuda")
    add = p4(dx_add.numpy())
    check(lib.dincae_conv3x3_dgrad(None, ptr(gp), ptr(out_sign), SLOPE, coutp, ptr(wp), cinp, cpad,
                                   B, H, W, 0, None, 1.0, None, cinp, ptr(dx), ptr(add), stream()))
    ws_bytes = lib.dincae_conv3x3_wgrad_workspace(cinp, 0, cpad, B, H, W)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((9, cinp, cpad, 4), 7.0, device="cuda")
    db = torch.full((cpad,), 7.0, device="cuda")
    check(lib.dincae_conv3x3_wgrad(ptr(xd), cinp, 0, None, 0, None, ptr(gp), ptr(out_sign), SLOPE,
                                   coutp, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws), ws_bytes, stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p4(dx, Cin), (xr.grad + dx_add).numpy()) < TOL
    assert rel_err(layout.unpack_conv(dw.cpu().numpy(), [Cin], Cout), kr.grad.numpy()) < TOL
    assert rel_err(db.cpu().numpy()[:Cout], br.grad.numpy()) < TOL
    # padding of the packed gradient is exactly zero (keeps padded weights at zero forever)
    full = dw.cpu().numpy().copy()
    idx, _ = layout.cin_map([Cin])
    full[:, idx // 4, :Cout, idx % 4] = 0
    assert np.abs(full).max() == 0.0
    assert np.abs(db.cpu().numpy()[Cout:]).max() == 0.0 if cpad > Cout else True


DEC_CASES = [(2, 14, 14, 54, 36, 36, True), (3, 15, 11, 24, 16, 16, True), (2, 28, 28, 36, 24, 24, False),
             (1, 112, 112, 16, 10, 2, True), (2, 7, 9, 6, 0, 5, True),
             # configs[3] decoder widths: 180 -> 72 (data-gradient N = 192), 270 -> 108 (N = 272:
             # wider than one MMA, FFMA path)
             (1, 18, 16, 108, 72, 72, True), (1, 16, 24, 162, 108, 108, True),
             (1, 12, 160, 16, 10, 2, True), (1, 10, 258, 32, 28, 2, False)]


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,want_skip_grad", DEC_CASES)
def test_decoder_stage_forward_and_backward(B, H, W, C0, C1, Cout, want_skip_grad, backend):
    lib = _lib.load()
    TOL = TOLS[backend]
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + C0)
    Hh, Wh = (H + 1) // 2, (W + 1) // 2
    prev_pre = torch.randn(B, Hh, Wh, C0, generator=g)
    skip = feed(torch.randn(B, H, W, C1, generator=g), backend) if C1 else None
    k = torch.randn(3, 3, C0 + C1, Cout, generator=g) * 0.2
    b = torch.randn(Cout, generator=g) * 0.1
    G = feed(torch.randn(B, H, W, Cout, generator=g), backend)   # gradient w.r.t. the conv pre-activation

    pr = prev_pre.clone().requires_grad_(True)
    kr = k.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    sr = skip.clone().requires_grad_(True) if C1 else None
    d_prev = F.leaky_relu(pr, SLOPE)
    u = up_ref(d_prev, (H, W))
    if C1:
        u = torch.cat([u, sr], dim=3)
    z = conv_ref(u, kr, br)
    act = F.leaky_relu(z, SLOPE)
    (z * G).sum().backward()

    c0p, c1p, coutp = layout.planes(C0), layout.planes(C1), layout.planes(Cout)
    cpad = layout.round_up(Cout, 16)
    splits = [C0] + ([C1] if C1 else [])
    wp = dev(layout.pack_conv(k.numpy(), splits, cpad))
    bias = torch.zeros(cpad, device="cuda")
    bias[:Cout] = b.cuda()
    d_prev_dev = p4(d_prev.detach().numpy())
    skip_dev = p4(skip.numpy()) if C1 else None
    out_full = torch.zeros(B, coutp, H, W, 4, device="cuda")
    check(lib.dincae_conv3x3_fwd(ptr(d_prev_dev), c0p, 1, ptr(skip_dev), c1p, ptr(wp), ptr(bias),
                                 cpad, B, H, W, coutp, SLOPE, ptr(out_full), None, None, stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p4(out_full, Cout), act.detach().numpy()) < TOL

    Gd = p4(G.numpy())
    g_prev = torch.zeros(B, c0p, Hh, Wh, 4, device="cuda")
    use_hi = bool(C1) and want_skip_grad
    dx_hi = torch.zeros(B, max(c1p, 1), H, W, 4, device="cuda")
    check(lib.dincae_conv3x3_dgrad(ptr(Gd), None, None, SLOPE, coutp, ptr(wp), c0p + c1p, cpad,
                                   B, H, W, c0p, ptr(d_prev_dev), SLOPE, ptr(g_prev),
                                   c1p if use_hi else 0, ptr(dx_hi) if use_hi else None, None,
                                   stream()))
    ws_bytes = lib.dincae_conv3x3_wgrad_workspace(c0p, c1p, cpad, B, H, W)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw = torch.full((9, c0p + c1p, cpad, 4), 7.0, device="cuda")
    db = torch.full((cpad,), 7.0, device="cuda")
    check(lib.dincae_conv3x3_wgrad(ptr(d_prev_dev), c0p, 1, ptr(skip_dev), c1p, ptr(Gd), None, None,
                                   SLOPE, coutp, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws), ws_bytes,
                                   stream()))
    torch.cuda.synchronize()
    assert rel_err(from_p4(g_prev, C0), pr.grad.numpy()) < TOL
    if use_hi:
        assert rel_err(from_p4(dx_hi, C1), sr.grad.numpy()) < TOL
    assert rel_err(layout.unpack_conv(dw.cpu().numpy(), splits, Cout), kr.grad.numpy()) < TOL
    assert rel_err(db.cpu().numpy()[:Cout], br.grad.numpy()) < TOL


def test_relu_bottleneck_slope_and_no_prev_act(backend):
    """prev_slope = 0 (dense ReLU below the first decoder conv) and prev_act = NULL."""
    lib = _lib.load()
    TOL = TOLS[backend]
    g = torch.Generator().manual_seed(3)
    B, H, W, C0, Cout = 2, 6, 10, 8, 4
    Hh, Wh = 3, 5
    pre = torch.randn(B, Hh, Wh, C0, generator=g)
    k = torch.randn(3, 3, C0, Cout, generator=g)
    G = feed(torch.randn(B, H, W, Cout, generator=g), backend)
    pr = pre.clone().requires_grad_(True)
    act = F.relu(pr)
    (conv_ref(up_ref(act, (H, W)), k) * G).sum().backward()
    ar = act.detach().clone().requires_grad_(True)
    (conv_ref(up_ref(ar, (H, W)), k) * G).sum().backward()
    cpad = 16
    wp = dev(layout.pack_conv(k.numpy(), [C0], cpad))
    act_dev = p4(act.detach().numpy())
    Gd = p4(G.numpy())
    for prev, slope, want in ((act_dev, 0.0, pr.grad), (None, 1.0, ar.grad)):
        g_prev = torch.zeros(B, 2, Hh, Wh, 4, device="cuda")
        check(lib.dincae_conv3x3_dgrad(ptr(Gd), None, None, SLOPE, 1, ptr(wp), 2, cpad, B, H, W,
                                       2, ptr(prev), slope, ptr(g_prev), 0, None, None, stream()))
        torch.cuda.synchronize()
        assert rel_err(from_p4(g_prev, C0), want.numpy()) < TOL


def test_wgrad_in_two_steps_equals_the_fused_call(backend):
    """dincae_conv3x3_wgrad_partial + _reduce (for callers that put the reduction on another
    stream) give bit-identical dw / db to dincae_conv3x3_wgrad."""
    import ctypes
    lib = _lib.load()
    g = torch.Generator().manual_seed(9)
    B, H, W, Cin, Cout = 3, 20, 24, 10, 16
    x = feed(torch.randn(B, H, W, Cin, generator=g), backend)
    G = feed(torch.randn(B, H, W, Cout, generator=g), backend)
    cinp, coutp, cpad = layout.planes(Cin), layout.planes(Cout), 16
    xd, Gd = p4(x.numpy()), p4(G.numpy())
    ws_bytes = lib.dincae_conv3x3_wgrad_workspace(cinp, 0, cpad, B, H, W)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw1 = torch.zeros(9, cinp, cpad, 4, device="cuda")
    db1 = torch.zeros(cpad, device="cuda")
    check(lib.dincae_conv3x3_wgrad(ptr(xd), cinp, 0, None, 0, ptr(Gd), None, None, SLOPE, coutp, cpad,
                                   B, H, W, ptr(dw1), ptr(db1), ptr(ws), ws_bytes, stream()))
    dw2, db2 = torch.zeros_like(dw1), torch.zeros_like(db1)
    n = ctypes.c_int(0)
    check(lib.dincae_conv3x3_wgrad_partial(ptr(xd), cinp, 0, None, 0, ptr(Gd), None, None, SLOPE, coutp,
                                           cpad, B, H, W, ptr(ws), ws_bytes, ctypes.byref(n), stream()))
    assert n.value >= 1
    check(lib.dincae_conv3x3_wgrad_reduce(ptr(ws), n.value, cinp, coutp, cpad, ptr(dw2), ptr(db2), stream()))
    torch.cuda.synchronize()
    assert torch.equal(dw1, dw2) and torch.equal(db1, db2)
    assert lib.dincae_conv3x3_wgrad_reduce(None, 1, cinp, coutp, cpad, ptr(dw2), ptr(db2), stream()) == -1


def test_invalid_arguments_are_rejected():
    lib = _lib.load()
    x = torch.zeros(16, device="cuda")
    assert lib.dincae_conv3x3_fwd(None, 1, 0, None, 0, ptr(x), ptr(x), 16, 1, 2, 2, 1, 0.2,
                                  ptr(x), None, None, stream()) == -1
    assert lib.dincae_conv3x3_fwd(ptr(x), 1, 0, None, 0, ptr(x), ptr(x), 16, 1, 2, 2, 1, 0.2,
                                  None, None, None, stream()) == -1
    assert lib.dincae_conv3x3_fwd(ptr(x), 1, 0, None, 0, ptr(x), ptr(x), 16, 1, 2, 2, 5, 0.2,
                                  ptr(x), None, None, stream()) == -1


def test_wide_images_take_the_tensor_core_weight_gradient():
    """Plan check (no device work): the column-tiled tcgen05 weight gradient covers the wide
    parity cases above and the 512x512 / 256x256 layers of BASELINE configs[3]."""
    lib = _lib.load()
    on = lib.dincae_conv3x3_wgrad_on_tensor_cores
    assert on(3, 4, 1, 20, 200) == 1 and on(2, 1, 2, 9, 131) == 1          # ENC_CASES (planes)
    assert on(4 + 3, 1, 1, 12, 160) == 1 and on(8 + 7, 1, 1, 10, 258) == 1  # DEC_CASES
    assert on(7, 8, 8, 512, 512) == 1 and on(8, 12, 8, 256, 256) == 1       # config D enc 1, enc 2
    assert on(20, 8, 8, 256, 256) == 1 and on(15, 1, 8, 512, 512) == 1      # config D dec 4, dec 5
    assert on(3, 4, 50, 112, 112) == 1                                      # config A enc 1
    assert on(68, 27, 8, 32, 32) == 0                                       # 270 -> 108: FFMA kernel
