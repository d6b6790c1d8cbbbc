"""Known-answer tests of the model oracle (oracle/model_torch.py).

TensorFlow 1.15 cannot run here and the reference holds no tensor-level golden vectors
(SURVEY.md section 8c), so the torch-CPU restatement of the graph is pinned op by op against
answers worked out BY HAND from the published TF-1.15 op semantics the reference relies on
(DINCAE/__init__.py:433-603) and against the closed forms / counts derived in SURVEY.md 8a/8d:
SAME convolution as cross-correlation with HWIO kernels, SAME average pooling that leaves
padded cells out of the mean, the legacy nearest-neighbour resize, concatenating skips, the
NHWC flatten of the bottleneck, the head, the loss and its closed-form gradient, global-norm
clipping and TF's Adam.  A drift guard replays a committed fp64 evaluation
(tests/golden/model_small.npz, written by oracle/make_golden_model.py).
"""
import math
import os

import numpy as np
import torch

from oracle import model_torch as M

F64 = torch.float64


def test_conv_is_same_cross_correlation_with_hwio_kernels():
    """tf.layers.conv2d, padding SAME, stride 1 (:442): out[y,x,o] = b[o] + sum k[dy,dx,i,o] *
    in[y+dy-1, x+dx-1, i], zeros outside."""
    x = torch.arange(12, dtype=F64).reshape(1, 3, 4, 1) + 1
    k = torch.zeros(3, 3, 1, 2, dtype=F64)
    k[0, 2, 0, 0] = 1.0          # tap (dy=0, dx=2): out[y,x] = in[y-1, x+1]
    k[1, 1, 0, 1] = 2.0          # centre tap, second output channel
    b = torch.tensor([0.5, -1.0], dtype=F64)
    y = M._conv(x, k, b)
    img = x[0, :, :, 0]
    want0 = torch.zeros(3, 4, dtype=F64)
    want0[1:, :3] = img[:2, 1:]
    assert torch.equal(y[0, :, :, 0], want0 + 0.5)
    assert torch.equal(y[0, :, :, 1], 2 * img - 1.0)
    # two input channels are summed
    x2 = torch.stack([img, 10 * img], dim=-1)[None]
    k2 = torch.zeros(3, 3, 2, 1, dtype=F64)
    k2[1, 1, 0, 0] = 1.0
    k2[1, 1, 1, 0] = 1.0
    assert torch.equal(M._conv(x2, k2, torch.zeros(1, dtype=F64))[0, :, :, 0], 11 * img)


def test_same_average_pool_excludes_padding():
    """average_pooling2d(2, 2, 'same') (:449): out = ceil(in/2); bottom/right windows are
    clipped and divide by the number of in-bounds cells."""
    x = torch.arange(15, dtype=F64).reshape(1, 3, 5, 1)
    y = M._pool(x)[0, :, :, 0]
    a = x[0, :, :, 0]
    want = torch.tensor([[(a[0, 0] + a[0, 1] + a[1, 0] + a[1, 1]) / 4, (a[0, 2] + a[0, 3] + a[1, 2] + a[1, 3]) / 4,
                          (a[0, 4] + a[1, 4]) / 2],
                         [(a[2, 0] + a[2, 1]) / 2, (a[2, 2] + a[2, 3]) / 2, a[2, 4]]], dtype=F64)
    assert y.shape == (2, 3) and torch.equal(y, want)


def test_legacy_nearest_resize_index_rule():
    """tf.image.resize_images(NEAREST_NEIGHBOR), align_corners=False (:496-499): source index
    floor(dst * in / out)."""
    for n_in, n_out in [(7, 14), (7, 13), (4, 7), (2, 3), (1, 2), (56, 112)]:
        x = torch.arange(n_in * n_in, dtype=F64).reshape(1, n_in, n_in, 1)
        y = M._resize_nearest(x, (n_out, n_out))[0, :, :, 0]
        src = [math.floor(d * n_in / n_out) for d in range(n_out)]
        want = x[0, :, :, 0][src][:, src]
        assert torch.equal(y, want), (n_in, n_out)
    # for the sizes the network produces (in = ceil(out/2)) this is dst >> 1
    for n_out in range(1, 130):
        n_in = (n_out + 1) // 2
        assert [math.floor(d * n_in / n_out) for d in range(n_out)] == [d >> 1 for d in range(n_out)]


def test_architecture_counts_match_the_survey():
    """20 variables in TF creation order; 2 881 367 parameters at config A
    (dense 2646 -> 529 -> 2646), 688 723 466 at config D; decoder inputs 90/60/40/26."""
    a = M.Arch(112, 112, 10)
    shapes = a.param_shapes()
    names = [n for n, _ in shapes]
    assert names[:2] == ["conv2d/kernel", "conv2d/bias"] and names[8:12] == \
        ["dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias"]
    assert names[12] == "conv2d_4/kernel" and names[-1] == "conv2d_7/bias" and len(names) == 20
    assert a.dense_units() == [529, 2646]
    assert sum(int(np.prod(s)) for _, s in shapes) == 2881367
    assert [sum(a.dec_cin(l)) for l in range(1, 5)] == [90, 60, 40, 26]
    assert [s[3] for n, s in shapes if n.startswith("conv2d") and n.endswith("kernel")][-4:] == [36, 24, 16, 2]
    d = M.Arch(512, 512, 28, [32, 48, 72, 108, 162], [0.2], [1, 2, 3, 4, 5])
    assert d.dense_units() == [8294, 41472]
    assert sum(int(np.prod(s)) for _, s in d.param_shapes()) == 688723466
    assert [sum(d.dec_cin(l)) for l in range(1, 6)] == [270, 180, 120, 80, 60]


def test_forward_wiring_concat_order_flatten_order_and_inert_dropout():
    """One-layer network small enough to evaluate by hand: skip = CONCAT with the upsampled
    tensor first (:503-506), dense input = NHWC flatten (:465-468), the last conv
    also gets the leaky ReLU (:508-513)."""
    arch = M.Arch(2, 2, 1, [1], [1.0], [1])
    shapes = dict(arch.param_shapes())
    assert shapes["conv2d_1/kernel"] == (3, 3, 2, 2) and shapes["dense/kernel"] == (1, 1)
    x = torch.tensor([[1.0, -2.0], [3.0, 4.0]], dtype=F64).reshape(1, 2, 2, 1)
    ke = torch.zeros(3, 3, 1, 1, dtype=F64)
    ke[1, 1, 0, 0] = 1.0                                   # identity conv
    kd = torch.zeros(3, 3, 2, 2, dtype=F64)
    kd[1, 1, 0, 0] = 1.0                                   # out 0 <- upsampled channel
    kd[1, 1, 1, 1] = -1.0                                  # out 1 <- minus the skip (raw input)
    assert [n for n, _ in arch.param_shapes()][2:6] == ["dense/kernel", "dense/bias", "dense_1/kernel", "dense_1/bias"]
    params = [ke, torch.zeros(1, dtype=F64), torch.tensor([[2.0]], dtype=F64), torch.tensor([0.5], dtype=F64),
              torch.tensor([[1.0]], dtype=F64), torch.zeros(1, dtype=F64),          # dense_1: back to n units
              kd, torch.zeros(2, dtype=F64)]
    keep = {}
    y = M.forward(arch, params, x, keep)
    enc = torch.tensor([[1.0, -0.4], [3.0, 4.0]], dtype=F64)           # leaky_relu(0.2)
    assert torch.allclose(keep["enc_conv"][1][0, :, :, 0], enc, atol=0, rtol=1e-15)
    pooled = enc.mean()
    bott = 2.0 * pooled + 0.5                                          # relu(x W + b) > 0, dropout inert
    assert torch.allclose(keep["dense"][0], bott.reshape(1, 1), rtol=1e-15)
    assert torch.allclose(y[0, :, :, 0], torch.full((2, 2), float(bott), dtype=F64), rtol=1e-15)
    neg = -x[0, :, :, 0]
    assert torch.allclose(y[0, :, :, 1], torch.where(neg > 0, neg, 0.2 * neg), rtol=1e-15)
    # NHWC flatten (C fastest): the first dense layer sees pool[y, x, c] at index (y*W + x)*C + c
    arch2 = M.Arch(4, 4, 1, [2], [1.0], [])
    p2 = M.init_params(arch2, 1, F64)
    x4 = torch.arange(16, dtype=F64).reshape(1, 4, 4, 1) / 7 - 1
    keep2 = {}
    M.forward(arch2, p2, x4, keep2)
    pool = keep2["enc_pool"][1]                                         # [1, 2, 2, 2]
    want = torch.zeros(1, p2[2].shape[1], dtype=F64)
    for yy in range(2):
        for xx in range(2):
            for c in range(2):
                want += pool[0, yy, xx, c] * p2[2][(yy * 2 + xx) * 2 + c][None]
    assert torch.allclose(keep2["dense"][0], torch.relu(want + p2[3]), rtol=1e-13, atol=1e-15)


def test_head_clamps_like_the_reference():
    """(:520-523, sinv :298-299) sigma2 = 1 / max(exp(min(b, 10)), 1e-3), m = a * sigma2."""
    xrec = torch.tensor([[2.0, 0.0], [3.0, 30.0], [-1.0, -30.0], [0.5, math.log(4.0)]], dtype=F64).reshape(1, 2, 2, 2)
    m, s2 = M.head(xrec)
    want_s2 = torch.tensor([1.0, math.exp(-10.0), 1000.0, 0.25], dtype=F64).reshape(1, 2, 2)
    assert torch.allclose(s2, want_s2, rtol=1e-15)
    assert torch.allclose(m, xrec[..., 0] * want_s2, rtol=1e-15)


def test_loss_and_its_closed_form_gradient():
    """Default cost (:558-559) and the truth_uncertain variant (:555-556) against direct
    formulas, and autograd against the closed-form gradient of SURVEY.md 8a row a9
    (tf.minimum / tf.maximum pass the gradient on ties)."""
    g = torch.Generator().manual_seed(0)
    xrec = (torch.randn(2, 5, 4, 2, generator=g, dtype=F64) * 3).requires_grad_(True)
    with torch.no_grad():
        xrec[0, 0, 0, 1] = 12.0        # b > 10: clamped, no gradient through b
        xrec[0, 0, 1, 1] = -9.0        # exp(b) < 1e-3: clamped, no gradient through b
    xtrue = torch.randn(2, 5, 4, 2, generator=g, dtype=F64)
    obs = torch.rand(2, 5, 4, generator=g) > 0.4
    xtrue[..., 1] = obs.to(F64)
    xtrue[..., 0] = xtrue[..., 0] * obs
    a, b = xrec[..., 0].detach(), xrec[..., 1].detach()
    w = obs.to(F64)
    n = w.sum()
    p = torch.exp(torch.clamp(b, max=10.0))
    q = torch.clamp(p, min=1e-3)
    s2 = 1 / q
    s2t = 1 / torch.clamp(xtrue[..., 1], min=1e-3)          # 1 where observed, 1000 elsewhere
    d = a * s2 - xtrue[..., 0] * s2t
    for uncertain in (False, True):
        cost, rms, nn = M.cost_terms(xrec, xtrue, uncertain)
        if uncertain:
            want = ((torch.log(s2 / s2t) + (s2t + d ** 2) / s2) * w).sum() / n
        else:
            want = ((torch.log(s2) + d ** 2 / s2) * w).sum() / n
        assert nn.item() == n.item()
        assert abs(cost.item() - want.item()) < 1e-12 * abs(want.item())
        assert abs(rms.item() - math.sqrt(((d ** 2) * w).sum().item() / n.item())) < 1e-12
        ga, = torch.autograd.grad(cost, xrec)
        live = ((p >= 1e-3) & (b <= 10.0)).to(F64)
        want_a = 2 * w * d / n
        want_b = w * (-1 / q + d ** 2 - 2 * d * a / q + (s2t if uncertain else 0.0)) / n * live * p
        assert torch.allclose(ga[..., 0], want_a, rtol=1e-11, atol=1e-14)
        assert torch.allclose(ga[..., 1], want_b, rtol=1e-11, atol=1e-14)
    assert ga[0, 0, 0, 1].item() == 0.0 and ga[0, 0, 1, 1].item() == 0.0


def test_l2_term_skips_biases():
    """beta * sum of tf.nn.l2_loss(v) = beta/2 * sum v^2 over variables without 'bias' in the
    name (:563-567): conv AND dense kernels."""
    arch = M.Arch(4, 4, 2, [3], [0.5], [1])
    params = [torch.full(s, 2.0, dtype=F64) for _, s in arch.param_shapes()]
    nk = sum(int(np.prod(s)) for n, s in arch.param_shapes() if "bias" not in n)
    assert M.l2_term(arch, params, 0.25).item() == 0.25 * 0.5 * 4.0 * nk


def test_clip_by_global_norm():
    """tf.clip_by_global_norm (:602): global norm over all tensors; unchanged below the clip."""
    g = [torch.tensor([3.0, 0.0], dtype=F64), torch.tensor([[4.0]], dtype=F64)]
    out, norm = M.clip_by_global_norm(g, 2.5)
    assert norm.item() == 5.0
    assert torch.allclose(out[0], torch.tensor([1.5, 0.0], dtype=F64)) and torch.allclose(out[1], torch.tensor([[2.0]], dtype=F64))
    out, norm = M.clip_by_global_norm(g, 10.0)
    assert torch.allclose(out[0], g[0]) and torch.allclose(out[1], g[1])


def test_tf_adam_steps_by_hand():
    """tf.train.AdamOptimizer (:588-600): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),
    m += (1-b1)(g-m), v += (1-b2)(g^2-v), theta -= lr_t * m / (sqrt(v) + eps) -- epsilon is
    NOT scaled by the bias correction (differs from torch.optim.Adam)."""
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    p = [torch.tensor([1.0, -2.0], dtype=F64)]
    opt = M.TFAdam(p)
    g1 = torch.tensor([0.5, -4.0], dtype=F64)
    opt.step(p, [g1], lr)
    m1, v1 = (1 - b1) * g1, (1 - b2) * g1 ** 2
    lr1 = lr * math.sqrt(1 - b2) / (1 - b1)
    want1 = torch.tensor([1.0, -2.0], dtype=F64) - lr1 * m1 / (torch.sqrt(v1) + eps)
    assert torch.allclose(p[0], want1, rtol=1e-14)
    # first step moves every weight by ~lr against the sign of its gradient
    assert torch.allclose(p[0], torch.tensor([1.0 - lr, -2.0 + lr], dtype=F64), atol=1e-9)
    g2 = torch.tensor([-1.0, 0.25], dtype=F64)
    opt.step(p, [g2], lr)
    m2 = m1 + (1 - b1) * (g2 - m1)
    v2 = v1 + (1 - b2) * (g2 ** 2 - v1)
    lr2 = lr * math.sqrt(1 - b2 ** 2) / (1 - b1 ** 2)
    want2 = want1 - lr2 * m2 / (torch.sqrt(v2) + eps)
    assert torch.allclose(p[0], want2, rtol=1e-13)
    # epsilon outside the bias correction: a tiny gradient shows the difference to torch.optim.Adam
    q = [torch.tensor([1.0], dtype=F64)]
    ref = torch.nn.Parameter(torch.tensor([1.0], dtype=F64))
    o2, ot = M.TFAdam(q), torch.optim.Adam([ref], lr=lr, betas=(b1, b2), eps=eps)
    gt = torch.tensor([1e-7], dtype=F64)
    o2.step(q, [gt], lr)
    ref.grad = gt.clone()
    ot.step()
    tf_step = lr1 * (1 - b1) * 1e-7 / (math.sqrt(1 - b2) * 1e-7 + eps)
    assert abs((1.0 - q[0].item()) - tf_step) < 1e-15
    assert abs((1.0 - q[0].item()) - (1.0 - ref.item())) > 1e-5 * lr


def test_train_step_matches_committed_fp64_evaluation():
    """Drift guard: tests/golden/model_small.npz holds an fp64 evaluation of forward, cost,
    gradients, clipping and three Adam steps on seeded inputs (oracle/make_golden_model.py)."""
    from oracle import make_golden_model as G
    path = os.path.join(os.path.dirname(__file__), "golden", "model_small.npz")
    z = np.load(path, allow_pickle=False)
    got = G.evaluate()
    assert sorted(got.keys()) == sorted(z.files)
    for k in z.files:
        a, b = np.asarray(got[k], dtype=np.float64), z[k].astype(np.float64)
        assert a.shape == b.shape, k
        assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max()), k
