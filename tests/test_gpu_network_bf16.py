"""Whole training step on the bf16 path (Plan(lanes=8) / NetworkBF16: tcgen05 kind::f16
convolutions, fp32 master weights / dense / loss / Adam) vs the float64 torch-CPU oracle with
identical weights and inputs (DINCAE/__init__.py:433-603).

Stated bf16 tolerances (8 mantissa bits: every stored activation / activation gradient and
every conv operand carries 2^-9 relative rounding; about ten layers deep):
  activations, xrec, m_rec: 3e-2 of the tensor's maximum;  cost / RMS: 3e-2 relative;
  parameter gradients: relative L2 error < 0.12 per tensor (leaky-ReLU sign flips of
  pre-activations at the rounding level included, see tests/test_gpu_network.py);
  after 3 Adam steps every weight within 2 * 3 * lr of the oracle's (Adam normalises steps).
Masks, indices and the observed-pixel count stay exact."""
import numpy as np
import pytest
import torch

from dincae_b200 import layout
from dincae_b200.engine import NetworkBF16, Plan, make_network
from oracle import model_torch as M
from tests.util import rel_err

pytestmark = pytest.mark.gpu

CASES = {
    "tiny": dict(B=2, H=16, W=16, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "odd": dict(B=3, H=30, W=22, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "configA": dict(B=4, H=112, W=112, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "multivar_deep": dict(B=2, H=40, W=48, nvar=22, filt=[8, 12, 18, 27, 40], frac=[0.2], skip=[1, 2, 3, 4, 5]),
    "nodense_partskip": dict(B=2, H=24, W=20, nvar=10, filt=[8, 12, 16], frac=[], skip=[2, 3]),
    "dense3": dict(B=2, H=16, W=12, nvar=6, filt=[8, 12], frac=[0.3, 0.1], skip=[1]),
    # channel widths of BASELINE configs[3] at a small size: 270 -> 108 input channels of the first
    # decoder layer exceed one MMA (split data gradient, three M-tiles in the weight gradient)
    "configD_widths": dict(B=1, H=64, W=96, nvar=28, filt=[32, 48, 72, 108, 162], frac=[0.2], skip=[1, 2, 3, 4, 5]),
}


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def from_p8(t, C):
    return layout.p8_to_nhwc(t.detach().float().cpu().numpy(), C)


def make(case, seed=0):
    c = CASES[case]
    arch = M.Arch(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    plan = Plan(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"], lanes=8)
    g = torch.Generator().manual_seed(seed)
    params = M.init_params(arch, seed)
    for (name, _), p in zip(arch.param_shapes(), params):
        if "bias" in name:
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    B, H, W, nvar = c["B"], c["H"], c["W"], c["nvar"]
    xin = bf(torch.randn(B, H, W, nvar, generator=g))
    obs = torch.rand(B, H, W, generator=g) > 0.5
    xtrue = torch.zeros(B, H, W, 2)
    xtrue[..., 1] = obs.float()
    xtrue[..., 0] = torch.randn(B, H, W, generator=g) * obs
    return c, arch, plan, params, xin, xtrue


def load_inputs(net, xin, xtrue):
    B = xin.shape[0]
    net.X[0][:B].copy_(torch.from_numpy(layout.nhwc_to_p4(xin.numpy())))
    net.xtrue[:B].copy_(xtrue)
    net.n_noncloud.fill_(int((xtrue[..., 1] != 0).sum()))


@pytest.mark.parametrize("case", list(CASES))
def test_param_layout_round_trip(case):
    c, arch, plan, params, _, _ = make(case)
    assert [e["name"] for e in plan.entries] == [n for n, _ in arch.param_shapes()]
    flat = plan.pack([p.numpy() for p in params])
    for a, b in zip(plan.unpack(flat), params):
        assert np.array_equal(a, b.numpy())
    assert all(p % 2 == 0 for p in plan.ep + plan.dp)


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("uncertain,beta", [(False, 0.0), (True, 0.001)])
def test_forward_loss_gradients_bf16(case, uncertain, beta):
    c, arch, plan, params, xin, xtrue = make(case, seed=1)
    B = c["B"]
    cost, rms, n, grads, xrec, keep = M.loss_and_grads(
        arch, [p.double() for p in params], xin.double(), xtrue.double(), uncertain, beta)
    net = make_network(plan, B, train=True)
    assert isinstance(net, NetworkBF16)
    net.set_params_tf([p.numpy() for p in params])
    load_inputs(net, xin, xtrue)
    net.forward(B)
    net.loss_backward(B, uncertain, normalise=True)
    net.l2_value()
    torch.cuda.synchronize()
    L = plan.L
    for l in range(1, L + 1):
        e1 = rel_err(from_p8(net.X8[l][:B], plan.enc_c[l]), keep["enc_pool"][l].detach().numpy())
        assert e1 < 3e-2, "enc_pool %d" % l
        if l < L:
            e2 = rel_err(from_p8(net.D8[l][:B], plan.dec_c[l]), keep["dec"][l].detach().numpy())
        else:
            e2 = rel_err(layout.p4_to_nhwc(net.xrec[:B].cpu().numpy(), 2), keep["dec"][l].detach().numpy())
        assert e2 < 3e-2, "dec %d" % l
        print("layer %d: enc_pool err %.2e dec err %.2e" % (l, e1, e2))
    got_cost = net.loss_out[0].item() + beta * net.l2_out.item()
    print("cost %.7f oracle %.7f" % (got_cost, cost.item()))
    assert abs(got_cost - cost.item()) < 3e-2 * max(1.0, abs(cost.item()))
    assert abs(net.loss_out[1].item() - rms.item()) < 3e-2 * max(1.0, abs(rms.item()))
    got_grads = net.get_grads_tf()
    names = [nm for nm, _ in arch.param_shapes()]
    worst = 0.0
    for nm, gg, gw, p in zip(names, got_grads, grads, params):
        want = gw.numpy() - (beta * p.numpy() if "bias" not in nm else 0.0)
        l2 = float(np.linalg.norm((gg - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-30))
        print("grad %-16s L2 err %.2e" % (nm, l2))
        worst = max(worst, l2)
        assert l2 < 0.12, nm
    print("worst gradient L2 error %.3e" % worst)
    net.head_recon(B)
    torch.cuda.synchronize()
    m_rec, s2 = M.head(xrec)
    assert rel_err(net.m_rec[:B].cpu().numpy(), m_rec.numpy()) < 3e-2
    assert rel_err(net.s2_rec[:B].cpu().numpy(), s2.numpy()) < 3e-2


@pytest.mark.parametrize("case", ["tiny", "odd", "nodense_partskip"])
def test_three_adam_steps_bf16(case):
    c, arch, plan, params, xin, xtrue = make(case, seed=2)
    B = c["B"]
    beta, uncertain, lr = 0.001, False, 1e-3
    net = make_network(plan, B, train=True)
    net.set_params_tf([p.numpy() for p in params])
    net.lr.fill_(lr)
    ps = [p.clone() for p in params]
    opt = M.TFAdam(ps)
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = bf(xin + 0.1 * torch.randn(xin.shape, generator=g))
        cost, rms, n, grads, norm = M.train_step(arch, ps, opt, x, xtrue, lr, 5.0, uncertain, beta)
        load_inputs(net, x, xtrue)
        net.forward(B)
        net.loss_backward(B, uncertain, normalise=False)
        net.l2_value()
        net.adam_step(clip_norm=5.0, l2_beta=beta, denom=net.loss_sums[3:4])
        torch.cuda.synchronize()
        assert abs(net.adam_state[2].item() - norm.item()) < 5e-2 * norm.item()
        got_cost = net.loss_out[0].item() + beta * net.l2_out.item()
        assert abs(got_cost - cost.item()) < 3e-2 * max(1.0, abs(cost.item()))
    for nm, a, b in zip([n_ for n_, _ in arch.param_shapes()], net.get_params_tf(), ps):
        assert np.abs(a - b.numpy()).max() < 6.1 * lr, nm
    flat = net.params.cpu().numpy()
    assert np.array_equal(plan.pack(net.get_params_tf()), flat)     # padding never moves


def test_bf16_step_is_deterministic_and_partial_batches_work():
    c, arch, plan, params, xin, xtrue = make("odd", seed=3)
    net = make_network(plan, 8, train=True)
    net.set_params_tf([p.numpy() for p in params])
    B = 3
    outs = []
    for _ in range(2):
        load_inputs(net, xin, xtrue)
        net.forward(B)
        net.loss_backward(B, False, normalise=True)
        torch.cuda.synchronize()
        outs.append((net.grads.clone(), net.xrec[:B].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    net.head_recon(B)
    torch.cuda.synchronize()
    m_rec, s2 = M.head(M.forward(arch, params, xin))
    assert rel_err(net.m_rec[:B].cpu().numpy(), m_rec.numpy()) < 3e-2


@pytest.mark.parametrize("factored", ["1", "0"])
def test_graph_captured_bf16_trainer_equals_eager(factored, monkeypatch):
    """The warm-up step run before graph capture is undone completely -- including the tiled bf16
    copies of the dense kernels that the factored Adam kernel rewrites: losses and parameters of a
    captured trainer equal the eager trainer's bit for bit from the first step on."""
    import random
    from dincae_b200.data import DeviceCube
    from dincae_b200.sampler import TrainSampler
    from dincae_b200.trainer import Trainer
    from tests.util import random_cube
    monkeypatch.setenv("DINCAE_FACTORED", factored)
    rs = np.random.RandomState(21)
    T, H, W, B = 12, 32, 32, 4
    lon, lat, time_, fields = random_cube(rs, T, H, W, 2)
    missing = [np.ma.getmaskarray(f) for f in fields]
    cube = DeviceCube(lon, lat, time_, fields, missing, [1.0, 1.3], 3)
    plan = Plan(H, W, cube.nvar, (16, 24, 40), (0.2,), (1, 2, 3), lanes=8)
    runs = []
    for use_graph in (False, True):
        tr = Trainer(plan, cube, B, noise_seed=3, jitter_std=[0.05, 0.02], use_graph=use_graph, init_seed=4)
        tr.set_lr(1e-3)
        random.seed(5)
        sm = TrainSampler(T, B, 4, seed=9)
        for _ in range(3):
            tr.train_step(*sm.next_batch())
        runs.append((tr.flush_losses(), tr.net.params.clone(), [w.clone() for w in tr.net.w8]))
        assert bool(tr.net.factored) == (factored == "1")
    assert runs[0][0] == runs[1][0]
    assert torch.equal(runs[0][1], runs[1][1])
    for a, b in zip(runs[0][2], runs[1][2]):
        assert torch.equal(a, b)
