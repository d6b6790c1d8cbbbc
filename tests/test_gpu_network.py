"""Whole training step and reconstruction on the GPU vs the torch-CPU oracle with identical
weights and inputs: forward activations, cost, RMS, all 20 gradients, weights after 3 Adam
steps, m_rec / sigma2_rec.  fp32 FFMA path; tolerances written at each assert."""
import numpy as np
import pytest
import torch

from dincae_b200 import layout
from dincae_b200.engine import Network, Plan
from oracle import model_torch as M
from tests.util import from_p4, rel_err, tf32_round

pytestmark = pytest.mark.gpu

CASES = {
    "tiny": dict(B=2, H=16, W=16, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "odd": dict(B=3, H=30, W=22, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "configA": dict(B=4, H=112, W=112, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4]),
    "multivar_deep": dict(B=2, H=40, W=48, nvar=22, filt=[8, 12, 18, 27, 40], frac=[0.2], skip=[1, 2, 3, 4, 5]),
    "nodense_partskip": dict(B=2, H=24, W=20, nvar=10, filt=[8, 12, 16], frac=[], skip=[2, 3]),
    "dense3": dict(B=2, H=16, W=12, nvar=6, filt=[8, 12], frac=[0.3, 0.1], skip=[1]),
}


def make(case, seed=0):
    c = CASES[case]
    arch = M.Arch(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    plan = Plan(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    g = torch.Generator().manual_seed(seed)
    params = M.init_params(arch, seed)
    # non-zero biases so that bias paths are exercised
    for (name, _), p in zip(arch.param_shapes(), params):
        if "bias" in name:
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    B, H, W, nvar = c["B"], c["H"], c["W"], c["nvar"]
    xin = torch.randn(B, H, W, nvar, generator=g)
    obs = torch.rand(B, H, W, generator=g) > 0.5
    xtrue = torch.zeros(B, H, W, 2)
    xtrue[..., 1] = obs.float()
    xtrue[..., 0] = torch.randn(B, H, W, generator=g) * obs
    return c, arch, plan, params, xin, xtrue


def load_inputs(net, xin, xtrue):
    # the batch builder hands TF32-representable inputs to the tensor-core back end
    # (dincae_build_batch); tests that bypass it follow the same contract
    if net.lib.dincae_get_conv_backend() == 1:
        xin = tf32_round(xin)
    B = xin.shape[0]
    net.X[0][:B].copy_(torch.from_numpy(layout.nhwc_to_p4(xin.numpy())))
    net.xtrue[:B].copy_(xtrue)
    net.n_noncloud.fill_(int((xtrue[..., 1] != 0).sum()))


@pytest.mark.parametrize("case", list(CASES))
def test_param_layout_matches_tf_variables(case):
    c, arch, plan, params, _, _ = make(case)
    assert [e["name"] for e in plan.entries] == [n for n, _ in arch.param_shapes()]
    assert [tuple(e["tf_shape"]) for e in plan.entries] == [tuple(s) for _, s in arch.param_shapes()]
    flat = plan.pack([p.numpy() for p in params])
    back = plan.unpack(flat)
    for a, b in zip(back, params):
        assert np.array_equal(a, b.numpy())


# tolerance multipliers of the tcgen05 kind::tf32 back end relative to the fp32 one
# (operands carry 10 mantissa bits; measured errors are printed by the tests)
TF32 = {0: 1.0, 1: 100.0}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("uncertain,beta", [(False, 0.0), (True, 0.001)])
def test_forward_loss_gradients(case, uncertain, beta, backend):
    k = TF32[backend]
    c, arch, plan, params, xin, xtrue = make(case, seed=1)
    B = c["B"]
    # float64 oracle: the arbiter when two float32 evaluations (CPU / GPU) disagree
    cost, rms, n, grads, xrec, keep = M.loss_and_grads(
        arch, [p.double() for p in params], xin.double(), xtrue.double(), uncertain, beta)
    grads32 = M.loss_and_grads(arch, params, xin, xtrue, uncertain, beta)[3]
    net = Network(plan, B, train=True)
    net.set_params_tf([p.numpy() for p in params])
    load_inputs(net, xin, xtrue)
    net.forward(B)
    net.loss_backward(B, uncertain, normalise=True)
    net.l2_value()
    torch.cuda.synchronize()
    L = plan.L
    for l in range(1, L + 1):
        got = from_p4(net.X[l][:B], plan.enc_c[l])
        e1 = rel_err(got, keep["enc_pool"][l].detach().numpy())
        assert e1 < 2e-5 * k, "enc_pool %d" % l
        got = from_p4(net.D[l][:B], plan.dec_c[l])
        e2 = rel_err(got, keep["dec"][l].detach().numpy())
        assert e2 < 5e-5 * k, "dec %d" % l
        print("layer %d: enc_pool err %.2e dec err %.2e" % (l, e1, e2))
    got_cost = net.loss_out[0].item() + beta * net.l2_out.item()
    print("cost %.7f oracle %.7f" % (got_cost, cost.item()))
    assert abs(got_cost - cost.item()) < 5e-5 * k * max(1.0, abs(cost.item()))
    assert abs(net.loss_out[1].item() - rms.item()) < 5e-5 * k * max(1.0, abs(rms.item()))
    got_grads = net.get_grads_tf()
    names = [nm for nm, _ in arch.param_shapes()]
    for nm, gg, gw, g32, p in zip(names, got_grads, grads, grads32, params):
        want = gw.numpy() - (beta * p.numpy() if "bias" not in nm else 0.0)   # L2 term joins in Adam
        scale = max(np.abs(want).max(), 1e-6)
        # fp32 kernels vs fp64 oracle.  A pre-activation that rounds to the other side of zero
        # flips a leaky-ReLU slope (1 <-> 0.2): a discrete O(1e-4..1e-3) effect on the first
        # layer's gradient that the fp32 CPU evaluation shows as well, so the bound is the
        # larger of 1e-4 and 3x the fp32-CPU error against the same fp64 arbiter.
        cpu_err = float((g32.double() - gw).abs().max()) / scale
        err = np.abs(gg - want).max() / scale
        l2 = float(np.linalg.norm((gg - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-30))
        print("grad %-16s max-norm err %.2e  L2 err %.2e" % (nm, err, l2))
        if backend == 0:
            assert err < max(1e-4, 3 * cpu_err), nm
        else:
            # TF32 operands perturb every pre-activation by ~3e-4 of its scale; the ones that
            # land on the other side of zero switch their leaky-ReLU slope, so a weight
            # gradient (a sum of +-terms over all pixels) moves by ~sqrt(flipped/total).
            # Bound the relative L2 error of each gradient tensor and its worst entry.
            assert l2 < 3e-2 and err < 6e-2, nm
    # head for reconstruction
    net.head_recon(B)
    torch.cuda.synchronize()
    m_rec, s2 = M.head(xrec)
    assert rel_err(net.m_rec[:B].cpu().numpy(), m_rec.numpy()) < 5e-5 * k
    assert rel_err(net.s2_rec[:B].cpu().numpy(), s2.numpy()) < 5e-5 * k


@pytest.mark.parametrize("case", ["tiny", "odd", "nodense_partskip"])
def test_three_adam_steps_match_oracle(case, backend):
    k = TF32[backend]
    c, arch, plan, params, xin, xtrue = make(case, seed=2)
    B = c["B"]
    beta, uncertain, lr = 0.001, False, 1e-3
    net = Network(plan, B, train=True)
    net.set_params_tf([p.numpy() for p in params])
    net.lr.fill_(lr)
    ps = [p.clone() for p in params]
    opt = M.TFAdam(ps)
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = xin + 0.1 * torch.randn(xin.shape, generator=g)
        cost, rms, n, grads, norm = M.train_step(arch, ps, opt, x, xtrue, lr, 5.0, uncertain, beta)
        load_inputs(net, x, xtrue)
        net.forward(B)
        net.loss_backward(B, uncertain, normalise=False)      # production mode: divide in Adam
        net.l2_value()
        net.adam_step(clip_norm=5.0, l2_beta=beta, denom=net.loss_sums[3:4])
        torch.cuda.synchronize()
        assert abs(net.adam_state[2].item() - norm.item()) < 2e-4 * k * norm.item()
        got_cost = net.loss_out[0].item() + beta * net.l2_out.item()
        assert abs(got_cost - cost.item()) < 1e-4 * k * max(1.0, abs(cost.item()))
    for nm, a, b in zip([n_ for n_, _ in arch.param_shapes()], net.get_params_tf(), ps):
        # Adam normalises the step: |delta| <= lr per step whatever the gradient scale.  fp32:
        # rounding-level agreement.  TF32: an entry whose gradient is at the operand-rounding
        # level can take its 3 steps in the opposite direction: |diff| <= 2 * 3 * lr
        assert np.abs(a - b.numpy()).max() < (2e-5 if backend == 0 else 6.1 * lr), nm
    # padded entries of the flat parameter vector never move away from zero
    flat = net.params.cpu().numpy()
    assert np.array_equal(plan.pack(net.get_params_tf()), flat)


def test_partial_batch_uses_prefix_of_buffers(backend):
    k = TF32[backend]
    c, arch, plan, params, xin, xtrue = make("odd", seed=3)
    net = Network(plan, 8, train=False)
    net.set_params_tf([p.numpy() for p in params])
    B = 3
    load_inputs(net, xin, xtrue)
    net.forward(B, save_signs=False)
    net.head_recon(B)
    torch.cuda.synchronize()
    m_rec, s2 = M.head(M.forward(arch, params, xin))
    assert rel_err(net.m_rec[:B].cpu().numpy(), m_rec.numpy()) < 5e-5 * k
