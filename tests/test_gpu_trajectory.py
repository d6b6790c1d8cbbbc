"""Multi-step parity on BASELINE configs[0] (test_DINCAE_SST-style run: synthetic 112x112 cube,
200 frames, batch 50, 2 epochs = 8 optimisation steps, then the reconstruction) against the
oracle with identical weights, frame indices, added-cloud indices and jitter draws
(DINCAE/__init__.py:196-227 generator, :433-603 graph, :637-689 loops), and a backward pass run
from the ORACLE's forward state (sign masks, activations) that isolates the arithmetic of the
backward kernels from leaky-ReLU sign flips of the forward pass."""
import numpy as np
import pytest
import torch

from dincae_b200 import layout, synthetic
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Plan, make_network
from oracle import datagen_np
from oracle import model_torch as M
from tests.util import from_p4, rel_err, tf32_round

pytestmark = pytest.mark.gpu

# per-step relative loss tolerance / reconstruction tolerance (of the field's maximum) per arithmetic
# fp32 kernels follow the oracle to rounding level.  TF32 / bf16 operands perturb every gradient at
# the 5e-4 / 4e-3 level and Adam normalises step sizes, so the weight trajectories separate by
# O(lr) per step: the per-step loss then agrees to about 1e-2 over the 8 steps (measured worst
# case: TF32 5.7e-3 at a loss spike of step 5, bf16 3.5e-3; printed by the test).
# Measured (B200): TF32 per-step errors 1.6e-5 .. 1.1e-3 plus 5.7e-3 at the loss spike of step 5,
# reconstruction 2.3e-2 of the field's maximum after the 8 steps; fp32 1.1e-6 / 1.3e-4.
TOL = {"fp32": (1e-4, 1e-3), "tf32": (1e-2, 4e-2), "bf16": (2e-2, 5e-2)}


@pytest.mark.parametrize("mode", ["fp32", "tf32", "bf16"])
def test_configs0_trajectory_matches_oracle(mode):
    from dincae_b200 import _lib
    lib = _lib.load()
    lib.dincae_set_conv_backend(0 if mode == "fp32" else 1)
    try:
        T, H, W, B, epochs, jit = 200, 112, 112, 50, 2, 0.05
        lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=W, seed=77)
        missing = np.ma.getmaskarray(data)
        feat = datagen_np.StaticFeatures(lon, lat, time_, [data], [1.0])
        cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
        arch = M.Arch(H, W, cube.nvar)
        plan = Plan(H, W, cube.nvar, lanes=8 if mode == "bf16" else 4)
        params = M.init_params(arch, 0)
        ps = [p.clone() for p in params]
        opt = M.TFAdam(ps)
        net = make_network(plan, B, train=True)
        net.set_params_tf([p.numpy() for p in params])
        net.lr.fill_(1e-3)
        rs = np.random.RandomState(5)
        steps = epochs * ((T + B - 1) // B)
        tol_loss, tol_rec = TOL[mode]
        worst = 0.0
        for step in range(steps):
            fi = rs.randint(0, T, size=B).astype(np.int32)
            ci = rs.randint(0, T, size=B).astype(np.int32)
            noise = rs.standard_normal((B, 3, H, W))
            xs, ts = [], []
            for b in range(B):
                xin = datagen_np.stack_frame(feat, int(fi[b]), 3)
                xt = xin[..., 0:2].copy()
                datagen_np.perturb_frame(xin, [missing], int(ci[b]), list(noise[b]), [jit])
                xs.append(xin)
                ts.append(xt)
            xin_t, xtrue_t = torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ts))
            cost, rms, n, grads, norm = M.train_step(arch, ps, opt, xin_t, xtrue_t, 1e-3, 5.0, False, 0.)
            net.n_noncloud.zero_()
            cube.build_batch(torch.from_numpy(fi).cuda(), torch.from_numpy(ci).cuda(), B, net.X[0], net.xtrue,
                             net.n_noncloud, jitter_std=[jit], noise=torch.from_numpy(noise).cuda())
            net.forward(B)
            net.loss_backward(B, False, normalise=False)
            net.adam_step(clip_norm=5.0, denom=net.loss_sums[3:4], B=B)
            torch.cuda.synchronize()
            s = net.loss_sums.cpu().numpy()
            got = (s[0] + s[1]) / s[3]
            assert int(s[3]) == int(n.item())                       # observed-pixel count: exact
            err = abs(got - cost.item()) / abs(cost.item())
            worst = max(worst, err)
            print("step %d cost %.6f oracle %.6f rel %.2e" % (step, got, cost.item(), err))
            assert err < tol_loss, (step, got, cost.item())
        # reconstruction of every frame (test mode: no added clouds, no jitter)
        worst_m = worst_s = 0.0
        for s0 in range(0, T, B):
            idx = np.arange(s0, min(T, s0 + B), dtype=np.int32)
            nb = len(idx)
            x = torch.from_numpy(np.stack([datagen_np.stack_frame(feat, int(i), 3) for i in idx]))
            with torch.no_grad():
                m_ref, s2_ref = M.head(M.forward(arch, ps, x))
            net.n_noncloud.zero_()
            cube.build_batch(torch.from_numpy(idx).cuda(), None, nb, net.X[0], net.xtrue, net.n_noncloud)
            net.forward(nb, save_signs=False)
            net.head_recon(nb)
            torch.cuda.synchronize()
            worst_m = max(worst_m, rel_err(net.m_rec[:nb].cpu().numpy(), m_ref.numpy()))
            worst_s = max(worst_s, rel_err(np.sqrt(net.s2_rec[:nb].cpu().numpy()), np.sqrt(s2_ref.numpy())))
        print("mode %s: worst loss err %.2e, m_rec err %.2e, sigma_rec err %.2e" % (mode, worst, worst_m, worst_s))
        assert worst_m < tol_rec and worst_s < tol_rec
    finally:
        lib.dincae_set_conv_backend(1)


def test_backward_from_oracle_forward_state_tf32():
    """The TF32 network test accepts 3e-2 relative L2 error per gradient tensor and explains it with
    leaky-ReLU sign flips of pre-activations that TF32 operand rounding pushes across zero.  Here
    the backward pass runs from the ORACLE's forward state (pooled activations, sign masks, dense
    activations, decoder activations, xrec -- rounded to TF32 where a convolution reads them), so
    no sign can flip: every gradient then agrees with the fp64 oracle to 2e-3 relative L2."""
    c = dict(B=3, H=30, W=22, nvar=10, filt=[16, 24, 36, 54], frac=[0.2], skip=[1, 2, 3, 4])
    arch = M.Arch(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    plan = Plan(c["H"], c["W"], c["nvar"], c["filt"], c["frac"], c["skip"])
    g = torch.Generator().manual_seed(1)
    params = M.init_params(arch, 1)
    for (name, _), p in zip(arch.param_shapes(), params):
        if "bias" in name:
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    B, H, W, nvar = c["B"], c["H"], c["W"], c["nvar"]
    xin = tf32_round(torch.randn(B, H, W, nvar, generator=g))
    obs = torch.rand(B, H, W, generator=g) > 0.5
    xtrue = torch.zeros(B, H, W, 2)
    xtrue[..., 1] = obs.float()
    xtrue[..., 0] = torch.randn(B, H, W, generator=g) * obs
    p64 = [p.double() for p in params]
    cost, rms, n, grads, xrec, keep = M.loss_and_grads(arch, p64, xin.double(), xtrue.double(), False, 0.)
    net = make_network(plan, B, train=True)
    net.set_params_tf([p.numpy() for p in params])
    L = plan.L

    def put(dst, nhwc):
        dst[:B].copy_(torch.from_numpy(layout.nhwc_to_p4(tf32_round(nhwc.detach().float()).numpy())))

    put(net.X[0], xin)
    net.xtrue[:B].copy_(xtrue)
    net.n_noncloud.fill_(int((xtrue[..., 1] != 0).sum()))
    for l in range(1, L + 1):
        put(net.X[l], keep["enc_pool"][l])
        net.S[l][:B].copy_(torch.from_numpy(layout.sign_mask(keep["enc_conv"][l].detach().numpy()).view(np.int16)))
    # dense activations: F[0] aliases X[L]; middle layers are plain [B, units]; the last one aliases D[0]
    for i, a in enumerate(keep["dense"][:-1]):
        net.F[i + 1][:B].zero_()
        net.F[i + 1][:B, :a.shape[1]].copy_(a.detach().float())
    for l in range(0, L + 1):
        put(net.D[l], keep["dec"][l])
    net.loss_backward(B, False, normalise=True)
    torch.cuda.synchronize()
    names = [nm for nm, _ in arch.param_shapes()]
    worst = 0.0
    for nm, gg, gw in zip(names, net.get_grads_tf(), grads):
        want = gw.numpy()
        l2 = float(np.linalg.norm((gg - want).ravel()) / max(np.linalg.norm(want.ravel()), 1e-30))
        print("grad %-16s L2 err %.2e" % (nm, l2))
        worst = max(worst, l2)
    print("worst %.2e" % worst)
    assert worst < 2e-3
