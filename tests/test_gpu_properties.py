"""Size-independent properties of the hot path at BASELINE.json's full size (config A: batch
50, 112x112, default architecture), where the CPU oracle would take minutes: per-frame
independence, additivity of the un-normalised gradient over the batch, determinism, and the
TF32 contract of the tensor-core back end (everything a convolution reads is representable)."""
import numpy as np
import pytest
import torch

from dincae_b200 import synthetic
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Network, Plan

pytestmark = pytest.mark.gpu
T, H, W, B = 64, 112, 112, 50


@pytest.fixture(scope="module")
def setup():
    lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=W, seed=4)
    missing = np.ma.getmaskarray(data)
    cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
    plan = Plan(H, W, cube.nvar)
    params = plan.glorot_init(3)
    rs = np.random.RandomState(1)
    frames = rs.randint(0, T, size=B).astype(np.int32)
    clouds = rs.randint(0, T, size=B).astype(np.int32)
    seq = np.arange(B, dtype=np.int64) + 1000
    return cube, plan, params, frames, clouds, seq


def run(cube, plan, params, frames, clouds, seq, backward=True):
    n = len(frames)
    net = Network(plan, n, train=True)
    net.set_params_tf(params)
    net.n_noncloud.zero_()
    cube.build_batch(torch.from_numpy(frames).cuda(), torch.from_numpy(clouds).cuda(), n, net.X[0],
                     net.xtrue, net.n_noncloud, jitter_std=[0.05], seed=77,
                     noise_seq=torch.from_numpy(seq).cuda())
    net.forward(n)
    net.head_recon(n)
    if backward:
        net.loss_backward(n, False, normalise=False)
    torch.cuda.synchronize()
    return net


def test_frames_are_independent_of_their_batch(setup):
    """A frame's reconstruction does not depend on which batch (or batch size) it is in: the
    50-frame batch equals the same frames pushed through in batches of 17 -- bit for bit.
    Batches of 16 or fewer take the weight-streaming dense kernels (another summation order in
    the bottleneck): bit-identical among themselves, rounding-level agreement with the rest."""
    cube, plan, params, frames, clouds, seq = setup
    full = run(cube, plan, params, frames, clouds, seq, backward=False)
    m_full, s_full = full.m_rec.cpu().numpy(), full.s2_rec.cpu().numpy()
    for lo, hi in ((0, 17), (17, 34), (33, 50)):
        part = run(cube, plan, params, frames[lo:hi], clouds[lo:hi], seq[lo:hi], backward=False)
        assert np.array_equal(part.m_rec.cpu().numpy(), m_full[lo:hi])
        assert np.array_equal(part.s2_rec.cpu().numpy(), s_full[lo:hi])
    small = [run(cube, plan, params, frames[:n], clouds[:n], seq[:n], backward=False) for n in (5, 8, 13, 16)]
    for net in small[1:]:
        assert np.array_equal(net.m_rec.cpu().numpy()[:5], small[0].m_rec.cpu().numpy())
        assert np.array_equal(net.s2_rec.cpu().numpy()[:5], small[0].s2_rec.cpu().numpy())
    m16 = small[-1].m_rec.cpu().numpy()
    assert np.abs(m16 - m_full[:16]).max() <= 2e-3 * np.abs(m_full).max()    # TF32 conv back end
    assert np.isfinite(m_full).all() and (s_full > 0).all()


def test_unnormalised_gradient_is_additive_over_the_batch(setup):
    """grads(50 frames) == grads(first 25) + grads(last 25) for the un-normalised loss (what the
    data-parallel all-reduce relies on); only the fp32 summation order differs."""
    cube, plan, params, frames, clouds, seq = setup
    full = run(cube, plan, params, frames, clouds, seq)
    a = run(cube, plan, params, frames[:25], clouds[:25], seq[:25])
    b = run(cube, plan, params, frames[25:], clouds[25:], seq[25:])
    g = full.grads.double()
    gs = a.grads.double() + b.grads.double()
    rel = float((g - gs).norm() / g.norm())
    print('gradient additivity rel err', rel)
    assert rel < 2e-5, rel
    sums = full.loss_sums.cpu().numpy()
    want = a.loss_sums.cpu().numpy() + b.loss_sums.cpu().numpy()
    assert sums[3] == want[3]                                   # observed-pixel count: exact
    assert np.allclose(sums[:3], want[:3], rtol=1e-6)         # per-thread fp32 partial sums
    assert int(full.n_noncloud.item()) == int(sums[3])


def test_step_is_deterministic_and_conv_inputs_are_tf32(setup):
    cube, plan, params, frames, clouds, seq = setup
    n1 = run(cube, plan, params, frames, clouds, seq)
    n2 = run(cube, plan, params, frames, clouds, seq)
    assert torch.equal(n1.grads, n2.grads)
    assert torch.equal(n1.loss_sums, n2.loss_sums)
    assert torch.equal(n1.m_rec, n2.m_rec)
    if n1.lib.dincae_get_conv_backend() == 1:
        # every tensor a convolution reads has its low 13 mantissa bits clear
        # (dX[L] comes from the fp32 dense layer and is read through the rounding register
        # loader, so it is not part of the contract)
        for t in [n1.X[k] for k in range(plan.L)] + n1.D[1:plan.L] + n1.GD[1:] + \
                 [d for d in n1.dX[2:plan.L] if d is not None]:
            bits = t.view(torch.int32)
            assert int((bits & 0x1FFF).abs().max()) == 0
