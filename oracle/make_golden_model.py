"""Write tests/golden/model_small.npz: an fp64 evaluation of the model oracle on seeded inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  TensorFlow 1.15 cannot run here, so this is
not a reference output: it freezes the restatement (oracle/model_torch.py, whose ops are
pinned by the hand-worked answers in tests/test_oracle_model.py) so that later edits cannot
drift unnoticed, and gives GPU parity work a fixed target:

    python -m oracle.make_golden_model

Contents: inputs (xin, xtrue, the 12 variables), every intermediate activation, xrec, m_rec,
sigma2_rec, cost / RMS / n (both loss variants, with and without L2), the 12 gradients, the
global norm, and the variables after 1 and 3 clipped TF-Adam steps.
"""
import os

import numpy as np
import torch

from . import model_torch as M

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "model_small.npz")


def evaluate():
    dt = torch.float64
    arch = M.Arch(9, 14, 10, [4, 6], [0.2], [1, 2])          # odd sizes: SAME pool / resize edges
    g = torch.Generator().manual_seed(2024)
    B = 3
    xin = torch.randn(B, arch.H, arch.W, arch.nvar, generator=g, dtype=dt)
    obs = torch.rand(B, arch.H, arch.W, generator=g) > 0.5
    xtrue = torch.zeros(B, arch.H, arch.W, 2, dtype=dt)
    xtrue[..., 0] = torch.randn(B, arch.H, arch.W, generator=g, dtype=dt) * obs
    xtrue[..., 1] = obs.to(dt)
    params = M.init_params(arch, 7, dt)
    for p in params[1::2]:                                    # non-zero biases
        p += 0.1 * torch.randn(p.shape, generator=g, dtype=dt)
    out = {"xin": xin, "xtrue": xtrue}
    for i, p in enumerate(params):
        out["param_%02d" % i] = p.clone()
    cost, rms, n, grads, xrec, keep = M.loss_and_grads(arch, params, xin, xtrue)
    for l in range(1, arch.nlayers + 1):
        out["enc_conv_%d" % l] = keep["enc_conv"][l].detach()
        out["enc_pool_%d" % l] = keep["enc_pool"][l].detach()
        out["dec_%d" % l] = keep["dec"][l].detach()
    for i, d in enumerate(keep["dense"]):
        out["dense_%d" % i] = d.detach()
    m_rec, s2 = M.head(xrec)
    out.update(xrec=xrec, m_rec=m_rec, sigma2_rec=s2, cost=cost, rms=rms, n=n)
    for i, gr in enumerate(grads):
        out["grad_%02d" % i] = gr
    c2, r2, _, g2, _, _ = M.loss_and_grads(arch, params, xin, xtrue, truth_uncertain=True, beta=1e-3)
    out.update(cost_uncertain_l2=c2, rms_uncertain_l2=r2, grad_uncertain_l2_00=g2[0], grad_uncertain_l2_05=g2[5])
    ps = [p.clone() for p in params]
    opt = M.TFAdam(ps)
    for step in (1, 2, 3):
        _, _, _, _, norm = M.train_step(arch, ps, opt, xin, xtrue, 1e-3, clip_grad=0.5)
        if step == 1:
            out["global_norm"] = norm
        if step in (1, 3):
            for i in (0, 4, 5, 8, 11):
                out["param_%02d_after_%d" % (i, step)] = ps[i].clone()
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()}


def main():
    np.savez_compressed(OUT, **evaluate())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
