"""Is the bf16 network's forward pass independent of the batch it runs in?  Frames 0..3 through a
4-frame network and as the first half of an 8-frame network (same weights, same noise keys):
every stored activation of those frames must be bit-identical.  python tools/batch_invariance.py [H]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dincae_b200 import synthetic  # noqa: E402
from dincae_b200.data import DeviceCube  # noqa: E402
from dincae_b200.engine import Plan, make_network  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T, nf = 24, 4
fields, missing = [], []
for k, kind in enumerate(["sst", "chl", "wind", "wind"]):
    lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=H, seed=100 + k, kind=kind)
    fields.append(data)
    missing.append(np.ma.getmaskarray(data))
cube = DeviceCube(lon, lat, time_, fields, missing, [1.0] * nf, 3)
plan = Plan(H, H, cube.nvar, (32, 48, 72, 108, 162), (0.2,), (1, 2, 3, 4, 5), lanes=8)
frames = torch.arange(8, dtype=torch.int32, device="cuda")
clouds = torch.arange(8, 16, dtype=torch.int32, device="cuda")
seq = torch.arange(100, 108, dtype=torch.int64, device="cuda")
nets = {}
for B in (4, 8):
    net = make_network(plan, B, True, "cuda")
    net.set_params_tf(plan.glorot_init(5))
    nets[B] = net
for B, net in nets.items():
    net.n_noncloud.zero_()
    net.build_input(cube, frames[:B], clouds[:B], B, jitter_std=[0.05] * nf, seed=11, noise_seq=seq[:B])
    net.forward(B)
    net.loss_backward(B, False, normalise=False)
torch.cuda.synchronize()
a, b = nets[4], nets[8]


def cmp(name, x, y):
    x, y = x[:4].float(), y[:4].float()
    same = torch.equal(x, y)
    print("%-28s %s  max|d| %.3e  max|x| %.3e" % (name, "identical" if same else "DIFFERENT", float((x - y).abs().max()),
                                                 float(x.abs().max())))


cmp("input X8[0]", a.X8[0], b.X8[0])
for l in range(1, plan.L + 1):
    cmp("enc %d X8" % l, a.X8[l], b.X8[l])
for name in ("Fx", "dZf"):
    if hasattr(a, name):
        pass
for l in range(len(a.D8)):
    if a.D8[l] is not None:
        cmp("dec %d D8" % l, a.D8[l], b.D8[l])
cmp("xrec", a.xrec, b.xrec)
s_lo = a.loss_sums.clone()
a.loss_sums.zero_()
a.n_noncloud.zero_()
a.build_input(cube, frames[4:], clouds[4:], 4, jitter_std=[0.05] * nf, seed=11, noise_seq=seq[4:])
a.forward(4)
a.loss_backward(4, False, normalise=False)
torch.cuda.synchronize()
cmp("xrec (frames 4-7)", a.xrec, b.xrec[4:])
tot = s_lo + a.loss_sums
print("loss sums 4+4", tot.tolist())
print("loss sums 8  ", b.loss_sums.tolist())
print("rel diff", ((tot - b.loss_sums).abs() / b.loss_sums.abs()).tolist())
