"""PyTorch-CPU restatement of the TF-1.15 graph of ``DINCAE.reconstruct``.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- never imported by ``dincae_b200``.

PARITY UNPINNED at tensor level: the arithmetic of DINCAE/__init__.py:433-603 is executed
by ``tensorflow==1.15.4`` (setup.py:19), which is not vendored in the reference tree and
cannot be installed here; the reference's tests only bound a scalar loss on a downloaded
file (test/test_DINCAE_SST.py:78,107,131).  This file restates the graph with the published
TF-1.15 op semantics:

  encoder   DINCAE/__init__.py:433-456   conv3x3 SAME + leaky_relu(0.2), avg-pool 2x2/2 SAME
  dense     :458-485                     flatten (NHWC order) -> relu(xW+b) x2; dropout inert (Q2)
  decoder   :487-518                     legacy nearest resize, concat [upsampled, skip], conv
  head      :520-529, sinv :298-299
  loss      :532-570
  optimizer :586-603                     TF-Adam on gradients clipped by global norm

Tensors are NHWC, conv kernels HWIO ``[3,3,Cin,Cout]``, dense kernels ``[in,out]``, exactly
as TF stores them, so weights can be exchanged with ``dincae_b200`` without reinterpretation.
"""
import math
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn.functional as F


@dataclass
class Arch:
    """Static description of the network (keyword arguments of ``reconstruct``, :302-329)."""
    H: int
    W: int
    nvar: int = 10
    enc_nfilter_internal: List[int] = field(default_factory=lambda: [16, 24, 36, 54])
    frac_dense_layer: List[float] = field(default_factory=lambda: [0.2])
    skipconnections: List[int] = field(default_factory=lambda: [1, 2, 3, 4])

    @property
    def enc_nfilter(self):
        return [self.nvar] + list(self.enc_nfilter_internal)

    @property
    def dec_nfilter(self):
        """Indexed by decoder layer l = 0..L like the reference (:385); entry 0 is the
        channel count of dec_conv[0] (the bottleneck output), conv l emits entry l."""
        return list(self.enc_nfilter_internal[::-1]) + [2]

    @property
    def nlayers(self):
        return len(self.enc_nfilter_internal)

    def sizes(self):
        """Spatial size after k pools, k = 0..L (SAME pooling: ceil)."""
        out = [(self.H, self.W)]
        for _ in range(self.nlayers):
            h, w = out[-1]
            out.append(((h + 1) // 2, (w + 1) // 2))
        return out

    def dense_units(self):
        if not self.frac_dense_layer:
            return []
        h, w = self.sizes()[-1]
        n = h * w * self.enc_nfilter[-1]
        fr = list(self.frac_dense_layer) + list(reversed(self.frac_dense_layer[:-1]))
        return [math.floor(n * f) for f in fr] + [n]

    def dec_cin(self, l):
        """Input channels of decoder conv l (1-based): upsampled [+ skip] (:503-506)."""
        L = self.nlayers
        up = self.dec_nfilter[l - 1]      # dec_nfilter[0] = channels of the bottleneck output
        skip = self.enc_nfilter[L - l] if l in self.skipconnections else 0
        return up, skip

    def param_shapes(self):
        """(name, shape) in TF variable-creation order."""
        out = []
        k = 0

        def conv(cin, cout):
            nonlocal k
            name = "conv2d" if k == 0 else "conv2d_%d" % k
            k += 1
            out.append((name + "/kernel", (3, 3, cin, cout)))
            out.append((name + "/bias", (cout,)))

        enc = self.enc_nfilter
        for l in range(1, len(enc)):
            conv(enc[l - 1], enc[l])
        units = self.dense_units()
        if units:
            nin = units[-1]
            for i, u in enumerate(units):
                name = "dense" if i == 0 else "dense_%d" % i
                out.append((name + "/kernel", (nin, u)))
                out.append((name + "/bias", (u,)))
                nin = u
        for l in range(1, self.nlayers + 1):
            up, skip = self.dec_cin(l)
            conv(up + skip, self.dec_nfilter[l])
        # TF creates encoder convs, dense, decoder convs in that order; the conv counter
        # runs across encoder and decoder (conv2d_4.. for the decoder at L=4).
        return out


def init_params(arch, seed=0, dtype=torch.float32):
    """Glorot-uniform kernels, zero biases (tf.layers defaults)."""
    g = torch.Generator().manual_seed(seed)
    params = []
    for name, shape in arch.param_shapes():
        if name.endswith("bias"):
            p = torch.zeros(shape, dtype=dtype)
        else:
            if len(shape) == 4:
                fan_in, fan_out = 9 * shape[2], 9 * shape[3]
            else:
                fan_in, fan_out = shape
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            p = ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)
        params.append(p)
    return params


def _conv(x_nhwc, k_hwio, b):
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), k_hwio.permute(3, 2, 0, 1), b, padding=1)
    return y.permute(0, 2, 3, 1)


def _pool(x_nhwc):
    y = F.avg_pool2d(x_nhwc.permute(0, 3, 1, 2), 2, 2, ceil_mode=True, count_include_pad=False)
    return y.permute(0, 2, 3, 1)


def _resize_nearest(x_nhwc, size):
    y = F.interpolate(x_nhwc.permute(0, 3, 1, 2), size=size, mode="nearest")
    return y.permute(0, 2, 3, 1)


def forward(arch, params, xin, keep=None):
    """Network forward: ``xin[B,H,W,nvar]`` -> ``xrec[B,H,W,2]`` (:433-518).

    ``keep`` (dict) receives every intermediate activation (NHWC).
    """
    L = arch.nlayers
    it = iter(params)
    enc_pool = [xin]
    enc_conv = [None]
    for l in range(1, L + 1):
        k, b = next(it), next(it)
        c = F.leaky_relu(_conv(enc_pool[l - 1], k, b), 0.2)
        enc_conv.append(c)
        enc_pool.append(_pool(c))
    x = enc_pool[L]
    dense_act = []
    if arch.frac_dense_layer:
        shp = x.shape
        x = x.reshape(shp[0], -1)
        for _ in arch.dense_units():
            k, b = next(it), next(it)
            x = F.relu(x @ k + b)      # dropout: identity (tf.layers.dropout, training=False)
            dense_act.append(x)
        x = x.reshape(shp)
    dec = [x]
    for l in range(1, L + 1):
        l2 = L + 1 - l
        up = _resize_nearest(dec[l - 1], enc_conv[l2].shape[1:3])
        if l in arch.skipconnections:
            up = torch.cat([up, enc_pool[l2 - 1]], dim=3)
        k, b = next(it), next(it)
        dec.append(F.leaky_relu(_conv(up, k, b), 0.2))
    if keep is not None:
        keep.update(enc_conv=enc_conv, enc_pool=enc_pool, dense=dense_act, dec=dec)
    return dec[L]


def head(xrec):
    """(m_rec, sigma2_rec) from the two output channels (:520-523)."""
    inv = torch.exp(torch.clamp(xrec[..., 1], max=10.))
    s2 = 1 / torch.clamp(inv, min=1e-3)
    return xrec[..., 0] * s2, s2


def cost_terms(xrec, xtrue, truth_uncertain=False):
    """cost (without L2), RMS, n_noncloud (:526-570)."""
    m_rec, s2_rec = head(xrec)
    s2_true = 1 / torch.clamp(xtrue[..., 1], min=1e-3)
    m_true = xtrue[..., 0] * s2_true
    d = m_rec - m_true
    w = (xtrue[..., 1] != 0).to(xrec.dtype)
    n = w.sum()
    if truth_uncertain:
        cost = ((torch.log(s2_rec / s2_true) + (s2_true + d ** 2) / s2_rec) * w).sum() / n
    else:
        cost = ((torch.log(s2_rec) * w).sum() + ((d ** 2 / s2_rec) * w).sum()) / n
    rms = torch.sqrt(((d ** 2) * w).sum() / n)
    return cost, rms, n


def l2_term(arch, params, beta):
    """beta * sum of tf.nn.l2_loss over the non-bias variables (:563-567)."""
    tot = 0.
    for (name, _), p in zip(arch.param_shapes(), params):
        if "bias" not in name:
            tot = tot + (p ** 2).sum() / 2
    return beta * tot


def loss_and_grads(arch, params, xin, xtrue, truth_uncertain=False, beta=0.):
    """One evaluation of ``optimizer.compute_gradients(cost)`` (:601)."""
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    keep = {}
    xrec = forward(arch, ps, xin, keep)
    cost, rms, n = cost_terms(xrec, xtrue, truth_uncertain)
    if beta != 0:
        cost = cost + l2_term(arch, ps, beta)
    grads = torch.autograd.grad(cost, ps)
    return cost.detach(), rms.detach(), n.detach(), [g.detach() for g in grads], xrec.detach(), keep


def clip_by_global_norm(grads, clip):
    """tf.clip_by_global_norm (:602): t * clip * min(1/norm, 1/clip)."""
    norm = torch.sqrt(sum((g ** 2).sum() for g in grads))
    scale = clip * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip, dtype=norm.dtype))
    return [g * scale for g in grads], norm


class TFAdam:
    """tf.train.AdamOptimizer(lr, 0.9, 0.999, 1e-8) (:588-600); epsilon OUTSIDE the
    bias correction, beta powers kept as running products."""

    def __init__(self, params, beta1=0.9, beta2=0.999, eps=1e-8):
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        dt = params[0].dtype
        self.b1p = torch.tensor(beta1, dtype=dt)
        self.b2p = torch.tensor(beta2, dtype=dt)

    def step(self, params, grads, lr):
        dt = params[0].dtype
        lr = torch.tensor(lr, dtype=dt)
        alpha = lr * torch.sqrt(1 - self.b2p) / (1 - self.b1p)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m += (g - m) * (1 - torch.tensor(self.b1, dtype=dt))     # T(1) - beta1() in dtype
            v += (g * g - v) * (1 - torch.tensor(self.b2, dtype=dt))
            p -= (m * alpha) / (torch.sqrt(v) + self.eps)
        self.b1p = self.b1p * self.b1
        self.b2p = self.b2p * self.b2


def train_step(arch, params, opt, xin, xtrue, lr, clip_grad=5.0, truth_uncertain=False, beta=0.):
    """One ``sess.run([cost, RMS, opt])`` (:649-654).  Mutates ``params`` and ``opt``."""
    cost, rms, n, grads, xrec, _ = loss_and_grads(arch, params, xin, xtrue, truth_uncertain, beta)
    clipped, norm = clip_by_global_norm(grads, clip_grad)
    opt.step(params, clipped, lr)
    return cost, rms, n, grads, norm
