"""Compatibility path of ``reconstruct`` for ARBITRARY user generators.

The reference accepts any Python generator function yielding ``(xin, xtrue)`` frames
(DINCAE/__init__.py:399-410).  When the generators were not made by this package's
``data_generator*`` (so there is no device cube to build batches from), frames are pulled
on the host exactly like tf.data does -- repeat -> shuffle buffer -> batch -- uploaded, laid
out as planar-4 by a torch copy, and the same CUDA kernels run the step.  Every
arithmetic op of the network still runs in the sm_100a kernels; this path only feeds them.
"""
import os
from math import ceil

import numpy as np
import torch

from .engine import Network, make_network


def _repeat(gen_fn):
    while True:
        n = 0
        for item in gen_fn():
            n += 1
            yield item
        if n == 0:
            return


def _upload(net, frames_x, frames_t):
    n = len(frames_x)
    P = net.plan
    x = torch.from_numpy(np.stack(frames_x).astype(np.float32, copy=False)).to(net.dev)
    t = torch.from_numpy(np.stack(frames_t).astype(np.float32, copy=False)).to(net.dev)
    C = x.shape[-1]
    planes = net.X[0].shape[1]
    pad = planes * 4 - C
    if pad:
        x = torch.nn.functional.pad(x, (0, pad))
    if P.lanes == 4 and net.lib.dincae_get_conv_backend() == 1:
        # TF32 contract of the tensor-core back end (include/dincae_b200.h): tensors a convolution
        # fetches by TMA must be TF32-representable, rounded to nearest like dincae_build_batch does
        # (tcgen05 would otherwise truncate the low 13 mantissa bits)
        x = ((x.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    net.X[0][:n].copy_(x.view(n, P.H, P.W, planes, 4).permute(0, 3, 1, 2, 4))
    if getattr(net, "input_p8", False):
        net.input_p8 = False                      # the bf16 network converts X[0] itself
    net.xtrue[:n].copy_(t)
    return n


class _Holder:
    """What save_checkpoint needs of a trainer."""

    def __init__(self, net):
        self.net, self.step_count = net, 0


def reconstruct_from_generators(plan, lon, lat, mask, meandata, train_datagen, train_len,
                                test_datagen, test_len, outdir, epochs, batch_size, save_each,
                                truth_uncertain, shuffle_buffer_size, clip_grad,
                                regularization_L2_beta, transfun, savesample, learning_rate,
                                learning_rate_decay_epoch, init_seed, shuffle_seed, loss,
                                output_name, close_writer, iseed=None, nepoch_keep_missing=0,
                                save_model_each=0, save_checkpoint=None, tensorboard_writer=None):
    import random
    net = make_network(plan, batch_size, train=True)
    net.set_params_tf(plan.glorot_init(init_seed))
    holder = _Holder(net)
    rs = np.random.RandomState(shuffle_seed)
    stream = _repeat(train_datagen)
    buf = []
    fname = None
    index = 0
    for e in range(epochs):
        if nepoch_keep_missing > 0:
            random.seed(iseed + e // nepoch_keep_missing)        # :639-641 (user generators read `random`)
        lr_e = learning_rate * (0.5 ** (e / learning_rate_decay_epoch))
        net.lr.fill_(float(lr_e))
        pending = []                # (step in epoch, loss sums, l2) still on the device

        def drain():
            nonlocal index
            for ii_, sums, l2 in pending:
                s = sums.cpu().numpy()
                cost = (s[0] + s[1]) / s[3]
                if regularization_L2_beta != 0:
                    cost += regularization_L2_beta * float(l2.cpu()[0])
                rms = float(np.sqrt(s[2] / s[3]))
                loss.append(float(np.float32(cost)))
                if tensorboard_writer is not None:
                    tensorboard_writer.add_scalar("Validation/cost", float(cost), index)
                    tensorboard_writer.add_scalar("Validation/RMS", rms, index)
                index += 1
                if ii_ % 20 == 0:
                    print("Epoch: {}/{}...".format(e + 1, epochs),
                          "Training loss: {:.20f}".format(cost), "RMS: {:.20f}".format(rms), lr_e)
            del pending[:]

        for ii in range(ceil(train_len / batch_size)):
            xs, ts = [], []
            for _ in range(batch_size):
                while len(buf) < max(1, shuffle_buffer_size):
                    buf.append(next(stream))
                j = int(rs.randint(0, len(buf)))
                xs.append(buf[j][0])
                ts.append(buf[j][1])
                buf[j] = next(stream)
            n = _upload(net, xs, ts)
            net.forward(n)
            net.loss_backward(n, truth_uncertain, normalise=False)
            if regularization_L2_beta != 0:
                net.l2_value()
            net.adam_step(clip_norm=clip_grad, l2_beta=regularization_L2_beta,
                          denom=net.loss_sums[3:4], B=n)
            holder.step_count += 1
            # the loss is read back in bulk (every 20 steps / at the end of the epoch), not per step
            pending.append((ii, net.loss_sums.clone(), net.l2_out.clone()))
            if len(pending) >= 20:
                drain()
        drain()
        if ((e == epochs - 1) or ((save_each > 0) and (e % save_each == 0))) and outdir is not None:
            print("Save output", e)
            fname = output_name(outdir)
            it = test_datagen()
            for ii in range(ceil(test_len / batch_size)):
                xs, ts = [], []
                for _ in range(batch_size):
                    try:
                        x, t = next(it)
                    except StopIteration:
                        break
                    xs.append(x)
                    ts.append(t)
                if not xs:
                    break
                n = _upload(net, xs, ts)
                net.forward(n, save_signs=False)
                net.head_recon(n)
                savesample(fname, net.m_rec[:n].cpu().numpy(), net.s2_rec[:n].cpu().numpy(),
                           meandata, lon, lat, e, ii, ii * batch_size, transfun=transfun)
            close_writer(fname)
        if save_model_each > 0 and e % save_model_each == 0 and outdir is not None and save_checkpoint is not None:
            save_checkpoint(os.path.join(outdir, "model-{:03d}.ckpt".format(e + 1)), holder)
    return fname
