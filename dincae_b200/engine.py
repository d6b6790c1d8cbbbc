"""Device-side execution plan of the DINCAE network: buffers, parameter layout and the
sequence of C-ABI kernel calls that replaces the TF-1.15 graph of
DINCAE/__init__.py:433-603 (encoder, dense bottleneck, decoder, head, loss, gradients,
clip + Adam).  PyTorch is used only for device memory and streams.
"""
import math

import numpy as np
import torch

from . import _lib, layout
from ._lib import check, ptr, stream_ptr

NEG_SLOPE = 0.2          # tf.nn.leaky_relu default alpha (DINCAE/__init__.py:430)
COUT_ALIGN = 16


class Plan:
    """Static shapes of the network for one (H, W, nvar, filters) configuration."""

    def __init__(self, H, W, nvar=10, enc_nfilter_internal=(16, 24, 36, 54),
                 frac_dense_layer=(0.2,), skipconnections=(1, 2, 3, 4), lanes=4):
        """``lanes`` = 4: planar-4 fp32 activations (TF32 / fp32 kernels); 8: planar-8 bf16
        activations (kind::f16 kernels; every tensor padded to a multiple of 8 channels, so all
        plane counts in units of 4 channels are even)."""
        if lanes not in (4, 8):
            raise ValueError("lanes must be 4 (fp32 / TF32) or 8 (bf16)")
        self.lanes = int(lanes)
        self.H, self.W, self.nvar = int(H), int(W), int(nvar)
        self.enc_c = [self.nvar] + [int(c) for c in enc_nfilter_internal]
        self.L = len(self.enc_c) - 1
        self.dec_c = self.enc_c[:0:-1] + [2]              # index 0..L, like :385
        self.skip = set(int(s) for s in skipconnections)
        self.frac = [float(f) for f in frac_dense_layer]
        self.sizes = [(self.H, self.W)]
        for _ in range(self.L):
            h, w = self.sizes[-1]
            self.sizes.append(((h + 1) // 2, (w + 1) // 2))
        self.ep = [self._planes(c) for c in self.enc_c]
        self.dp = [self._planes(c) for c in self.dec_c]
        self.ep8 = [layout.planes8(c) for c in self.enc_c]
        self.dp8 = [layout.planes8(c) for c in self.dec_c]
        hL, wL = self.sizes[self.L]
        self.n_bottleneck = hL * wL * self.enc_c[self.L]
        self.flat_map, self.n_flat = layout.flat_map(hL, wL, self.enc_c[self.L], self.ep[self.L])
        if self.frac:
            fr = self.frac + self.frac[:-1][::-1]
            self.dense_units = [math.floor(self.n_bottleneck * f) for f in fr] + [self.n_bottleneck]
        else:
            self.dense_units = []
        # padded widths of the dense activations F[0..nd]
        self.dense_pad = [self.n_flat]
        for i, u in enumerate(self.dense_units):
            last = i == len(self.dense_units) - 1
            self.dense_pad.append(self.n_flat if last else layout.round_up(u, self.lanes))
        self._build_params()

    def _planes(self, c):
        """Planes of 4 channels a tensor of c channels occupies in this plan's layouts."""
        return layout.planes(c) if self.lanes == 4 else 2 * layout.planes8(c)

    # decoder conv l (1..L): sources and sizes -------------------------------------------
    def dec_sources(self, l):
        """(channels of upsampled source, channels of skip source or 0, resolution index)."""
        r = self.L - l                      # output resolution index (= l2 - 1)
        up = self.dec_c[l - 1]
        sk = self.enc_c[r] if l in self.skip else 0
        return up, sk, r

    def _build_params(self):
        ent = []

        def conv(name, splits, cout):
            cinp = sum(self._planes(c) for c in splits)
            cpad = layout.round_up(cout, COUT_ALIGN)
            ent.append(dict(name=name + "/kernel", kind="conv_w", tf_shape=(3, 3, sum(splits), cout),
                            splits=list(splits), cout=cout, cout_pad=cpad, cinp=cinp,
                            size=9 * cinp * cpad * 4))
            ent.append(dict(name=name + "/bias", kind="conv_b", tf_shape=(cout,), cout=cout,
                            size=cpad))

        k = 0

        def cname():
            nonlocal k
            n = "conv2d" if k == 0 else "conv2d_%d" % k
            k += 1
            return n

        for l in range(1, self.L + 1):
            conv(cname(), [self.enc_c[l - 1]], self.enc_c[l])
        nin, nin_pad = self.n_bottleneck, self.n_flat
        for i, u in enumerate(self.dense_units):
            name = "dense" if i == 0 else "dense_%d" % i
            upad = self.dense_pad[i + 1]
            ent.append(dict(name=name + "/kernel", kind="dense_w", tf_shape=(nin, u),
                            rows=nin_pad, cols=upad, size=nin_pad * upad, index=i))
            ent.append(dict(name=name + "/bias", kind="dense_b", tf_shape=(u,), cols=upad,
                            size=upad, index=i))
            nin, nin_pad = u, upad
        for l in range(1, self.L + 1):
            up, sk, _ = self.dec_sources(l)
            conv(cname(), [up] + ([sk] if sk else []), self.dec_c[l])
        # flat layout: all kernels first (they receive the L2 term), then all biases.  Among the
        # kernels, those whose gradient is final first in the backward pass (dense + decoder)
        # come first, so that the data-parallel trainer all-reduces ONE contiguous big bucket
        # while the encoder backward still runs, and one contiguous rest bucket afterwards.
        off = 0
        n_enc = 2 * self.L
        for e in ent[n_enc:]:
            if e["kind"].endswith("_w"):
                e["offset"] = off
                off = layout.round_up(off + e["size"], 64)
        self.n_early = off
        for e in ent[:n_enc]:
            if e["kind"].endswith("_w"):
                e["offset"] = off
                off = layout.round_up(off + e["size"], 64)
        self.n_decay = off
        for e in ent:
            if e["kind"].endswith("_b"):
                e["offset"] = off
                off = layout.round_up(off + e["size"], 64)
        self.n_params_flat = off
        self.entries = ent
        self.n_params_tf = sum(int(np.prod(e["tf_shape"])) for e in ent)

    def _dense_maps(self, i):
        nd = len(self.dense_units)
        rows = self.flat_map if i == 0 else np.arange(self.dense_units[i - 1])
        cols = self.flat_map if i == nd - 1 else np.arange(self.dense_units[i])
        return rows, cols

    def pack(self, tf_params):
        """List of numpy arrays in TF variable order -> flat float32 vector."""
        flat = np.zeros(self.n_params_flat, dtype=np.float32)
        assert len(tf_params) == len(self.entries)
        for e, p in zip(self.entries, tf_params):
            p = np.asarray(p, dtype=np.float32)
            assert tuple(p.shape) == tuple(e["tf_shape"]), (e["name"], p.shape, e["tf_shape"])
            o, n = e["offset"], e["size"]
            if e["kind"] == "conv_w":
                flat[o:o + n] = layout.pack_conv(p, e["splits"], e["cout_pad"], self.lanes).reshape(-1)
            elif e["kind"] == "conv_b":
                flat[o:o + e["cout"]] = p
            elif e["kind"] == "dense_w":
                rows, cols = self._dense_maps(e["index"])
                flat[o:o + n] = layout.pack_dense(p, rows, e["rows"], cols, e["cols"]).reshape(-1)
            else:
                _, cols = self._dense_maps(e["index"])
                flat[o + cols] = p
        return flat

    def unpack(self, flat):
        """Flat vector (params or gradients) -> list of numpy arrays in TF order/shape."""
        out = []
        flat = np.asarray(flat)
        for e in self.entries:
            o, n = e["offset"], e["size"]
            if e["kind"] == "conv_w":
                wp = flat[o:o + n].reshape(9, e["cinp"], e["cout_pad"], 4)
                out.append(layout.unpack_conv(wp, e["splits"], e["cout"], self.lanes))
            elif e["kind"] == "conv_b":
                out.append(flat[o:o + e["cout"]].copy())
            elif e["kind"] == "dense_w":
                rows, cols = self._dense_maps(e["index"])
                out.append(layout.unpack_dense(flat[o:o + n].reshape(e["rows"], e["cols"]), rows, cols))
            else:
                _, cols = self._dense_maps(e["index"])
                out.append(flat[o + cols].copy())
        return out

    def glorot_init(self, seed):
        """tf.layers defaults: Glorot-uniform kernels, zero biases."""
        rs = np.random.RandomState(seed)
        ps = []
        for e in self.entries:
            shp = e["tf_shape"]
            if e["kind"].endswith("_b"):
                ps.append(np.zeros(shp, dtype=np.float32))
                continue
            if len(shp) == 4:
                fan_in, fan_out = 9 * shp[2], 9 * shp[3]
            else:
                fan_in, fan_out = shp
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            ps.append(rs.uniform(-lim, lim, size=shp).astype(np.float32))
        return ps


class Network:
    """Buffers + kernel sequences for one Plan on the current CUDA device."""

    def __init__(self, plan, max_batch, train=True, device=None, factored=False, world=1, rank=0):
        """``factored``: keep the weight gradients of the dense bottleneck as their factors (layer
        input x, pre-activation gradient dz) instead of materialising x^T dz (csrc/dense_factored.cu);
        ``world`` / ``rank``: data-parallel layout of the per-rank gradient blob (see _setup_factored)."""
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DincaeError("dincae_b200 needs a CUDA device (no CPU fallback)")
        self.plan = plan
        self.Bmax = int(max_batch)
        self.train = train
        self.dev = torch.device(device if device is not None else "cuda")
        P, B, dev = plan, self.Bmax, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        self._setup_factored(factored, world, rank)

        self.params = torch.zeros(P.n_params_flat, **f32)
        if train:
            # 64 extra floats: tail that carries the loss sums through the gradient all-reduce
            self.grads_ext = None if self.factored else torch.zeros(P.n_params_flat + 64, **f32)
            self.grads = None if self.factored else self.grads_ext[:P.n_params_flat]
            self.adam_m = torch.zeros(P.n_params_flat, **f32)
            self.adam_v = torch.zeros(P.n_params_flat, **f32)
            self.adam_state = torch.zeros(8, **f32)
            self.lr = torch.zeros(1, **f32)
            self.reset_optimizer()

        # activations --------------------------------------------------------------------
        self.X = [torch.zeros(B, P.ep[k], P.sizes[k][0], P.sizes[k][1], 4, **f32)
                  for k in range(P.L + 1)]
        self.xtrue = torch.zeros(B, P.H, P.W, 2, **f32)
        self.n_noncloud = torch.zeros(1, dtype=torch.int32, device=dev)
        nd = len(P.dense_units)
        if self.factored:
            # the bottleneck activations that are gradient factors live in the per-rank blob
            self.X[P.L] = self.Fx[0].view(B, P.ep[P.L], P.sizes[P.L][0], P.sizes[P.L][1], 4)
        self.F = [self.X[P.L].view(B, -1)]
        for i in range(nd):
            last = i == nd - 1
            self.F.append(self.Fx[i + 1] if (self.factored and not last) else torch.zeros(B, P.dense_pad[i + 1], **f32))
        self.D = [self.F[-1].view(B, P.ep[P.L], P.sizes[P.L][0], P.sizes[P.L][1], 4)]
        for l in range(1, P.L + 1):
            h, w = P.sizes[P.L - l]
            self.D.append(torch.zeros(B, P.dp[l], h, w, 4, **f32))
        self.m_rec = torch.zeros(B, P.H, P.W, **f32)
        self.s2_rec = torch.zeros(B, P.H, P.W, **f32)

        ws = [self.lib.dincae_dense_workspace(B, P.dense_pad[i], P.dense_pad[i + 1])
              for i in range(nd)]
        if train:
            self.S = [None] + [
                torch.zeros(B, P.ep[l], P.sizes[l - 1][0], (P.sizes[l - 1][1] + 3) // 4,
                            dtype=torch.int16, device=dev) for l in range(1, P.L + 1)]
            self.GD = [torch.zeros_like(d) for d in self.D]           # dZ of decoder convs / dense out
            if self.factored:
                self.GD[0] = self.dZf[nd].view(self.D[0].shape)
            self.dX = [None] + [torch.zeros_like(self.X[k]) for k in range(1, P.L + 1)]
            self.dXskip = [None] * (P.L + 1)
            for l in range(1, P.L + 1):
                _, sk, r = P.dec_sources(l)
                if sk and r >= 1:
                    self.dXskip[r] = torch.zeros_like(self.X[r])
            self.dZ = [None] + [self.dZf[i + 1] if self.factored else torch.zeros_like(self.F[i + 1])
                                for i in range(nd)]
            if nd:
                self.dZ[nd] = self.GD[0].view(B, -1)
            self.loss_sums = torch.zeros(4, dtype=torch.float64, device=dev)
            self.loss_out = torch.zeros(2, **f32)
            self.l2_out = torch.zeros(1, **f32)
            ws.append(self.lib.dincae_head_loss_workspace(B, P.H, P.W))
            ws.append(self.lib.dincae_adam_workspace(P.n_params_flat))
            ws += self._factored_workspaces()
            for e in P.entries:
                if e["kind"] == "conv_w":
                    h, w = self._conv_res(e)
                    ws.append(self.lib.dincae_conv3x3_wgrad_workspace(e["cinp"], 0, e["cout_pad"], B, h, w))
        self.workspace = torch.zeros(max(ws + [256]), dtype=torch.uint8, device=dev)
        # Weight gradients run on a side stream next to the data-gradient chain (they only
        # share inputs), with their own scratch so the two streams never touch the same bytes.
        import os
        self._side_parts = int(os.environ.get("DINCAE_SIDE_PARTS", "7"))
        self.side = torch.cuda.Stream(device=dev) if (train and os.environ.get("DINCAE_NO_SIDE_STREAM") != "1") else None
        self.workspace_side = torch.zeros_like(self.workspace) if train else None
        self._ent = {e["name"]: e for e in P.entries}
        self._conv_names = [e["name"][:-7] for e in P.entries if e["kind"] == "conv_w"]
        self.tracer = None       # optional callable(label, fn, args, alg_bytes, alg_flops)

    # ---- dropout (optional; inert in the reference) -------------------------------------
    dropout_rate = 0.0          # > 0: tf.layers.dropout as it was meant to act, after each dense layer
    dropout_seed = 0
    dropout_keys = None         # device int64 [B]: per-frame keys (the trainer's noise sequence numbers)
    dropout_masks = None        # optional list of uint8 [B, N] tensors that receive the keep masks

    def _dropout_fwd(self, i, B):
        """Dropout on the output of dense layer i (training forward only)."""
        if self.dropout_rate <= 0.0:
            return
        N = self.plan.dense_pad[i + 1]
        mk = self.dropout_masks[i] if self.dropout_masks is not None else None
        self._call("dropout_fwd(%d)" % i, "dincae_dropout_fwd", (
            ptr(self.F[i + 1]), B, N, N, float(self.dropout_rate), int(self.dropout_seed) & (2 ** 64 - 1), i, 0,
            ptr(self.dropout_keys), ptr(mk), stream_ptr()), 8 * B * N, 0)

    def _dropout_bwd(self, i, B):
        """1 / (1 - rate) on the gradient of dense layer i's (gated) output; the mask itself is
        applied by the ReLU gate of the kernel that produced this gradient (csrc/dropout.cu)."""
        if self.dropout_rate <= 0.0:
            return
        N = self.plan.dense_pad[i + 1]
        self._call("dropout_bwd(%d)" % i, "dincae_scale_rows", (
            ptr(self.dZ[i + 1]), B, N, N, 1.0 / (1.0 - float(self.dropout_rate)), stream_ptr()), 8 * B * N, 0)

    # ---- factored dense gradients / per-rank gradient blob ---------------------------------------
    def _setup_factored(self, factored, world, rank):
        """Factored mode: everything a rank contributes to the optimiser step sits in ONE
        contiguous per-rank blob of floats,
            [x_0 | dz_1 | x_1 | dz_2 | ... | gradients of all other parameters | 4 loss sums (f64)],
        so that data-parallel ranks exchange it with a single all-gather into ``blob_all``
        [world][blob_len]; the optimiser kernels read all ranks' blobs in place, in rank order."""
        P, B = self.plan, self.Bmax
        nd = len(P.dense_units)
        self.factored = bool(factored) and nd > 0 and self.train
        self.world, self.rank = int(world), int(rank)
        self.n_dense = 0
        if not self.factored:
            return
        f32 = dict(dtype=torch.float32, device=self.dev)
        ents = [e for e in P.entries if e["kind"] == "dense_w"]
        assert ents[0]["offset"] == 0, "dense kernels are expected at the start of the flat buffer"
        self.n_dense = layout.round_up(ents[-1]["offset"] + ents[-1]["size"], 64)
        self.n_rest = P.n_params_flat - self.n_dense
        off = 0
        self._fact = []
        for i in range(nd):
            kx, kz = P.dense_pad[i], P.dense_pad[i + 1]
            self._fact.append(dict(x=off, z=off + layout.round_up(B * kx, 64), K=kx, N=kz))
            off = self._fact[-1]["z"] + layout.round_up(B * kz, 64)
        self._rest_off = off
        off += layout.round_up(self.n_rest, 64)
        self._loss_off = off
        off += 64
        self.blob_len = off
        self.blob_all = torch.zeros(self.world, off, **f32)
        self.blob = self.blob_all[self.rank]
        self.Fx = [self.blob[f["x"]: f["x"] + B * f["K"]].view(B, f["K"]) for f in self._fact]
        self.dZf = [None] + [self.blob[f["z"]: f["z"] + B * f["N"]].view(B, f["N"]) for f in self._fact]
        self.grads_rest = self.blob[self._rest_off: self._rest_off + self.n_rest]
        self.loss_sums_local = self.blob[self._loss_off: self._loss_off + 8].view(torch.float64)
        self.extra_sumsq = torch.zeros(1, dtype=torch.float64, device=self.dev)

    def _factored_workspaces(self):
        if not self.factored:
            return []
        return [self.lib.dincae_dense_factored_workspace(self.world * self.Bmax, f["K"], f["N"]) for f in self._fact] + \
            [self.lib.dincae_adam_multi_workspace(self.n_rest)]

    def _g(self, name):
        """Gradient slice of variable ``name`` (in the per-rank blob when factored)."""
        e = self._ent[name]
        if self.factored:
            o = e["offset"] - self.n_dense
            assert o >= 0, "dense kernels have no materialised gradient in factored mode"
            return self.grads_rest[o: o + e["size"]]
        return self.grads[e["offset"]: e["offset"] + e["size"]]

    def _loss_sums_dst(self):
        return self.loss_sums_local if self.factored else self.loss_sums

    def pack_loss_sums(self):
        """Loss sums -> (hi, lo) float pairs in the tail of the gradient buffer (materialised path,
        data parallel: ONE float32 all-reduce of ``grads_ext`` carries gradients and loss sums)."""
        n = self.plan.n_params_flat
        self._call("pack_loss_sums", "dincae_f64_to_hilo", (
            ptr(self.loss_sums), 4, ptr(self.grads_ext[n:]), stream_ptr()), 64, 0)

    def unpack_loss_sums(self):
        n = self.plan.n_params_flat
        self._call("unpack_loss_sums", "dincae_hilo_to_f64", (
            ptr(self.grads_ext[n:]), 4, ptr(self.loss_sums), stream_ptr()), 64, 0)

    def _dense_weight_grad(self, i, B, part=2):
        """Weight + bias gradient of dense layer i -- or, factored, only the bias gradient."""
        P = self.plan
        name = "dense" if i == 0 else "dense_%d" % i
        K, N = P.dense_pad[i], P.dense_pad[i + 1]
        kk, nn = self._ent[name + "/kernel"]["tf_shape"]
        if self.factored:
            self._call("dense_bias_grad(%d)" % i, "dincae_dense_bias_grad", (
                ptr(self.dZ[i + 1]), B, N, N, ptr(self._g(name + "/bias")), stream_ptr()), 4 * (B * nn + nn), 0)
            return
        self._fork(lambda: self._call("dense_bwd_weight(%d)" % i, "dincae_dense_bwd_weight", (
            ptr(self.F[i]), ptr(self.dZ[i + 1]), B, K, N,
            ptr(self._g(name + "/kernel")), ptr(self._g(name + "/bias")),
            stream_ptr()), 4 * (B * kk + B * nn + kk * nn), 2 * B * kk * nn), part)

    def _fork(self, fn, part=1):
        """Run ``fn`` (which enqueues kernels) on the side stream, after everything enqueued so
        far on the current stream; joined again by ``_join`` (works inside graph capture)."""
        if self.side is None or not (self._side_parts & part):
            fn()
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            fn()

    def _join(self):
        if self.side is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def _call(self, label, fname, args, alg_bytes=0, alg_flops=0):
        """Enqueue one C-ABI call.  ``alg_bytes`` / ``alg_flops`` are the ALGORITHMIC traffic
        and work of the call (fp32, logical channel counts, every tensor touched once; see
        DESIGN.md) -- used by bench.py's roofline report through ``self.tracer``."""
        fn = getattr(self.lib, fname)
        if self.tracer is None:
            check(fn(*args), label)
        else:
            self.tracer(label, fn, args, alg_bytes, alg_flops)

    # ---- parameters -----------------------------------------------------------------------
    def _conv_res(self, e):
        idx = [x["name"] for x in self.plan.entries if x["kind"] == "conv_w"].index(e["name"])
        L = self.plan.L
        return self.plan.sizes[idx] if idx < L else self.plan.sizes[L - (idx - L + 1)]

    def reset_optimizer(self, beta1=0.9, beta2=0.999):
        self.adam_m.zero_()
        self.adam_v.zero_()
        self.adam_state.zero_()
        self.adam_state[0] = beta1
        self.adam_state[1] = beta2

    def set_params_tf(self, tf_params):
        self.params.copy_(torch.from_numpy(self.plan.pack(tf_params)))

    def get_params_tf(self):
        return self.plan.unpack(self.params.cpu().numpy())

    def flat_grads(self, B=None):
        """Flat gradient vector (a copy).  Factored mode: the dense kernels' gradients are expanded
        from their factors and the other gradients summed over the ranks' blobs."""
        if not self.factored:
            return self.grads.clone()
        B = self.Bmax if B is None else B
        flat = torch.zeros(self.plan.n_params_flat, dtype=torch.float32, device=self.dev)
        flat[self.n_dense:] = self.blob_all[:, self._rest_off: self._rest_off + self.n_rest].sum(dim=0)
        ba = self.blob_all
        for i, f in enumerate(self._fact):
            e = self._ent[("dense" if i == 0 else "dense_%d" % i) + "/kernel"]
            check(self.lib.dincae_dense_factored_expand(
                ba.data_ptr() + 4 * f["x"], self.blob_len, f["K"], ba.data_ptr() + 4 * f["z"], self.blob_len, f["N"],
                self.world, B, f["K"], f["N"], flat.data_ptr() + 4 * e["offset"], f["N"], stream_ptr()), "expand")
        return flat

    def get_grads_tf(self, B=None):
        return self.plan.unpack(self.flat_grads(B).cpu().numpy())

    def _w(self, name, buf=None):
        e = self._ent[name]
        buf = self.params if buf is None else buf
        return buf[e["offset"]: e["offset"] + e["size"]]

    def _enc_name(self, l):
        return self._conv_names[l - 1]

    def _dec_name(self, l):
        return self._conv_names[self.plan.L + l - 1]

    # ---- forward --------------------------------------------------------------------------
    def params_changed(self):
        """Call after writing ``self.params`` directly (restore, broadcast): refreshes copies
        derived from the parameters.  Nothing to do for the fp32 / TF32 network."""

    def build_input(self, cube, frame_idx, cloud_idx, B, **kw):
        """Stack the network input of frames ``frame_idx[:B]`` from the device cube
        (``DeviceCube.build_batch``) into this network's input buffer."""
        cube.build_batch(frame_idx, cloud_idx, B, self.X[0], self.xtrue, self.n_noncloud, **kw)

    def forward(self, B, save_signs=None):
        """Encoder -> dense -> decoder on self.X[0][:B]; leaves xrec in self.D[L]."""
        P, lib, st = self.plan, self.lib, stream_ptr()
        save_signs = self.train if save_signs is None else save_signs
        for l in range(1, P.L + 1):
            h, w = P.sizes[l - 1]
            e = self._ent[self._enc_name(l) + "/kernel"]
            ci, co = P.enc_c[l - 1], P.enc_c[l]
            hh, wh = P.sizes[l]
            nb = 4 * B * (h * w * ci + hh * wh * co) + 36 * ci * co + (B * h * w * co // 8 if save_signs else 0)
            self._call("conv3x3_fwd(enc %d)" % l, "dincae_conv3x3_fwd", (
                ptr(self.X[l - 1]), P.ep[l - 1], 0, None, 0,
                ptr(self._w(e["name"])), ptr(self._w(self._enc_name(l) + "/bias")), e["cout_pad"],
                B, h, w, P.ep[l], NEG_SLOPE, None, ptr(self.X[l]),
                ptr(self.S[l]) if save_signs else None, st), nb, 18 * ci * co * B * h * w)
        for i in range(len(P.dense_units)):
            name = "dense" if i == 0 else "dense_%d" % i
            kk, nn = self._ent[name + "/kernel"]["tf_shape"]
            self._call("dense_fwd(%d)" % i, "dincae_dense_fwd", (
                ptr(self.F[i]), ptr(self._w(name + "/kernel")), ptr(self._w(name + "/bias")),
                B, P.dense_pad[i], P.dense_pad[i + 1], 1, ptr(self.F[i + 1]),
                ptr(self.workspace), self.workspace.numel(), st),
                4 * (B * kk + kk * nn + B * nn), 2 * B * kk * nn)
            if save_signs:
                self._dropout_fwd(i, B)
        for l in range(1, P.L + 1):
            up, sk, r = P.dec_sources(l)
            h, w = P.sizes[r]
            e = self._ent[self._dec_name(l) + "/kernel"]
            hh, wh = P.sizes[r + 1]
            co = P.dec_c[l]
            nb = 4 * B * (hh * wh * up + h * w * (sk + co)) + 36 * (up + sk) * co
            self._call("conv3x3_fwd(dec %d)" % l, "dincae_conv3x3_fwd", (
                ptr(self.D[l - 1]), P.dp[l - 1], 1,
                ptr(self.X[r]) if sk else None, P.ep[r] if sk else 0,
                ptr(self._w(e["name"])), ptr(self._w(self._dec_name(l) + "/bias")), e["cout_pad"],
                B, h, w, P.dp[l], NEG_SLOPE, ptr(self.D[l]), None, None, st),
                nb, 18 * (up + sk) * co * B * h * w)

    def head_recon(self, B):
        P = self.plan
        self._call("head_recon", "dincae_head_recon", (
            ptr(self.D[P.L]), B, P.H, P.W, ptr(self.m_rec), ptr(self.s2_rec), stream_ptr()),
            16 * B * P.H * P.W, 0)

    def head_recon_records(self, B, meandata, never, fill, mean_is_f32, records, big_endian=True):
        """Head + the arithmetic of ``savesample`` (:237-248, 288-291) as file-ready records
        [B][2][H][W] (see include/dincae_b200.h)."""
        P = self.plan
        self._call("head_recon_records", "dincae_head_recon_records", (
            ptr(self.D[P.L]), B, P.H, P.W, ptr(meandata), ptr(never), float(fill), int(mean_is_f32),
            int(big_endian), ptr(records), stream_ptr()),
            (16 + 8) * B * P.H * P.W, 0)

    # ---- loss + backward --------------------------------------------------------------------
    def grad_buckets(self):
        """(start, end) ranges of the flat gradient buffer in the order they become final
        during ``loss_backward``: [dense + decoder kernels] after phase 1 (97 % of the bytes at
        config A), then [encoder kernels + all biases] after phase 2."""
        return (0, self.plan.n_early), (self.plan.n_early, self.plan.n_params_flat)

    def loss_backward(self, B, truth_uncertain=False, normalise=True, phase=0):
        """head + NLL, then gradients of every parameter into self.grads.
        phase 1 = loss head, decoder and dense backward; phase 2 = encoder backward; 0 = both
        (the data-parallel trainer all-reduces the phase-1 gradients while phase 2 runs)."""
        if phase == 0:
            self.loss_backward(B, truth_uncertain, normalise, 1)
            self.loss_backward(B, truth_uncertain, normalise, 2)
            return
        P, lib, st = self.plan, self.lib, stream_ptr()
        L = P.L
        wsp, wsn = ptr(self.workspace), self.workspace.numel()
        wss = ptr(self.workspace_side)
        has_dense = len(P.dense_units) > 0
        if phase == 2:
            return self._backward_encoder(B, has_dense, wss, wsn, st)
        self._call("head_loss", "dincae_head_loss", (
            ptr(self.D[L]), ptr(self.xtrue), B, P.H, P.W, int(bool(truth_uncertain)), NEG_SLOPE,
            ptr(self.n_noncloud) if normalise else None, ptr(self.GD[L]), ptr(self._loss_sums_dst()),
            ptr(self.loss_out), wsp, wsn, st), 24 * B * P.H * P.W, 0)
        has_dense = len(P.dense_units) > 0
        for l in range(L, 0, -1):
            up, sk, r = P.dec_sources(l)
            h, w = P.sizes[r]
            name = self._dec_name(l)
            e = self._ent[name + "/kernel"]
            c0p = P.dp[l - 1]
            c1p = P.ep[r] if sk else 0
            hh, wh = P.sizes[r + 1]
            co = P.dec_c[l]
            nb = 4 * B * (hh * wh * up + h * w * (sk + co)) + 36 * (up + sk) * co
            self._fork(lambda: self._call("conv3x3_wgrad(dec %d)" % l, "dincae_conv3x3_wgrad", (
                ptr(self.D[l - 1]), c0p, 1, ptr(self.X[r]) if sk else None, c1p,
                ptr(self.GD[l]), None, None, NEG_SLOPE, P.dp[l], e["cout_pad"], B, h, w,
                ptr(self._g(name + "/kernel")), ptr(self._g(name + "/bias")),
                wss, wsn, stream_ptr()), nb, 18 * (up + sk) * co * B * h * w), 1)
            prev_slope = NEG_SLOPE if l > 1 else (0.0 if has_dense else 1.0)
            want_skip = bool(sk) and r >= 1
            cin_eff = up + (sk if want_skip else 0)
            nb = 4 * B * (h * w * co + 2 * hh * wh * up + (h * w * sk if want_skip else 0)) + 36 * cin_eff * co
            self._call("conv3x3_dgrad(dec %d)" % l, "dincae_conv3x3_dgrad", (
                ptr(self.GD[l]), None, None, NEG_SLOPE, P.dp[l],
                ptr(self._w(name + "/kernel")), e["cinp"], e["cout_pad"], B, h, w,
                c0p, ptr(self.D[l - 1]), prev_slope, ptr(self.GD[l - 1]),
                c1p if want_skip else 0, ptr(self.dXskip[r]) if want_skip else None, None, st),
                nb, 18 * cin_eff * co * B * h * w)
        nd = len(P.dense_units)
        for i in range(nd - 1, -1, -1):
            name = "dense" if i == 0 else "dense_%d" % i
            K, N = P.dense_pad[i], P.dense_pad[i + 1]
            kk, nn = self._ent[name + "/kernel"]["tf_shape"]
            self._dropout_bwd(i, B)
            self._dense_weight_grad(i, B)
            dst = self.dZ[i] if i > 0 else self.dX[L].view(self.Bmax, -1)
            self._call("dense_bwd_data(%d)" % i, "dincae_dense_bwd_data", (
                ptr(self.dZ[i + 1]), ptr(self._w(name + "/kernel")),
                ptr(self.F[i]) if i > 0 else None, B, K, N, ptr(dst), wsp, wsn, st),
                4 * (B * nn + kk * nn + B * kk), 2 * B * kk * nn)
        self._join()

    def _backward_encoder(self, B, has_dense, wss, wsn, st):
        P, L = self.plan, self.plan.L
        top_grad = self.dX[L] if has_dense else self.GD[0]
        for l in range(L, 0, -1):
            h, w = P.sizes[l - 1]
            name = self._enc_name(l)
            e = self._ent[name + "/kernel"]
            gpool = top_grad if l == L else self.dX[l]
            ci, co = P.enc_c[l - 1], P.enc_c[l]
            hh, wh = P.sizes[l]
            nb = 4 * B * (h * w * ci + hh * wh * co) + B * h * w * co // 8 + 36 * ci * co
            self._fork(lambda: self._call("conv3x3_wgrad(enc %d)" % l, "dincae_conv3x3_wgrad", (
                ptr(self.X[l - 1]), P.ep[l - 1], 0, None, 0,
                None, ptr(gpool), ptr(self.S[l]), NEG_SLOPE, P.ep[l], e["cout_pad"], B, h, w,
                ptr(self._g(name + "/kernel")), ptr(self._g(name + "/bias")),
                wss, wsn, stream_ptr()), nb, 18 * ci * co * B * h * w), 4)
            if l > 1:
                has_add = self.dXskip[l - 1] is not None
                nb = 4 * B * (hh * wh * co + h * w * ci * (2 if has_add else 1)) + B * h * w * co // 8 + 36 * ci * co
                self._call("conv3x3_dgrad(enc %d)" % l, "dincae_conv3x3_dgrad", (
                    None, ptr(gpool), ptr(self.S[l]), NEG_SLOPE, P.ep[l],
                    ptr(self._w(name + "/kernel")), e["cinp"], e["cout_pad"], B, h, w,
                    0, None, 1.0, None,
                    P.ep[l - 1], ptr(self.dX[l - 1]),
                    ptr(self.dXskip[l - 1]) if has_add else None, st),
                    nb, 18 * ci * co * B * h * w)
        self._join()

    def l2_value(self):
        """0.5 * sum of squares of all kernels -> self.l2_out (device)."""
        self._call("l2_half_sumsq", "dincae_l2_half_sumsq", (
            ptr(self.params), self.plan.n_decay, ptr(self.l2_out), ptr(self.workspace),
            self.workspace.numel(), stream_ptr()), 4 * self.plan.n_params_tf, 0)

    def adam_step(self, clip_norm=5.0, l2_beta=0.0, beta1=0.9, beta2=0.999, eps=1e-8,
                  grad_scale=1.0, denom=None, B=None):
        """clip_by_global_norm + Adam on self.grads; learning rate read from self.lr.
        Factored mode: ``B`` rows of every rank's factors are valid; the loss sums of all ranks
        are reduced into ``self.loss_sums`` first (``denom`` usually points at its 4th entry)."""
        if self.factored:
            return self._adam_step_factored(clip_norm, l2_beta, beta1, beta2, eps, grad_scale, denom,
                                            self.Bmax if B is None else B)
        self._call("adam_step", "dincae_adam_step", (
            ptr(self.params), ptr(self.grads), ptr(self.adam_m), ptr(self.adam_v),
            self.plan.n_params_flat, self.plan.n_decay, l2_beta, clip_norm, ptr(self.lr),
            beta1, beta2, eps, grad_scale, ptr(denom), ptr(self.adam_state),
            ptr(self.workspace), self.workspace.numel(), stream_ptr()),
            28 * self.plan.n_params_tf, 0)

    def _adam_step_factored(self, clip_norm, l2_beta, beta1, beta2, eps, grad_scale, denom, B):
        if l2_beta != 0.0:
            raise ValueError("the factored optimiser path has no L2 term (use factored=False)")
        st = stream_ptr()
        ba, L = self.blob_all, self.blob_len
        base = ba.data_ptr()
        self._call("sum_loss_sums", "dincae_sum_ranks_f64", (
            base + 4 * self._loss_off, L // 2, self.world, 4, ptr(self.loss_sums), st), 32 * self.world, 0)
        self.extra_sumsq.zero_()
        wsp, wsn = ptr(self.workspace), self.workspace.numel()
        for i, f in enumerate(self._fact):
            self._call("dense_grad_sumsq(%d)" % i, "dincae_dense_factored_sumsq", (
                base + 4 * f["x"], L, f["K"], base + 4 * f["z"], L, f["N"], self.world, B, f["K"], f["N"],
                ptr(self.extra_sumsq), wsp, wsn, st), 4 * self.world * B * (f["K"] + f["N"]), 0)
        nd_, nr = self.n_dense, self.n_rest
        self._call("adam_step(conv + biases)", "dincae_adam_step_multi", (
            ptr(self.params[nd_:]), base + 4 * self._rest_off, L, self.world, ptr(self.adam_m[nd_:]),
            ptr(self.adam_v[nd_:]), nr, clip_norm, ptr(self.lr), beta1, beta2, eps, grad_scale, ptr(denom),
            ptr(self.extra_sumsq), ptr(self.adam_state), wsp, wsn, st),
            (24 + 8 * self.world) * nr, 0)
        for i, f in enumerate(self._fact):
            e = self._ent[("dense" if i == 0 else "dense_%d" % i) + "/kernel"]
            o = e["offset"]
            kk, nn = e["tf_shape"]
            w8 = self.w8[i] if getattr(self, "w8", None) else None
            self._call("adam_dense_factored(%d)" % i, "dincae_adam_update_factored", (
                ptr(self.params[o:]), ptr(self.adam_m[o:]), ptr(self.adam_v[o:]), f["N"],
                base + 4 * f["x"], L, f["K"], base + 4 * f["z"], L, f["N"], self.world, B, f["K"], f["N"],
                grad_scale, ptr(denom), beta1, beta2, eps, ptr(self.adam_state), ptr(w8), st),
                (26 if w8 is not None else 24) * kk * nn, 2 * self.world * B * kk * nn)


def make_network(plan, max_batch, train=True, device=None, **kw):
    """Network for ``plan``: planar-4 fp32 / TF32 kernels, or (``plan.lanes == 8``) the bf16 path."""
    cls = NetworkBF16 if plan.lanes == 8 else Network
    return cls(plan, max_batch, train=train, device=device, **kw)


class NetworkBF16(Network):
    """bf16 execution of the same graph (BASELINE configs[3]: "bf16 training"): activations and
    their gradients are planar-8 bf16, the convolutions run ``tcgen05.mma kind::f16`` with fp32
    accumulation (``csrc/conv_bf16.cu``, ``csrc/conv_wgrad_bf16.cu``), parameters, gradients of
    the parameters, Adam state, the dense bottleneck, the network output and the loss stay fp32.
    The fp32 master kernels are converted to their bf16 operand copies once per step."""

    def __init__(self, plan, max_batch, train=True, device=None, factored=False, world=1, rank=0):
        import os
        if plan.lanes != 8:
            raise ValueError("NetworkBF16 needs Plan(lanes=8)")
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DincaeError("dincae_b200 needs a CUDA device (no CPU fallback)")
        self.plan = plan
        self.Bmax = int(max_batch)
        self.train = train
        self.dev = torch.device(device if device is not None else "cuda")
        P, B, dev = plan, self.Bmax, self.dev
        f32 = dict(dtype=torch.float32, device=dev)
        b16 = dict(dtype=torch.bfloat16, device=dev)
        L = P.L
        self._setup_factored(factored, world, rank)

        self.params = torch.zeros(P.n_params_flat, **f32)
        if train:
            self.grads_ext = None if self.factored else torch.zeros(P.n_params_flat + 64, **f32)
            self.grads = None if self.factored else self.grads_ext[:P.n_params_flat]
            self.adam_m = torch.zeros(P.n_params_flat, **f32)
            self.adam_v = torch.zeros(P.n_params_flat, **f32)
            self.adam_state = torch.zeros(8, **f32)
            self.lr = torch.zeros(1, **f32)
            self.reset_optimizer()

        # the batch builder writes planar-4 fp32 with exactly ceil(nvar/4) planes
        self.x0_planes = layout.planes(P.nvar)
        self.input_p8 = False                 # True once build_input() fills X8[0] directly
        self.X = [torch.zeros(B, self.x0_planes, P.H, P.W, 4, **f32)]
        self.X8 = [torch.zeros(B, P.ep8[k], P.sizes[k][0], P.sizes[k][1], 8, **b16) for k in range(L + 1)]
        self.xtrue = torch.zeros(B, P.H, P.W, 2, **f32)
        self.n_noncloud = torch.zeros(1, dtype=torch.int32, device=dev)
        nd = len(P.dense_units)
        hL, wL = P.sizes[L]
        self.F = [self.Fx[0] if self.factored else torch.zeros(B, P.n_flat, **f32)] if nd else []
        for i in range(nd):
            last = i == nd - 1
            self.F.append(self.Fx[i + 1] if (self.factored and not last) else torch.zeros(B, P.dense_pad[i + 1], **f32))
        self.D8 = [torch.zeros(B, P.ep8[L], hL, wL, 8, **b16) if nd else self.X8[L]]
        for l in range(1, L):
            h, w = P.sizes[L - l]
            self.D8.append(torch.zeros(B, P.dp8[l], h, w, 8, **b16))
        self.D8.append(None)                                    # the last layer only writes xrec
        self.xrec = torch.zeros(B, 1, P.H, P.W, 4, **f32)
        self.D = [None] * L + [self.xrec]                       # head kernels read D[L]
        self.m_rec = torch.zeros(B, P.H, P.W, **f32)
        self.s2_rec = torch.zeros(B, P.H, P.W, **f32)

        # bf16 operand copies of the conv kernels: forward copy per layer, data-gradient copies
        # per channel range of at most 256 input channels
        self._wbf = {}
        total = 0
        import ctypes
        for e in P.entries:
            if e["kind"] != "conv_w":
                continue
            name = e["name"][:-7]
            nde = ctypes.c_size_t(0)
            nf = self.lib.dincae_conv_weights_bf16_elems(e["cinp"], e["cout_pad"], 16, ctypes.byref(nde))
            info = dict(wf=(total, nf), wd=[])
            total += layout.round_up(nf, 64)
            if train:
                for (ci_first, lo, hi) in self._dgrad_ranges(name):
                    ndr = layout.round_up(8 * (lo + hi), 16)
                    self.lib.dincae_conv_weights_bf16_elems(e["cinp"], e["cout_pad"], ndr, ctypes.byref(nde))
                    info["wd"].append(dict(ci_first=ci_first, lo=lo, hi=hi, nd=ndr, off=total, n=nde.value))
                    total += layout.round_up(nde.value, 64)
            self._wbf[name] = info
        self.wbf = torch.zeros(total, **b16)

        # tiled bf16 copies of the dense kernels for the tensor-core GEMMs (csrc/dense_tc.cu)
        self.w8 = [torch.zeros(P.dense_pad[i] * P.dense_pad[i + 1], **b16) for i in range(nd)]
        ws = [self.lib.dincae_dense_tc_workspace(min(B, 256), P.dense_pad[i], P.dense_pad[i + 1]) for i in range(nd)]
        ws += [self.lib.dincae_dense_workspace(B, P.dense_pad[i], P.dense_pad[i + 1]) for i in range(nd)]
        if train:
            self.S8 = [None] + [
                torch.zeros(B, P.ep8[l], P.sizes[l - 1][0], 4 * ((P.sizes[l - 1][1] + 3) // 4),
                            dtype=torch.uint8, device=dev) for l in range(1, L + 1)]
            self.GD8 = [torch.zeros_like(self.D8[l]) if (l > 0 or nd) else None for l in range(L)]
            self.GD8.append(torch.zeros(B, 1, P.H, P.W, 8, **b16))
            self.dX8 = [None] + [torch.zeros_like(self.X8[k]) for k in range(1, L + 1)]
            if not nd:
                self.GD8[0] = self.dX8[L]                       # decoder 1 back-propagates straight into the encoder
            self.dXskip8 = [None] * (L + 1)
            for l in range(1, L + 1):
                _, sk, r = P.dec_sources(l)
                if sk and r >= 1:
                    self.dXskip8[r] = torch.zeros_like(self.X8[r])
            self.dZ = [None] + [self.dZf[i + 1] if self.factored else torch.zeros_like(self.F[i + 1])
                                for i in range(nd)]
            self.dF0 = torch.zeros(B, P.n_flat, **f32) if nd else None
            self.loss_sums = torch.zeros(4, dtype=torch.float64, device=dev)
            self.loss_out = torch.zeros(2, **f32)
            self.l2_out = torch.zeros(1, **f32)
            ws.append(self.lib.dincae_head_loss_workspace(B, P.H, P.W))
            ws.append(self.lib.dincae_adam_workspace(P.n_params_flat))
            ws += self._factored_workspaces()
            for e in P.entries:
                if e["kind"] == "conv_w":
                    ws.append(self.lib.dincae_conv3x3_wgrad_bf16_workspace(e["cinp"], e["cout_pad"]))
        self.workspace = torch.zeros(max(ws + [256]), dtype=torch.uint8, device=dev)
        self._side_parts = int(os.environ.get("DINCAE_SIDE_PARTS", "7"))
        self.side = torch.cuda.Stream(device=dev) if (train and os.environ.get("DINCAE_NO_SIDE_STREAM") != "1") else None
        self.workspace_side = torch.zeros_like(self.workspace) if train else None
        self._ent = {e["name"]: e for e in P.entries}
        self._conv_names = [e["name"][:-7] for e in P.entries if e["kind"] == "conv_w"]
        self.tracer = None

    # ---- bf16 operand copies of the kernels -----------------------------------------------------
    def _dgrad_ranges(self, name):
        """[(first input channel, planes of the half-res source, planes of the skip source)]: the
        channel ranges (at most 256 channels = one MMA each) whose data gradient layer ``name``
        computes."""
        P = self.plan
        names = [e["name"][:-7] for e in P.entries if e["kind"] == "conv_w"]
        idx = names.index(name)
        if idx < P.L:                                   # encoder layer idx+1: gradient of its input
            return [(0, 0, P.ep8[idx])] if idx >= 1 else []
        l = idx - P.L + 1
        up, sk, r = P.dec_sources(l)
        lo = P.dp8[l - 1] if l > 1 else P.ep8[P.L]
        hi = P.ep8[r] if (sk and r >= 1) else 0
        if 8 * (lo + hi) <= 256:
            return [(0, lo, hi)]
        return [(0, lo, 0)] + ([(8 * lo, 0, hi)] if hi else [])

    def _wf(self, name):
        off, n = self._wbf[name]["wf"]
        return self.wbf[off: off + n]

    def tile_dense_weights(self):
        """fp32 master dense kernels -> tiled bf16 copies."""
        P, st = self.plan, stream_ptr()
        for i in range(len(P.dense_units)):
            name = "dense" if i == 0 else "dense_%d" % i
            K, N = P.dense_pad[i], P.dense_pad[i + 1]
            self._call("dense_tile_weights(%d)" % i, "dincae_dense_tc_tile_weights", (
                ptr(self._w(name + "/kernel")), N, K, N, ptr(self.w8[i]), st), 6 * K * N, 0)

    def set_params_tf(self, tf_params):
        super().set_params_tf(tf_params)
        self.tile_dense_weights()

    def params_changed(self):
        """The tiled bf16 dense copies follow the fp32 master weights (the factored Adam kernels
        keep them current during training; any other write to ``params`` must come through here)."""
        self.tile_dense_weights()

    def prepare_weights(self):
        """fp32 master kernels -> bf16 operand copies (one small launch per layer and copy).  The
        dense copies are refreshed here too unless the factored Adam kernel keeps them current."""
        st = stream_ptr()
        if not self.factored and self.train:
            self.tile_dense_weights()
        for name, info in self._wbf.items():
            e = self._ent[name + "/kernel"]
            w = self._w(e["name"])
            wds = info["wd"]
            first = wds[0] if wds else None
            self._call("conv_weights_bf16(%s)" % name, "dincae_conv_weights_bf16", (
                ptr(w), e["cinp"], e["cout_pad"], first["ci_first"] if first else 0,
                first["nd"] if first else 16, ptr(self._wf(name)),
                ptr(self.wbf[first["off"]:]) if first else None, st), 6 * e["size"], 0)
            for d in wds[1:]:
                self._call("conv_weights_bf16(%s)" % name, "dincae_conv_weights_bf16", (
                    ptr(w), e["cinp"], e["cout_pad"], d["ci_first"], d["nd"], None,
                    ptr(self.wbf[d["off"]:]), st), 6 * e["size"], 0)

    def build_input(self, cube, frame_idx, cloud_idx, B, **kw):
        """The batch builder writes the bf16 P8 input directly (dincae_build_batch_p8): no fp32
        copy, no conversion pass.  ``forward`` then skips ``p4_to_p8(input)``."""
        self.input_p8 = True
        cube.build_batch(frame_idx, cloud_idx, B, None, self.xtrue, self.n_noncloud,
                         xin8=self.X8[0], planes8=self.plan.ep8[0], **kw)

    # ---- forward ------------------------------------------------------------------------------------
    def forward(self, B, save_signs=None):
        P, st = self.plan, stream_ptr()
        L = P.L
        save_signs = self.train if save_signs is None else save_signs
        self.prepare_weights()
        h0, w0 = P.sizes[0]
        if not self.input_p8:                       # input supplied as fp32 P4 (host feed, tests)
            self._call("p4_to_p8(input)", "dincae_p4_to_p8", (
                ptr(self.X[0]), self.x0_planes, P.ep8[0], B, h0, w0, ptr(self.X8[0]), st),
                6 * B * h0 * w0 * P.nvar, 0)
        for l in range(1, L + 1):
            h, w = P.sizes[l - 1]
            name = self._enc_name(l)
            e = self._ent[name + "/kernel"]
            ci, co = P.enc_c[l - 1], P.enc_c[l]
            hh, wh = P.sizes[l]
            nb = 2 * B * (h * w * ci + hh * wh * co) + 18 * ci * co + (B * h * w * co // 8 if save_signs else 0)
            self._call("conv3x3_fwd(enc %d)" % l, "dincae_conv3x3_fwd_bf16", (
                ptr(self.X8[l - 1]), P.ep8[l - 1], 0, None, 0,
                ptr(self._wf(name)), ptr(self._w(name + "/bias")), e["cout_pad"],
                B, h, w, P.ep8[l], NEG_SLOPE, None, ptr(self.X8[l]),
                ptr(self.S8[l]) if save_signs else None, None, 0, st), nb, 18 * ci * co * B * h * w)
        nd = len(P.dense_units)
        hL, wL = P.sizes[L]
        if nd:
            self._call("p8_to_p4(bottleneck)", "dincae_p8_to_p4", (
                ptr(self.X8[L]), P.ep[L], P.ep8[L], B, hL, wL, ptr(self.F[0]), st),
                6 * B * P.n_bottleneck, 0)
        for i in range(nd):
            name = "dense" if i == 0 else "dense_%d" % i
            kk, nn = self._ent[name + "/kernel"]["tf_shape"]
            K, N = P.dense_pad[i], P.dense_pad[i + 1]
            if B <= 256:
                # tensor cores over the tiled bf16 copy: one pass over 2 bytes per weight
                self._call("dense_fwd(%d)" % i, "dincae_dense_tc_fwd", (
                    ptr(self.F[i]), K, ptr(self.w8[i]), ptr(self._w(name + "/bias")), B, K, N, 1,
                    ptr(self.F[i + 1]), N, ptr(self.workspace), self.workspace.numel(), st),
                    4 * (B * kk + B * nn) + 2 * kk * nn, 2 * B * kk * nn)
            else:
                self._call("dense_fwd(%d)" % i, "dincae_dense_fwd", (
                    ptr(self.F[i]), ptr(self._w(name + "/kernel")), ptr(self._w(name + "/bias")),
                    B, K, N, 1, ptr(self.F[i + 1]),
                    ptr(self.workspace), self.workspace.numel(), st),
                    4 * (B * kk + kk * nn + B * nn), 2 * B * kk * nn)
            if save_signs:
                self._dropout_fwd(i, B)
        if nd:
            self._call("p4_to_p8(bottleneck)", "dincae_p4_to_p8", (
                ptr(self.F[nd]), P.ep[L], P.ep8[L], B, hL, wL, ptr(self.D8[0]), st),
                6 * B * P.n_bottleneck, 0)
        for l in range(1, L + 1):
            up, sk, r = P.dec_sources(l)
            h, w = P.sizes[r]
            name = self._dec_name(l)
            e = self._ent[name + "/kernel"]
            hh, wh = P.sizes[r + 1]
            co = P.dec_c[l]
            last = l == L
            nb = 2 * B * (hh * wh * up + h * w * sk) + (4 if last else 2) * B * h * w * co + 18 * (up + sk) * co
            c0p8 = P.dp8[l - 1] if l > 1 else P.ep8[L]
            self._call("conv3x3_fwd(dec %d)" % l, "dincae_conv3x3_fwd_bf16", (
                ptr(self.D8[l - 1]), c0p8, 1,
                ptr(self.X8[r]) if sk else None, P.ep8[r] if sk else 0,
                ptr(self._wf(name)), ptr(self._w(name + "/bias")), e["cout_pad"],
                B, h, w, P.dp8[l], NEG_SLOPE, None if last else ptr(self.D8[l]), None, None,
                ptr(self.xrec) if last else None, 1 if last else 0, st),
                nb, 18 * (up + sk) * co * B * h * w)

    # ---- loss + backward ------------------------------------------------------------------------------
    def loss_backward(self, B, truth_uncertain=False, normalise=True, phase=0):
        if phase == 0:
            self.loss_backward(B, truth_uncertain, normalise, 1)
            self.loss_backward(B, truth_uncertain, normalise, 2)
            return
        P, st = self.plan, stream_ptr()
        L = P.L
        wsp, wsn = ptr(self.workspace), self.workspace.numel()
        wss = ptr(self.workspace_side)
        nd = len(P.dense_units)
        if phase == 2:
            return self._backward_encoder(B, wss, wsn, st)
        self._call("head_loss", "dincae_head_loss_bf16", (
            ptr(self.xrec), ptr(self.xtrue), B, P.H, P.W, int(bool(truth_uncertain)), NEG_SLOPE,
            ptr(self.n_noncloud) if normalise else None, ptr(self.GD8[L]), ptr(self._loss_sums_dst()),
            ptr(self.loss_out), wsp, wsn, st), 40 * B * P.H * P.W, 0)
        for l in range(L, 0, -1):
            up, sk, r = P.dec_sources(l)
            h, w = P.sizes[r]
            name = self._dec_name(l)
            e = self._ent[name + "/kernel"]
            c0p8 = P.dp8[l - 1] if l > 1 else P.ep8[L]
            c1p8 = P.ep8[r] if sk else 0
            hh, wh = P.sizes[r + 1]
            co = P.dec_c[l]
            nb = 2 * B * (hh * wh * up + h * w * (sk + co)) + 36 * (up + sk) * co
            self._fork(lambda: self._call("conv3x3_wgrad(dec %d)" % l, "dincae_conv3x3_wgrad_bf16", (
                ptr(self.D8[l - 1]), c0p8, 1, ptr(self.X8[r]) if sk else None, c1p8,
                ptr(self.GD8[l]), None, None, NEG_SLOPE, P.dp8[l], e["cinp"], e["cout_pad"], B, h, w,
                ptr(self._g(name + "/kernel")), ptr(self._g(name + "/bias")),
                wss, wsn, stream_ptr()), nb, 18 * (up + sk) * co * B * h * w), 1)
            prev_slope = NEG_SLOPE if l > 1 else (0.0 if nd else 1.0)
            prev_act = self.D8[l - 1] if (l > 1 or nd) else None
            want_skip = bool(sk) and r >= 1
            cin_eff = up + (sk if want_skip else 0)
            nb = 2 * B * (h * w * co + 2 * hh * wh * up + (h * w * sk if want_skip else 0)) + 18 * cin_eff * co
            for d in self._wbf[name]["wd"]:
                lo, hi = d["lo"] > 0, d["hi"] > 0
                self._call("conv3x3_dgrad(dec %d)" % l, "dincae_conv3x3_dgrad_bf16", (
                    ptr(self.GD8[l]), None, None, NEG_SLOPE, P.dp8[l], ptr(self.wbf[d["off"]:]), d["nd"],
                    B, h, w,
                    c0p8 if lo else 0, ptr(prev_act) if lo else None, prev_slope if lo else 1.0,
                    ptr(self.GD8[l - 1]) if lo else None,
                    c1p8 if hi else 0, ptr(self.dXskip8[r]) if hi else None, None, st),
                    nb, 18 * cin_eff * co * B * h * w)
        hL, wL = P.sizes[L]
        if nd:
            self._call("p8_to_p4(bottleneck grad)", "dincae_p8_to_p4", (
                ptr(self.GD8[0]), P.ep[L], P.ep8[L], B, hL, wL, ptr(self.dZ[nd]), st),
                6 * B * P.n_bottleneck, 0)
        for i in range(nd - 1, -1, -1):
            name = "dense" if i == 0 else "dense_%d" % i
            K, N = P.dense_pad[i], P.dense_pad[i + 1]
            kk, nn = self._ent[name + "/kernel"]["tf_shape"]
            self._dropout_bwd(i, B)
            self._dense_weight_grad(i, B)
            dst = self.dZ[i] if i > 0 else self.dF0
            if B <= 256:
                self._call("dense_bwd_data(%d)" % i, "dincae_dense_tc_bwd_data", (
                    ptr(self.dZ[i + 1]), N, ptr(self.w8[i]), ptr(self.F[i]) if i > 0 else None, B, K, N,
                    ptr(dst), K, wsp, wsn, st), 4 * (B * nn + B * kk) + 2 * kk * nn, 2 * B * kk * nn)
            else:
                self._call("dense_bwd_data(%d)" % i, "dincae_dense_bwd_data", (
                    ptr(self.dZ[i + 1]), ptr(self._w(name + "/kernel")),
                    ptr(self.F[i]) if i > 0 else None, B, K, N, ptr(dst), wsp, wsn, st),
                    4 * (B * nn + kk * nn + B * kk), 2 * B * kk * nn)
        if nd:
            self._call("p4_to_p8(bottleneck grad)", "dincae_p4_to_p8", (
                ptr(self.dF0), P.ep[L], P.ep8[L], B, hL, wL, ptr(self.dX8[L]), st),
                6 * B * P.n_bottleneck, 0)
        self._join()

    def _backward_encoder(self, B, wss, wsn, st):
        P, L = self.plan, self.plan.L
        for l in range(L, 0, -1):
            h, w = P.sizes[l - 1]
            name = self._enc_name(l)
            e = self._ent[name + "/kernel"]
            gpool = self.dX8[l]
            ci, co = P.enc_c[l - 1], P.enc_c[l]
            hh, wh = P.sizes[l]
            nb = 2 * B * (h * w * ci + hh * wh * co) + B * h * w * co // 8 + 36 * ci * co
            self._fork(lambda: self._call("conv3x3_wgrad(enc %d)" % l, "dincae_conv3x3_wgrad_bf16", (
                ptr(self.X8[l - 1]), P.ep8[l - 1], 0, None, 0,
                None, ptr(gpool), ptr(self.S8[l]), NEG_SLOPE, P.ep8[l], e["cinp"], e["cout_pad"], B, h, w,
                ptr(self._g(name + "/kernel")), ptr(self._g(name + "/bias")),
                wss, wsn, stream_ptr()), nb, 18 * ci * co * B * h * w), 4)
            if l > 1:
                has_add = self.dXskip8[l - 1] is not None
                d = self._wbf[name]["wd"][0]
                nb = 2 * B * (hh * wh * co + h * w * ci * (2 if has_add else 1)) + B * h * w * co // 8 + 18 * ci * co
                self._call("conv3x3_dgrad(enc %d)" % l, "dincae_conv3x3_dgrad_bf16", (
                    None, ptr(gpool), ptr(self.S8[l]), NEG_SLOPE, P.ep8[l], ptr(self.wbf[d["off"]:]), d["nd"],
                    B, h, w, 0, None, 1.0, None,
                    P.ep8[l - 1], ptr(self.dX8[l - 1]),
                    ptr(self.dXskip8[l - 1]) if has_add else None, st),
                    nb, 18 * ci * co * B * h * w)
        self._join()
