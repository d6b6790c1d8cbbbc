"""Training / reconstruction driver on top of engine.Network and data.DeviceCube.

One ``Trainer.train_step`` = one ``sess.run([cost, RMS, opt])`` of the reference
(DINCAE/__init__.py:649-654): batch assembly, forward, loss, backward, global-norm clip and
Adam -- all on the device, captured once into a CUDA graph and replayed.  Per step the host
sends 16 bytes per frame of indices and reads back 40 bytes of loss sums.

Data parallelism (new capability): one process per GPU; every rank
draws the same global index stream and builds its own slice; gradients and the loss sums
are left un-normalised, summed with an NCCL all-reduce, and divided by the global count of
observed pixels inside the Adam kernel, so N ranks x B frames is numerically a single
batch of N*B.
"""
import os

import numpy as np
import torch

from .engine import Network, make_network


class Trainer:
    def __init__(self, plan, cube, batch_size, truth_uncertain=False, clip_grad=5.0,
                 l2_beta=0.0, noise_seed=0, use_graph=True, dist_group=None, tf_params=None,
                 init_seed=0, jitter_std=None, dropout_rate=0.0):
        self.plan, self.cube = plan, cube
        self.B = int(batch_size)
        self.truth_uncertain = bool(truth_uncertain)
        self.clip_grad = float(clip_grad)
        self.l2_beta = float(l2_beta)
        self.noise_seed = int(noise_seed)
        self.jitter_std = jitter_std
        import os
        self.dist = None
        self.rank, self.world = 0, 1
        if dist_group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.group = dist_group if dist_group is not True else None
            self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        # Factored dense gradients (csrc/dense_factored.cu): default whenever there is a dense
        # bottleneck and no L2 term.  Data parallel, the ranks then exchange ONE all-gather of
        # their gradient blobs (factors + conv / bias gradients + loss sums) per step.
        n_dense = sum(e["size"] for e in plan.entries if e["kind"] == "dense_w")
        want = os.environ.get("DINCAE_FACTORED", "auto")
        # worth it when the dense kernels are big enough for their gradient traffic to matter
        # (configs[3]: 688 M parameters, 2.75 GB to all-reduce otherwise); at config A (2.8 M) the
        # step is latency-bound and the materialised path is a dozen small launches shorter
        # (measured on 2 GPUs: 0.872 ms factored vs the single all-reduce below)
        auto = n_dense >= 32 * 1024 * 1024
        factored = (len(plan.dense_units) > 0 and self.l2_beta == 0.0 and self.world * self.B <= 1024
                    and (want == "1" or (want == "auto" and auto)))
        self.net = make_network(plan, self.B, train=True, device=cube.dev, factored=factored,
                                world=self.world, rank=self.rank)
        self.net.set_params_tf(tf_params if tf_params is not None else plan.glorot_init(init_seed))
        if self.dist is not None:
            self.dist.broadcast(self.net.params, src=0, group=self.group)
            self.net.params_changed()
        self._nccl_in_graph = os.environ.get("DINCAE_NCCL_IN_GRAPH", "1") == "1"
        dev = cube.dev
        B = self.B
        # [frame(B) | cloud(B) | seq int64 (2B int32)]
        # pinned staging ring for the per-step indices: the host runs ahead of the stream, so a
        # slot may only be rewritten once the asynchronous copy that read it has executed
        self._idx_ring = [torch.zeros(4 * B, dtype=torch.int32).pin_memory() for _ in range(8)]
        self._idx_events = [None] * len(self._idx_ring)
        self._idx_pos = 0
        self.idx_dev = torch.zeros(4 * B, dtype=torch.int32, device=dev)
        self.frame_idx = self.idx_dev[:B]
        self.cloud_idx = self.idx_dev[B:2 * B]
        self.noise_seq = self.idx_dev[2 * B:].view(torch.int64)
        # optional dropout after the dense layers (inert in the reference: default 0), keyed by the
        # frames' global sequence numbers so that the masks do not depend on the GPU count
        self.net.dropout_rate = float(dropout_rate)
        self.net.dropout_seed = self.noise_seed + 0x5DEECE66D
        self.net.dropout_keys = self.noise_seq
        self.use_graph = use_graph
        self._graphs = {}
        self._ring = torch.zeros(1024, 4, dtype=torch.float64).pin_memory()
        self._ring_l2 = torch.zeros(1024, 1, dtype=torch.float32).pin_memory()
        self._pending = []          # ring slots not yet converted to floats
        self._stash = []            # entries drained because the ring filled up between two flushes
        self._ring_pos = 0
        self.step_count = 0
        self.launches_per_step = None

    # ---- one optimisation step ---------------------------------------------------------------
    def _enqueue_fwd_bwd(self, phases=(1, 2)):
        net, B = self.net, self.B
        if 1 in phases:
            net.n_noncloud.zero_()
            net.build_input(self.cube, self.frame_idx, self.cloud_idx, B, jitter_std=self.jitter_std,
                            seed=self.noise_seed, noise_seq=self.noise_seq)
            net.forward(B)
            net.loss_backward(B, self.truth_uncertain, normalise=False, phase=1)
        if 2 in phases:
            net.loss_backward(B, self.truth_uncertain, normalise=False, phase=2)
            if self.l2_beta != 0.0:
                net.l2_value()

    def _enqueue_update(self):
        net = self.net
        net.adam_step(clip_norm=self.clip_grad, l2_beta=self.l2_beta, denom=net.loss_sums[3:4], B=self.B)

    def _run(self, key, fn):
        if not self.use_graph:
            fn()
            return
        g = self._graphs.get(key)
        if g is None:
            # warm up on a side stream (required before capture), then capture
            s = torch.cuda.Stream(device=self.cube.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._snapshot()
                fn()
                self._restore()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: CUDA calls of other host threads (writer threads waiting on events)
            # must not invalidate this capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fn()
            self._graphs[key] = g
        g.replay()

    def _snapshot(self):
        n = self.net
        self._snap = [t.clone() for t in (n.params, n.adam_m, n.adam_v, n.adam_state)]

    def _restore(self):
        n = self.net
        for t, s in zip((n.params, n.adam_m, n.adam_v, n.adam_state), self._snap):
            t.copy_(s)
        n.params_changed()          # the warm-up step also advanced the bf16 dense copies
        self._snap = None

    def set_lr(self, lr):
        self.net.lr.fill_(float(lr))

    def upload_indices(self, frame_idx, cloud_idx, seq):
        B = self.B
        slot = self._idx_pos % len(self._idx_ring)
        self._idx_pos += 1
        if self._idx_events[slot] is not None:
            self._idx_events[slot].synchronize()
        h = self._idx_ring[slot]
        h[:B] = torch.from_numpy(np.ascontiguousarray(frame_idx, dtype=np.int32))
        h[B:2 * B] = torch.from_numpy(np.ascontiguousarray(cloud_idx, dtype=np.int32))
        h[2 * B:].view(torch.int64)[:] = torch.from_numpy(np.ascontiguousarray(seq, dtype=np.int64))
        self.idx_dev.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._idx_events[slot] = ev

    def train_step(self, frame_idx, cloud_idx, seq):
        """Enqueue one step for this rank's B frames; the loss is fetched lazily."""
        self.upload_indices(frame_idx, cloud_idx, seq)
        self.run_step()

    def run_step(self):
        """One step on the indices already in ``self.idx_dev``."""
        if self.dist is None:
            self._run("step", lambda: (self._enqueue_fwd_bwd(), self._enqueue_update()))
        elif self.net.factored:
            # one exchange per step: every rank's blob (dense factors, the other gradients, loss
            # sums) is all-gathered in place; the optimiser kernels then sum the ranks' copies in
            # rank order, so all ranks apply bit-identical updates
            net = self.net
            gather = lambda: self.dist.all_gather_into_tensor(net.blob_all.view(-1), net.blob, group=self.group)
            if self._nccl_in_graph:
                self._run("step", lambda: (self._enqueue_fwd_bwd(), gather(), self._enqueue_update()))
            else:
                self._run("fwd_bwd", self._enqueue_fwd_bwd)
                gather()
                self._run("update", self._enqueue_update)
        else:
            # materialised gradients: ONE all-reduce of [all gradients | loss sums as (hi, lo) float
            # pairs] after the backward pass (round 1 issued three collectives with host-side waits
            # in between, pure latency at 11.5 MB), captured with the kernels into one graph when
            # NCCL capture is enabled
            net = self.net
            n_all = net.plan.n_params_flat + 8
            (b0, b1), _ = net.grad_buckets()
            split = os.environ.get("DINCAE_DP_BUCKETS", "2") == "2" and b1 < n_all

            def fwd_bwd_reduce():
                # the big bucket (dense + decoder kernels, 97 % of the bytes) is final after phase 1:
                # its all-reduce runs next to the encoder backward; the rest (encoder kernels, biases,
                # loss sums) follows as one small collective
                if split:
                    self._enqueue_fwd_bwd((1,))
                    h = self.dist.all_reduce(net.grads_ext[b0:b1], group=self.group, async_op=True)
                    self._enqueue_fwd_bwd((2,))
                    net.pack_loss_sums()
                    self.dist.all_reduce(net.grads_ext[b1:n_all], group=self.group)
                    h.wait()
                else:
                    self._enqueue_fwd_bwd()
                    net.pack_loss_sums()
                    self.dist.all_reduce(net.grads_ext[:n_all], group=self.group)

            if self._nccl_in_graph:
                self._run("step", lambda: (fwd_bwd_reduce(), net.unpack_loss_sums(), self._enqueue_update()))
            else:
                self._run("fwd_bwd", lambda: (self._enqueue_fwd_bwd(), net.pack_loss_sums()))
                self.dist.all_reduce(net.grads_ext[:n_all], group=self.group)
                self._run("update", lambda: (net.unpack_loss_sums(), self._enqueue_update()))
        # loss bookkeeping: 4 sums + L2 value -> pinned ring (async)
        slot = self._ring_pos % self._ring.shape[0]
        if len(self._pending) >= self._ring.shape[0] - 1:
            # ring full in the middle of an epoch: keep the drained entries for the next
            # flush_losses() of the caller instead of dropping them
            self._stash += self._drain()
        self._ring[slot].copy_(self.net.loss_sums, non_blocking=True)
        if self.l2_beta != 0.0:
            self._ring_l2[slot].copy_(self.net.l2_out, non_blocking=True)
        self._pending.append(slot)
        self._ring_pos += 1
        self.step_count += 1

    def flush_losses(self):
        """Synchronise and return [(cost, RMS), ...] of the steps since the last flush."""
        out = self._stash + self._drain()
        self._stash = []
        return out

    def _drain(self):
        if not self._pending:
            return []
        torch.cuda.current_stream().synchronize()
        out = []
        for slot in self._pending:
            s = self._ring[slot].numpy()
            n = s[3]
            cost = (s[0] + s[1]) / n if n != 0 else float("nan")
            if self.l2_beta != 0.0:
                cost = cost + self.l2_beta * float(self._ring_l2[slot, 0])
            rms = float(np.sqrt(s[2] / n)) if n != 0 else float("nan")
            out.append((float(np.float32(cost)), float(np.float32(rms))))
        self._pending = []
        return out


class Reconstructor:
    """Forward-only sweep (DINCAE/__init__.py:678-689): test-mode batches, head, D2H."""

    def __init__(self, plan, cube, batch_size, net=None, train_mode_inputs=False, noise_seed=0,
                 use_graph=True, py_random=None, jitter_std=None, records=None, n_out_slots=2):
        """``records=(meandata, fill_value)``: ``run_batch`` returns file-ready records
        (big-endian float32 [B][2][H][W]: mean_rec = m_rec + meandata, sigma_rec, fill where
        ``meandata`` is masked) instead of (m_rec, sigma2_rec) -- the per-batch arithmetic of
        ``savesample`` runs on the device (``dincae_head_recon_records``)."""
        self.plan, self.cube, self.B = plan, cube, int(batch_size)
        self.net = net if net is not None else make_network(plan, self.B, train=False, device=cube.dev)
        self.train_mode_inputs = train_mode_inputs       # reference quirk Q1
        self.noise_seed = noise_seed
        self.jitter_std = jitter_std
        self.use_graph = use_graph
        self.py_random = py_random
        dev = cube.dev
        B = self.B
        self._idx_ring = [torch.zeros(4 * B, dtype=torch.int32).pin_memory() for _ in range(8)]
        self._idx_events = [None] * len(self._idx_ring)
        self._idx_pos = 0
        self.idx_dev = torch.zeros(4 * B, dtype=torch.int32, device=dev)
        self.records = records is not None
        if self.records:
            md = np.ma.asarray(records[0])
            self._fill = float(records[1])
            self._mean_is_f32 = md.dtype == np.float32
            self._mean_dev = torch.from_numpy(
                np.ascontiguousarray(np.ma.filled(md, 0).reshape(plan.H, plan.W), dtype=np.float64)).to(dev)
            self._never_dev = torch.from_numpy(
                np.ascontiguousarray(np.ma.getmaskarray(md).reshape(plan.H, plan.W)).view(np.uint8)).to(dev)
            self.out_host = [torch.zeros(B, 2, plan.H, plan.W, dtype=torch.int32).pin_memory()
                             for _ in range(n_out_slots)]
            self.out_dev = torch.zeros(B, 2, plan.H, plan.W, dtype=torch.int32, device=dev)
        else:
            self.out_host = [torch.zeros(2, B, plan.H, plan.W).pin_memory() for _ in range(n_out_slots)]
            self.out_dev = torch.zeros(2, B, plan.H, plan.W, device=dev)
        self._graphs = {}
        self._sweep = -1           # index of the current train-mode sweep (keys the draws)
        self._cloud = None

    def begin_sweep(self):
        """Start a new pass over the frames (every rank of a sharded sweep calls this once per
        pass): new added-cloud draws and noise keys for train-mode inputs."""
        self._sweep += 1
        if self.train_mode_inputs:
            seed = (int(self.noise_seed) * 1000003 + self._sweep) % (2 ** 32)
            self._cloud = np.random.RandomState(seed).randint(0, self.cube.T, size=self.cube.T).astype(np.int32)

    def record_event(self):
        """Event after everything enqueued so far (the D2H copy of the last ``run_batch``)."""
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def _enqueue(self, n):
        net, B = self.net, self.B
        net.n_noncloud.zero_()
        cloud = self.idx_dev[B:2 * B] if self.train_mode_inputs else None
        net.build_input(self.cube, self.idx_dev[:B], cloud, n, jitter_std=self.jitter_std,
                        seed=self.noise_seed, noise_seq=self.idx_dev[2 * B:].view(torch.int64))
        net.forward(n, save_signs=False)
        if self.records:
            net.head_recon_records(n, self._mean_dev, self._never_dev, self._fill, self._mean_is_f32,
                                   self.out_dev)
            return
        net.head_recon(n)
        self.out_dev[0, :n].copy_(net.m_rec[:n])
        self.out_dev[1, :n].copy_(net.s2_rec[:n])

    def capture(self, n):
        """CUDA graph of the sweep of a batch of ``n`` frames (captured on first use; callers that
        run host threads next to the sweep capture the sizes they need up front)."""
        g = self._graphs.get(n)
        if g is None:
            s = torch.cuda.Stream(device=self.cube.dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue(n)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._enqueue(n)
            self._graphs[n] = g
        return g

    def run_batch(self, frames, slot=0):
        """frames: int array (n <= B).  Returns pinned host tensor [2,B,H,W] (valid after a
        stream sync): [0] = m_rec, [1] = sigma2_rec -- or, in records mode, the int32 view of
        the big-endian float32 records [B,2,H,W]."""
        n, B = len(frames), self.B
        islot = self._idx_pos % len(self._idx_ring)
        self._idx_pos += 1
        if self._idx_events[islot] is not None:
            self._idx_events[islot].synchronize()      # the copy that read this slot has run
        h = self._idx_ring[islot]
        h[:n] = torch.from_numpy(np.ascontiguousarray(frames, dtype=np.int32))
        if self.train_mode_inputs:
            # Reference quirk Q1: the "test" generator of reconstruct_gridded_nc adds clouds and
            # jitter.  The cloud index and the noise key of a frame depend only on (seed, sweep
            # number, ABSOLUTE frame number): every rank of a sharded sweep draws the same table, the
            # result does not depend on the world size, and the global `random` is never touched.
            T = self.cube.T
            if self._cloud is None:
                self.begin_sweep()
            fr = np.asarray(frames, dtype=np.int64)
            h[B:B + n] = torch.from_numpy(self._cloud[fr])
            h[2 * B:].view(torch.int64)[:n] = torch.from_numpy(fr + self._sweep * T)
        self.idx_dev.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._idx_events[islot] = ev
        if self.use_graph:
            self.capture(n).replay()
        else:
            self._enqueue(n)
        out = self.out_host[slot]
        if self.records:
            out[:n].copy_(self.out_dev[:n], non_blocking=True)
        else:
            out.copy_(self.out_dev, non_blocking=True)
        return out
