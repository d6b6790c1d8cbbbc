#!/usr/bin/env python3
"""Benchmark of the DINCAE hot path (BASELINE.json metric: train frames/s at 112x112, plus
reconstruction frames/s) on synthetic AVHRR-like cubes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one optimisation step (batch assembly, forward, loss, backward, clip + Adam)
on one batch of 50 frames of 112x112 per GPU -- BASELINE.json configs[1] (5266-frame cube,
default architecture [16,24,36,54]); with N > 1 every rank processes its own 50 frames
(weak scaling, gradients summed with an NCCL all-reduce).  Prints ONE JSON line.

* value     : frames/s with the cube resident in HBM and the step's indices on the device
              (CUDA events, barrier + synchronize on both sides, max over ranks).
* e2e       : frames/s through the public host-facing path: the cube is uploaded from pinned
              host memory INSIDE the timed region, every step copies its indices host->device
              and its loss sums device->host.
* roofline  : dominant kernel call of the step, timed live with CUDA events in an eager
              (non-graph) pass; achieved = algorithmic bytes of that call / its duration.
* cpu_baseline / --impl reference : the oracle (numpy generator port + torch-CPU graph) on the
              box's host cores; the reference itself (TF 1.15) cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


METRIC = "train frames/sec at 112x112 (value) & 512x512 (config_d); recon frames/sec alongside"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ntime", type=int, default=5266)
    ap.add_argument("--size", type=int, default=112)
    ap.add_argument("--batch", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--recon-shard", action="store_true",
                    help="(kept for compatibility: the sharded sweep is timed by default at N > 1)")
    ap.add_argument("--no-recon-shard", action="store_true")
    ap.add_argument("--no-config-d", action="store_true", help="skip the 512x512 bf16 leg (configs[3])")
    ap.add_argument("--batch-d", type=int, default=16, help="frames per GPU and step of the 512x512 leg")
    ap.add_argument("--steps-d", type=int, default=0, help="timed steps of the 512x512 leg (default min(steps, 20))")
    ap.add_argument("--no-e2e-api", action="store_true",
                    help="time e2e through Trainer.train_step instead of DINCAE.reconstruct_gridded_nc")
    ap.add_argument("--e2e-epochs", type=int, default=100,
                    help="epochs of the e2e run through DINCAE.reconstruct_gridded_nc (reference default: 1000)")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {"workload": "configs[1]: AVHRR-like synthetic %dx%dx%d SST cube, default architecture "
                        "[16,24,36,54], training step" % (a.size, a.size, a.ntime),
            "frames_per_gpu_per_step": a.batch, "global_batch": a.batch * n_gpus,
            "ntime": a.ntime, "height": a.size, "width": a.size, "nvar": 10,
            "parallelism": "dp%d" % n_gpus,
            "l2_policy": "working set per step (activations ~0.45 GB + cube 0.33 GB) exceeds the "
                         "126 MB L2; no explicit flush"}


def make_cube(a, pinned):
    from dincae_b200 import synthetic
    lon, lat, time_, data, mask = synthetic.sst_like_cube(T=a.ntime, H=a.size, W=a.size, seed=12345)
    if pinned:
        # both halves of the masked array (values and mask) live in pinned host memory
        raw = torch.empty(data.shape, dtype=torch.float32).pin_memory()
        raw.copy_(torch.from_numpy(np.ma.getdata(data)))
        mk = torch.empty(data.shape, dtype=torch.bool).pin_memory()
        mk.copy_(torch.from_numpy(np.ma.getmaskarray(data)))
        data = np.ma.masked_array(raw.numpy(), mk.numpy())
        data._pinned_owner = (raw, mk)
    return lon, lat, time_, data, mask


class ClockSampler:
    """SM clock / throttle-reason samples taken IN PROCESS through NVML by a polling thread
    (every ~2 ms) from construction to ``stop()`` -- warm-up plus the timed region.  (A
    ``nvidia-smi -lms 100`` child saw no sample inside a 14 ms timed region.)"""

    def __init__(self, gpu_index, period=0.002):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.err = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:            # noqa: BLE001
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for nm, bit in names.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception as e:            # noqa: BLE001
                        self.err = repr(e)
                        return
                    self._stop.wait(period)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception as e:            # noqa: BLE001
            self.err = repr(e)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
               "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples), "source": "NVML, in-process polling thread"}
        if self.err:
            out["error"] = self.err
        return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy port of the reference generator + torch-CPU restatement of the
# TF graph) driven like the reference: python generator -> repeat -> shuffle(45) -> batch.
# ------------------------------------------------------------------------------------------
def cpu_arm(a, steps, warmup, ntime_cap=400):
    from oracle import datagen_np
    from oracle import model_torch as M
    import random
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = argparse.Namespace(**vars(a))
    b.ntime = min(a.ntime, ntime_cap)              # bounded sample of the cube (host RAM/time)
    lon, lat, time_, data, mask = make_cube(b, pinned=False)
    missing = np.ma.getmaskarray(data)
    random.seed(1)
    np.random.seed(1)
    gen, nvar, ntime, meandata = datagen_np.data_generator(lon, lat, time_, data, missing)
    arch = M.Arch(a.size, a.size, nvar)
    params = M.init_params(arch, 0)
    opt = M.TFAdam(params)

    def repeat():
        while True:
            for item in gen():
                yield item

    stream, buf, rs = repeat(), [], np.random.RandomState(0)

    def next_batch():
        xs, ts = [], []
        for _ in range(a.batch):
            while len(buf) < 45:
                buf.append(next(stream))
            j = rs.randint(0, len(buf))
            xs.append(buf[j][0])
            ts.append(buf[j][1])
            buf[j] = next(stream)
        return torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ts))

    t_gen = t_model = 0.0
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        xin, xtrue = next_batch()
        t1 = time.perf_counter()
        M.train_step(arch, params, opt, xin, xtrue, 1e-3)
        t2 = time.perf_counter()
        if k >= warmup:
            t_gen += t1 - t0
            t_model += t2 - t1
    # reconstruction: forward + head on test-mode frames
    feat = datagen_np.StaticFeatures(lon, lat, time_, [data], [1.0])
    t0 = time.perf_counter()
    nrec = 2 * a.batch
    with torch.no_grad():
        for s in range(0, nrec, a.batch):
            x = torch.from_numpy(np.stack([datagen_np.stack_frame(feat, i % ntime, 3)
                                           for i in range(s, s + a.batch)]))
            M.head(M.forward(arch, params, x))
    t_rec = time.perf_counter() - t0
    total = t_gen + t_model
    return {"value": a.batch * steps / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d train steps of %d frames (%dx%d) drawn from the first %d frames of the "
                      "cube: reference generator port (single python thread) + torch-CPU fp32 graph "
                      "on %d threads" % (steps, a.batch, a.size, a.size, b.ntime, cores),
            "generator_only_fps": a.batch * steps / t_gen, "model_only_fps": a.batch * steps / t_model,
            "recon_fps": nrec / t_rec, "seconds": total}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_arm(a, a.steps, a.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * a.batch / res["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, a.gpus),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "recon_fps": res["recon_fps"], "generator_only_fps": res["generator_only_fps"],
            "model_only_fps": res["model_only_fps"],
            "note": "TF 1.15 (setup.py:19) is not installable here; this arm times the oracle "
                    "restatement of the reference's CPU path (kind=port)"}
    emit(line)


def peak_bf16():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 2250.0, "fallback (nominal dense bf16)"


def traced_step(trainer, B, jitter, reps):
    """Per-call CUDA-event times of one optimisation step, eager, side stream off (two kernels
    sharing the SMs would read 2-4x longer each).  Returns {label: [ms, alg_bytes, alg_flops]}."""
    from dincae_b200 import _lib
    records = []

    def tracer(label, fn, args, nbytes, nflops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(*args), label)
        e1.record()
        records.append((label, e0, e1, nbytes, nflops))

    net = trainer.net
    side, net.side = net.side, None
    net.tracer = tracer
    ebb0, ebb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bb_ms = 0.0
    for _ in range(reps):
        net.n_noncloud.zero_()
        # keep the GPU busy while the host enqueues the step, so that the kernels run back to
        # back and the events measure execution, not the python launch gaps
        torch.cuda._sleep(8_000_000)
        ebb0.record()
        net.build_input(trainer.cube, trainer.frame_idx, trainer.cloud_idx, B, jitter_std=jitter, seed=1,
                        noise_seq=trainer.noise_seq)
        ebb1.record()
        net.forward(B)
        net.loss_backward(B, False, normalise=False)
        net.adam_step(denom=net.loss_sums[3:4])
        torch.cuda.synchronize()
        bb_ms += ebb0.elapsed_time(ebb1)
    net.tracer = None
    net.side = side
    agg = {}
    for label, e0, e1, nb, nf in records:
        d = agg.setdefault(label, [0.0, nb, nf])
        d[0] += e0.elapsed_time(e1) / reps
    cube = trainer.cube
    agg["build_batch"] = [bb_ms / reps, B * cube.H * cube.W * (3 * 5 * cube.ndata + 1 + 4 * cube.nvar + 8), 0]
    return agg


def config_d_leg(a, world, rank, dist, barrier):
    """BASELINE configs[3]: 512x512 multivariate cube (SST + chlorophyll + u/v wind: nvar 28),
    encoder [32,48,72,108,162], bf16 convolutions (tcgen05 kind::f16), fp32 master weights /
    dense bottleneck / loss / Adam.  One training step per iteration, `--batch-d` frames per GPU."""
    from dincae_b200 import synthetic
    from dincae_b200.data import DeviceCube
    from dincae_b200.engine import Plan
    from dincae_b200.sampler import TrainSampler
    from dincae_b200.trainer import Trainer
    import random
    B = a.batch_d
    K = a.steps_d if a.steps_d > 0 else min(a.steps, 20)
    T, H, W = max(24, B + 8), 512, 512
    fields, missing = [], []
    for kind, seed in (("sst", 100), ("chl", 101), ("wind", 102), ("wind", 103)):
        lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=W, seed=seed, kind=kind)
        fields.append(data)
        missing.append(np.ma.getmaskarray(data))
    cube = DeviceCube(lon, lat, time_, fields, missing, [1.0] * 4, 3)
    filt = (32, 48, 72, 108, 162)
    plan = Plan(H, W, cube.nvar, filt, (0.2,), (1, 2, 3, 4, 5), lanes=8)
    jitter = [0.05] * 4
    trainer = Trainer(plan, cube, B, noise_seed=4321, jitter_std=jitter, use_graph=not a.no_graph,
                      dist_group=True if world > 1 else None)
    trainer.set_lr(1e-3)
    random.seed(5)
    sampler = TrainSampler(T, B * world, 8, seed=11)

    def next_slice():
        f, c, q = sampler.next_batch()
        sl = slice(rank * B, (rank + 1) * B)
        return f[sl], c[sl], q[sl]

    for _ in range(3):
        trainer.train_step(*next_slice())
    trainer.flush_losses()
    staged = torch.zeros(K, 4 * B, dtype=torch.int32)
    for k in range(K):
        f, c, q = next_slice()
        staged[k, :B] = torch.from_numpy(f)
        staged[k, B:2 * B] = torch.from_numpy(c)
        staged[k, 2 * B:].view(torch.int64)[:] = torch.from_numpy(q)
    staged = staged.cuda()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(K):
        trainer.idx_dev.copy_(staged[k])
        trainer.run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    losses = trainer.flush_losses()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out = {"workload": "configs[3]: synthetic 512x512 cube, 4 variables (SST, chlorophyll, u-wind, v-wind; "
                       "nvar 28), encoder [32,48,72,108,162], 688.7 M parameters, training step",
           "dtype": "bf16 (convolution operands and stored activations; fp32 accumulation, master "
                    "weights, dense bottleneck, loss, Adam)",
           "n_gpus": world, "frames_per_gpu_per_step": B, "global_batch": B * world, "steps": K,
           "warmup": 3, "ms_per_step": ms / K, "value": world * B * K / (ms / 1e3), "unit": UNIT,
           "scaling": "weak", "loss_first_last": [losses[0][0], losses[-1][0]] if losses else None,
           "l2_policy": "parameters + Adam state (11 GB) and activations exceed the 126 MB L2; no explicit flush"}
    if rank == 0:
        agg = traced_step(trainer, B, jitter, 3)
        total_ms = sum(v[0] for v in agg.values())
        conv = {k: v for k, v in agg.items() if k.startswith("conv3x3")}
        conv_ms = sum(v[0] for v in conv.values())
        conv_fl = sum(v[2] for v in conv.values())
        pk_tf, pk_tf_src = peak_bf16()
        pk_bw, pk_bw_src = peaks()
        top = max(agg.items(), key=lambda kv: kv[1][0])
        step_bytes = sum(v[1] for v in agg.values())
        out["roofline"] = {
            "bound": "hbm", "kernel": top[0], "kernel_ms": top[1][0],
            "achieved": top[1][1] / (top[1][0] * 1e-3) / 1e9, "peak": pk_bw, "unit": "GB/s",
            "frac": top[1][1] / (top[1][0] * 1e-3) / 1e9 / pk_bw, "peak_source": pk_bw_src,
            "algorithmic_bytes": top[1][1], "kernel_share_of_step": top[1][0] / total_ms,
            "step_algorithmic_bytes": step_bytes,
            "step_hbm_frac": step_bytes / (ms / K * 1e-3) / 1e9 / pk_bw,
            "conv_ms": conv_ms, "conv_algorithmic_tflops": conv_fl / (conv_ms * 1e-3) / 1e12,
            "conv_tensor_frac": conv_fl / (conv_ms * 1e-3) / 1e12 / pk_tf, "tensor_peak_tflops": pk_tf,
            "tensor_peak_source": pk_tf_src,
            "tensor_pipe_ncu": "profiles/r2_configD_ncu.md (sm__pipe_tensor_cycles_active per kernel)"}
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        t = json.load(open(tp)).get(top[0]) if os.path.isfile(tp) else None
        out["roofline"]["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"] if t else None
        out["kernel_breakdown_ms"] = {k: round(v[0], 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        out["conv_tflops_by_kernel"] = {k: round(v[2] / (v[0] * 1e-3) / 1e12, 1)
                                        for k, v in sorted(conv.items(), key=lambda kv: -kv[1][0])}
    del trainer
    torch.cuda.empty_cache()
    return out


def e2e_api_leg(a, world, rank, dist, barrier, K):
    """The metric through the reference-facing call on HOST data: rank 0 writes the synthetic
    cube to a netCDF file (untimed), then every rank calls
    DINCAE.reconstruct_gridded_nc(file, "SST", outdir, epochs=E, batch_size=B, save_each=0)
    -- file read, cube upload, E x ceil(T / (N B)) optimisation steps with `loss.append`, the
    final reconstruction sweep and the output file are all inside the timed region."""
    import contextlib
    import glob
    import tempfile
    from math import ceil
    import DINCAE
    from dincae_b200 import synthetic
    box = [None]
    if rank == 0:
        d = tempfile.mkdtemp(prefix="dincae_bench_")
        lon, lat, time_, data, mask = synthetic.sst_like_cube(T=a.ntime, H=a.size, W=a.size, seed=12345)
        synthetic.write_input_file(os.path.join(d, "sst.nc"), "SST", lon, lat, time_, data, mask)
        box[0] = d
    if dist is not None:
        dist.broadcast_object_list(box, src=0)
    d = box[0]
    fname, outdir = os.path.join(d, "sst.nc"), os.path.join(d, "out")
    spe = ceil(a.ntime / (a.batch * world))
    # The reference trains epochs=1000 by default (DINCAE/__init__.py:307); 100 keeps the bench
    # short.  Every one-off cost of the call (netCDF read, upload, kernel / graph set-up, the final
    # reconstruction and its file) is inside the timed region and is amortised over these epochs.
    # Weak scaling like the device-timed leg: the epoch count grows with the number of ranks, so every
    # rank runs (about) the same number of optimisation steps at every N and the one-off costs are
    # amortised over the same per-rank work.
    E = a.e2e_epochs * world if a.e2e_epochs > 0 else max(1, ceil(K / spe))
    loss = []
    barrier()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        DINCAE.reconstruct_gridded_nc(fname, "SST", outdir, epochs=E, batch_size=a.batch, save_each=0,
                                      save_model_each=0, iseed=12345, loss=loss)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    files = glob.glob(os.path.join(outdir, "data-*.nc"))
    in_bytes = os.path.getsize(fname)
    out_bytes = os.path.getsize(files[0]) if files else 0
    steps = E * spe
    res = {"value": steps * a.batch * world / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(a.ntime * a.size * a.size * 5 / steps + 16 * a.batch),
           "d2h_bytes_per_step": int(32 + 2 * 4 * a.ntime * a.size * a.size / steps),
           "seconds": dt, "epochs": E, "steps": steps, "train_frames": steps * a.batch * world,
           "reconstructed_frames": a.ntime, "input_file_bytes": in_bytes, "output_file_bytes": out_bytes,
           "final_loss": loss[-1] if loss else None,
           "note": "timed region = DINCAE.reconstruct_gridded_nc(file, 'SST', outdir, epochs=%d, batch_size=%d, "
                   "save_each=0): netCDF read of the %d-frame cube, upload, %d optimisation steps "
                   "(loss.append each), reconstruction of all %d frames and the output netCDF file; "
                   "value = trained frames / wall seconds of the whole call; epochs = %d x n_gpus (weak "
                   "scaling: the same number of steps per rank at every N)" % (E, a.batch, a.ntime, steps, a.ntime,
                                                                              a.e2e_epochs)}
    if rank == 0:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    return res


# ------------------------------------------------------------------------------------------
def run_ours(a):
    from dincae_b200 import _lib
    from dincae_b200.data import DeviceCube
    from dincae_b200.engine import Plan
    from dincae_b200.sampler import TrainSampler, sequential_batches
    from dincae_b200.trainer import Reconstructor, Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # keep stdout to the one JSON line (some NCCL builds print their version banner there)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    B, K, W = a.batch, a.steps, max(3, a.warmup)

    # ---- CPU baseline first, at N = 1 only (at N > 1 the other ranks would spin in NCCL
    #      barriers on the host cores the baseline needs) ------------------------------------
    cpu_base = None
    if world == 1 and not a.no_cpu_baseline:
        cb = cpu_arm(a, a.cpu_steps, 1)
        cpu_base = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu_base["recon_fps"] = cb["recon_fps"]
        cpu_base["generator_only_fps"] = cb["generator_only_fps"]

    lon, lat, time_, data, mask = make_cube(a, pinned=True)
    missing = np.ma.getmaskarray(data)
    cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
    plan = Plan(a.size, a.size, cube.nvar)
    trainer = Trainer(plan, cube, B, noise_seed=1234, jitter_std=[0.05],
                      use_graph=not a.no_graph, dist_group=True if world > 1 else None)
    trainer.set_lr(1e-3)
    import random
    random.seed(99)
    sampler = TrainSampler(a.ntime, B * world, 45, seed=7)

    def next_slice():
        f, c, q = sampler.next_batch()
        s = slice(rank * B, (rank + 1) * B)
        return f[s], c[s], q[s]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)          # NVML polling across warm-up + timed region
    # warm-up (captures the CUDA graphs) ---------------------------------------------------
    for _ in range(W):
        trainer.train_step(*next_slice())
    trainer.flush_losses()

    # ---- value: indices pre-staged on the device ----------------------------------------
    staged = torch.zeros(K, 4 * B, dtype=torch.int32)
    for k in range(K):
        f, c, q = next_slice()
        staged[k, :B] = torch.from_numpy(f)
        staged[k, B:2 * B] = torch.from_numpy(c)
        staged[k, 2 * B:].view(torch.int64)[:] = torch.from_numpy(q)
    staged = staged.cuda()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(K):
        trainer.idx_dev.copy_(staged[k])
        trainer.run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    losses = trainer.flush_losses()
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms / 1e3)

    # ---- e2e, host-driven trainer: cube upload from pinned host memory + K steps ------------
    barrier()
    t0 = time.perf_counter()
    cube.upload([data], [missing], fetch_mean=False)
    for k in range(K):
        trainer.train_step(*next_slice())
    e2e_losses = trainer.flush_losses()
    barrier()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    cube_bytes = data.size * 5                       # f32 data + u8 mask
    e2e_trainer = {"value": world * B * K / t_e2e, "unit": UNIT,
                   "h2d_bytes_per_step": int(16 * B + cube_bytes / K), "d2h_bytes_per_step": 32,
                   "note": "upload of the %d-frame cube from pinned host memory (%.0f MB, amortised over "
                           "the %d steps) + per-step index H2D (16 B/frame) + per-step loss-sum D2H, "
                           "through Trainer.train_step" % (a.ntime, cube_bytes / 1e6, K)}

    # ---- reconstruction sweep sharded by time range over all ranks, one output file -----------
    recon_shard = None
    if world > 1 and not a.no_recon_shard:
        import tempfile
        from dincae_b200 import api
        rec_s = Reconstructor(plan, cube, B, net=trainer.net, use_graph=not a.no_graph,
                              records=(cube.meandata[0], -9999.), n_out_slots=api._SAVE_SLOTS)
        box = [tempfile.mkdtemp() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        for name in ("warm.nc", "data.nc"):                 # the first pass captures the graphs
            barrier()
            t0 = time.perf_counter()
            api._save_records(rec_s, os.path.join(box[0], name), lon, lat, cube.meandata[0], a.ntime, B,
                              rank, world, dist.barrier)
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            recon_shard = a.ntime / float(t.item())
        del rec_s

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clk, "e2e": e2e_trainer,
            "e2e_trainer": e2e_trainer,
            "final_loss": losses[-1][0] if losses else None,
            "loss_first_last": [losses[0][0], e2e_losses[-1][0]] if losses and e2e_losses else None}
    if recon_shard is not None:
        line["recon_to_file_fps_all_gpus"] = recon_shard

    if rank == 0:
        # ---- launches per step + per-call timing (eager pass, CUDA events) ------------------
        lib.dincae_launch_count(1)
        trainer._enqueue_fwd_bwd()
        trainer._enqueue_update()
        torch.cuda.synchronize()
        per_step = int(lib.dincae_launch_count(1))
        line["gpu_launches"] = per_step * K
        line["gpu_launches_per_step"] = per_step

        agg = traced_step(trainer, B, [0.05], 5)
        total_ms = sum(v[0] for v in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][0])
        peak, peak_src = peaks()
        ach = top[1][1] / (top[1][0] * 1e-3) / 1e9
        # DRAM bytes of that kernel from the committed `ncu --set full` capture (profiles/),
        # per launch; null when the capture holds another kernel
        traffic = None
        for tp in ("r2_traffic.json", "r1_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if traffic is None and os.path.isfile(tp):
                t = json.load(open(tp)).get(top[0])
                if t:
                    traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": traffic, "kernel": top[0],
                            "kernel_ms": top[1][0], "kernel_share_of_step": top[1][0] / total_ms,
                            "algorithmic_bytes": top[1][1], "algorithmic_flops": top[1][2],
                            "peak_source": peak_src,
                            "note": "per-call times from an eager pass with the side stream off; the "
                                    "captured step overlaps weight gradients with the data-gradient chain",
                            "step_algorithmic_bytes": sum(v[1] for v in agg.values()),
                            "step_hbm_frac": sum(v[1] for v in agg.values()) / (ms / K * 1e-3) / 1e9 / peak}
        line["kernel_breakdown_ms"] = {k: round(v[0], 4) for k, v in
                                       sorted(agg.items(), key=lambda kv: -kv[1][0])}

        # ---- reconstruction sweep (forward only, results copied to pinned host memory) ------
        rec = Reconstructor(plan, cube, B, net=trainer.net, use_graph=not a.no_graph)
        nrec_frames = min(a.ntime, 40 * B)
        batches = list(sequential_batches(nrec_frames, B))
        for fr in batches[:3]:
            rec.run_batch(fr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, fr in enumerate(batches):
            rec.run_batch(fr, slot=i % 2)
        torch.cuda.synchronize()
        line["recon_fps"] = nrec_frames / (time.perf_counter() - t0)
        # forward pass of one frame: input 4*nvar + xrec/records 24 B per pixel + every activation once
        fwd_bytes = sum(v[1] for k, v in agg.items() if k.startswith("conv3x3_fwd") or k.startswith("dense_fwd"))
        fwd_bytes += agg["build_batch"][1] + 24 * B * a.size * a.size
        line["recon_hbm_frac"] = fwd_bytes / B * line["recon_fps"] / 1e9 / peak

        # ---- the same sweep through to the output file (a12: forward + savesample arithmetic on
        #      the device, records written by threads; NetCDF-3 unless netCDF4 is installed) -----
        import tempfile
        from dincae_b200 import api
        rec2 = Reconstructor(plan, cube, B, net=trainer.net, use_graph=not a.no_graph,
                             records=(cube.meandata[0], -9999.), n_out_slots=api._SAVE_SLOTS)
        with tempfile.TemporaryDirectory() as d:
            api._save_records(rec2, os.path.join(d, "warm.nc"), lon, lat, cube.meandata[0], 2 * B + (a.ntime % B), B)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            api._save_records(rec2, os.path.join(d, "data.nc"), lon, lat, cube.meandata[0], a.ntime, B)
            line["recon_to_file_fps"] = a.ntime / (time.perf_counter() - t0)
            line["recon_file_bytes"] = os.path.getsize(os.path.join(d, "data.nc"))
        del rec, rec2
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        elif world > 1:
            line["cpu_baseline"] = None       # measured at N = 1 only (see the N = 1 line / --impl reference)

    # the config-A trainer is done: free it before the other legs
    del trainer, cube
    torch.cuda.empty_cache()

    # ---- e2e through the reference-facing call (file in, file out) ----------------------------
    if not a.no_e2e_api:
        try:
            e2e_api = e2e_api_leg(a, world, rank, dist, barrier, K)
            line["e2e"] = e2e_api
        except Exception as e:            # noqa: BLE001 -- keep the line; say what failed
            line["e2e_api_error"] = repr(e)[:300]
        torch.cuda.empty_cache()

    # ---- BASELINE configs[3]: 512x512, bf16 ------------------------------------------------------
    if not a.no_config_d:
        try:
            line["config_d"] = config_d_leg(a, world, rank, dist, barrier)
        except Exception as e:            # noqa: BLE001
            line["config_d"] = {"error": repr(e)[:300]}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version
    # banner, prints of the reference-facing API) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (dincae_b200 has no CPU fallback); "
                             "use --impl reference for the CPU arm")
        run_ours(a)


if __name__ == "__main__":
    main()
