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

METRIC = "train frames/sec at 112x112 (recon frames/sec alongside)"
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
                    help="N > 1: also time the reconstruction sweep sharded over all ranks into one file")
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
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:            # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:            # noqa: BLE001
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    print(json.dumps(line), flush=True)


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
    clocks = ClockSampler(local)
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

    # ---- e2e: cube upload from pinned host memory + K host-driven steps --------------------
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
    e2e = {"value": world * B * K / t_e2e, "unit": UNIT,
           "h2d_bytes_per_step": int(16 * B + cube_bytes / K),
           "d2h_bytes_per_step": 32,
           "note": "timed region = upload of the %d-frame cube (%.0f MB, amortised over the %d "
                   "steps) + per-step index H2D (16 B/frame) + per-step loss-sum D2H, through "
                   "Trainer.train_step" % (a.ntime, cube_bytes / 1e6, K)}

    # ---- opt-in: reconstruction sweep sharded by time range over all ranks, one output file ---
    recon_shard = None
    if a.recon_shard and world > 1:
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

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": workload_config(a, world), "clocks": clk, "e2e": e2e,
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

        records = []

        def tracer(label, fn, args, nbytes, nflops):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(*args), label)
            e1.record()
            records.append((label, e0, e1, nbytes, nflops))

        reps = 5
        trainer.net.tracer = tracer
        ebb0, ebb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bb_ms = 0.0
        for _ in range(reps):
            trainer.net.n_noncloud.zero_()
            # keep the GPU busy while the host enqueues the step, so that the kernels run back
            # to back and the events measure execution, not the python launch gaps
            torch.cuda._sleep(8_000_000)
            ebb0.record()
            trainer.cube.build_batch(trainer.frame_idx, trainer.cloud_idx, B, trainer.net.X[0],
                                     trainer.net.xtrue, trainer.net.n_noncloud, jitter_std=[0.05],
                                     seed=1, noise_seq=trainer.noise_seq)
            ebb1.record()
            trainer.net.forward(B)
            trainer.net.loss_backward(B, False, normalise=False)
            trainer.net.adam_step(denom=trainer.net.loss_sums[3:4])
            torch.cuda.synchronize()
            bb_ms += ebb0.elapsed_time(ebb1)
        trainer.net.tracer = None
        agg = {}
        for label, e0, e1, nb, nf in records:
            d = agg.setdefault(label, [0.0, nb, nf])
            d[0] += e0.elapsed_time(e1) / reps
        nvar = cube.nvar
        agg["build_batch"] = [bb_ms / reps, B * a.size * a.size * (3 * 5 + 1 + 4 * nvar + 8), 0]
        total_ms = sum(v[0] for v in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][0])
        peak, peak_src = peaks()
        ach = top[1][1] / (top[1][0] * 1e-3) / 1e9
        # DRAM bytes of that kernel from the committed `ncu --set full` capture (profiles/),
        # per launch; null when the capture holds another kernel
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.isfile(tp):
            t = json.load(open(tp)).get(top[0])
            if t:
                traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
        line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": traffic, "kernel": top[0],
                            "kernel_ms": top[1][0], "kernel_share_of_step": top[1][0] / total_ms,
                            "algorithmic_bytes": top[1][1], "algorithmic_flops": top[1][2],
                            "peak_source": peak_src,
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

        if not a.no_cpu_baseline:
            cb = cpu_arm(a, a.cpu_steps, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["recon_fps"] = cb["recon_fps"]
            line["cpu_baseline"]["generator_only_fps"] = cb["generator_only_fps"]
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
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
