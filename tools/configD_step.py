"""Functional check of BASELINE.json configs[3] shapes: 512x512, 4 variables (nvar 28), encoder
[32,48,72,108,162]: a few training steps + one reconstruction batch; prints losses and time."""
import sys, time, numpy as np, torch, random
sys.path.insert(0, "/root/repo")
from dincae_b200 import synthetic
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Plan
from dincae_b200.sampler import TrainSampler
from dincae_b200.trainer import Trainer, Reconstructor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T, H, W = max(12, B + 4), 512, 512
LANES = 8 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else 4
fields, missing = [], []
for k in range(4):
    lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=W, seed=100 + k)
    fields.append(data); missing.append(np.ma.getmaskarray(data))
cube = DeviceCube(lon, lat, time_, fields, missing, [1.0] * 4, 3)
print("nvar", cube.nvar)
plan = Plan(H, W, cube.nvar, (32, 48, 72, 108, 162), (0.2,), (1, 2, 3, 4, 5), lanes=LANES)
tr = Trainer(plan, cube, B, noise_seed=1, jitter_std=[0.05] * 4, use_graph=False)
tr.set_lr(1e-3)
random.seed(3)
sm = TrainSampler(T, B, 8, seed=7)
for k in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.train_step(*sm.next_batch())
    out = tr.flush_losses()
    torch.cuda.synchronize()
    print("step %d loss %.5f rms %.5f  %.1f ms" % (k, out[0][0], out[0][1], 1e3 * (time.perf_counter() - t0)))
    assert np.isfinite(out[0][0])
rec = Reconstructor(plan, cube, B, net=tr.net, use_graph=False)
o = rec.run_batch(np.arange(B))
torch.cuda.synchronize()
print("recon finite:", bool(torch.isfinite(o).all()), "params", plan.n_params_tf,
      "GB allocated %.1f" % (torch.cuda.max_memory_allocated() / 1e9))

# ---- per-call breakdown of one training step (eager, CUDA events; side stream serialised) ------
from dincae_b200 import _lib
records = []


def tracer(label, fn, args, nbytes, nflops):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args), label)
    e1.record()
    records.append((label, e0, e1, nbytes, nflops))


tr.net.tracer = tracer
reps = 3
for _ in range(reps):
    tr.net.n_noncloud.zero_()
    torch.cuda._sleep(20_000_000)
    tr.net.forward(B)
    tr.net.loss_backward(B, False, normalise=False)
    tr.net.adam_step(denom=tr.net.loss_sums[3:4])
    torch.cuda.synchronize()
tr.net.tracer = None
agg = {}
for label, e0, e1, nb, nf in records:
    d = agg.setdefault(label, [0.0, nb, nf])
    d[0] += e0.elapsed_time(e1) / reps
tot = sum(v[0] for v in agg.values())
print("B=%d  sum of calls %.2f ms" % (B, tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%8.3f ms  %5.1f%%  %7.1f GB/s alg  %7.2f TFLOP/s alg  %s" % (
        v[0], 100 * v[0] / tot, v[1] / v[0] / 1e6, v[2] / v[0] / 1e9, k))
