"""BASELINE.json configs[4]: inference-only reconstruction sweep over the batch size, plus the
ensemble averaging of K saved reconstructions (device kernels only, data resident)."""
import sys, time, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
from dincae_b200 import synthetic, ensemble
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Plan
from dincae_b200.sampler import sequential_batches
from dincae_b200.trainer import Reconstructor

T, H, W = 2048, 112, 112
lon, lat, time_, data, mask = synthetic.sst_like_cube(T=T, H=H, W=W, seed=12345)
missing = np.ma.getmaskarray(data)
cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
plan = Plan(H, W, cube.nvar)
out = {"recon_fps": {}, "config": "112x112, default architecture, T=%d, results copied to pinned host memory" % T}
for B in (1, 8, 32, 50, 128, 512):
    rec = Reconstructor(plan, cube, B, use_graph=True)
    rec.net.set_params_tf(plan.glorot_init(0))
    batches = list(sequential_batches(min(T, max(40 * B, 256)), B))
    for fr in batches[:3]:
        rec.run_batch(fr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for i, fr in enumerate(batches):
        rec.run_batch(fr, slot=i % 2)
        n += len(fr)
    torch.cuda.synchronize()
    out["recon_fps"][B] = round(n / (time.perf_counter() - t0), 1)
    del rec
    torch.cuda.empty_cache()
# ensemble averaging of K members of a full reconstruction (T x H x W), device-resident
K, nel = 10, T * H * W
acc = ensemble.EnsembleAccumulator(nel)
mem = [(torch.randn(nel, device="cuda"), torch.rand(nel, device="cuda") + 0.1) for _ in range(2)]
for m, s in mem:
    acc.add(m, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    acc.add(*mem[k % 2])
mean, sig = acc.result("mixture")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
bytes_ = K * nel * (8 + 3 * 16 + 8) + nel * (3 * 8 + 4 + 8)
out["ensemble"] = {"members": K, "elements": nel, "ms": round(ms, 3), "GB_per_s": round(bytes_ / ms / 1e6, 1)}
print(json.dumps(out))
