"""Stand-alone timing of the dense-bottleneck kernels at BASELINE configs[3] sizes (K = 43008,
N = 8296): python tools/dense_bench.py [B] [world]"""
import os
import sys

import torch

sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib  # noqa: E402
from dincae_b200._lib import check, ptr  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K, N = 43008, 8296
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for (k, n) in ((K, N), (N, K)):
    w = torch.randn(k, n, device="cuda") * 0.01
    m = torch.zeros_like(w)
    v = torch.zeros_like(w)
    bias = torch.zeros(n, device="cuda")
    x = torch.randn(B, k, device="cuda")
    y = torch.zeros(B, n, device="cuda")
    dz = torch.randn(B, n, device="cuda")
    dx = torch.zeros(B, k, device="cuda")
    ws_b = max(lib.dincae_dense_workspace(B, k, n), lib.dincae_dense_factored_workspace(world * B, k, n), 1 << 20)
    ws = torch.zeros(ws_b, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: check(lib.dincae_dense_fwd(ptr(x), ptr(w), ptr(bias), B, k, n, 1, ptr(y), ptr(ws), ws_b, st)))
    print("dense_fwd      B=%d %dx%d  %7.3f ms  %6.0f GB/s" % (B, k, n, ms, 4.0 * k * n / ms / 1e6), flush=True)
    ms = timeit(lambda: check(lib.dincae_dense_bwd_data(ptr(dz), ptr(w), None, B, k, n, ptr(dx), ptr(ws), ws_b, st)))
    print("dense_bwd_data B=%d %dx%d  %7.3f ms  %6.0f GB/s" % (B, k, n, ms, 4.0 * k * n / ms / 1e6), flush=True)
    # factors of `world` ranks, each B rows
    xf = torch.randn(world, B * k, device="cuda")
    zf = torch.randn(world, B * n, device="cuda")
    w8 = torch.zeros(k * n, dtype=torch.bfloat16, device="cuda")
    state = torch.tensor([0.9, 0.999, 0.0, 1e-3, 1.0, 0, 0, 0], dtype=torch.float32, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    for tma in ("0", "1"):
        os.environ["DINCAE_ADAM_TMA"] = tma                 # register-staged kernel / TMA pipeline (Bt <= 16)
        ms = timeit(lambda: check(lib.dincae_adam_update_factored(
            ptr(w), ptr(m), ptr(v), n, ptr(xf), B * k, k, ptr(zf), B * n, n, world, B, k, n, 1.0, None, 0.9, 0.999,
            1e-8, ptr(state), ptr(w8), st)))
        print("adam_factored  Bt=%d %dx%d  DINCAE_ADAM_TMA=%s %7.3f ms  %6.0f GB/s (26 B/param incl. bf16 copy)" %
              (world * B, k, n, tma, ms, 26.0 * k * n / ms / 1e6), flush=True)
    del os.environ["DINCAE_ADAM_TMA"]
    ms = timeit(lambda: check(lib.dincae_dense_factored_sumsq(
        ptr(xf), B * k, k, ptr(zf), B * n, n, world, B, k, n, ptr(ss), ptr(ws), ws_b, st)))
    print("gram sumsq     Bt=%d %dx%d  %7.3f ms" % (world * B, k, n, ms), flush=True)
    del w, m, v
