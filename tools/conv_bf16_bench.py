"""Stand-alone timing of the bf16 convolution kernels at BASELINE configs[3] layer shapes
(CUDA events, L2 flushed between repetitions): python tools/conv_bf16_bench.py [B] [layer ...]
Quick to put under ncu: ncu --set full -k regex:conv3x3 -c 6 python tools/conv_bf16_bench.py 8 enc1"""
import ctypes
import sys

import torch

sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib, layout  # noqa: E402
from dincae_b200._lib import check, ptr  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
want = sys.argv[2:] or ["enc1", "enc2", "enc3", "dec3", "dec4", "dec5"]
# name: (H, C0 (half-res source for dec), C1 (skip or 0), Cout, decoder?)
LAYERS = {"enc1": (512, 28, 0, 32, False), "enc2": (256, 32, 0, 48, False), "enc3": (128, 48, 0, 72, False),
          "enc4": (64, 72, 0, 108, False), "enc5": (32, 108, 0, 162, False),
          "dec1": (32, 162, 108, 108, True), "dec2": (64, 108, 72, 72, True), "dec3": (128, 72, 48, 48, True),
          "dec4": (256, 48, 32, 32, True), "dec5": (512, 32, 28, 2, True)}
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
b16 = dict(dtype=torch.bfloat16, device="cuda")


def timeit(fn, reps=5):
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


for name in want:
    H, C0, C1, Cout, dec = LAYERS[name]
    W = H
    Hh, Wh = (H + 1) // 2, (W + 1) // 2
    c0p, c1p, co8 = layout.planes8(C0), layout.planes8(C1), layout.planes8(Cout)
    cpad = layout.round_up(Cout, 16)
    cinp4 = 2 * (c0p + c1p)
    wp = torch.randn(9, cinp4, cpad, 4, device="cuda") * 0.05
    bias = torch.zeros(cpad, device="cuda")
    nd_planes = c0p + (c1p if name != "dec5" else 0)
    nd = layout.round_up(8 * min(nd_planes, 32), 16)
    nde = ctypes.c_size_t(0)
    nfe = lib.dincae_conv_weights_bf16_elems(cinp4, cpad, nd, ctypes.byref(nde))
    wf = torch.zeros(nfe, **b16)
    wd = torch.zeros(nde.value, **b16)
    check(lib.dincae_conv_weights_bf16(ptr(wp), cinp4, cpad, 0, nd, ptr(wf), ptr(wd), st))
    if dec:
        src0 = torch.randn(B, c0p, Hh, Wh, 8, **b16)
        src1 = torch.randn(B, c1p, H, W, 8, **b16)
    else:
        src0 = torch.randn(B, c0p, H, W, 8, **b16)
        src1 = None
    out_full = torch.zeros(B, co8, H, W, 8, **b16)
    out_pool = torch.zeros(B, co8, Hh, Wh, 8, **b16)
    sign = torch.zeros(B, co8, H, 4 * ((W + 3) // 4), dtype=torch.uint8, device="cuda")
    G = torch.randn(B, co8, H, W, 8, **b16)
    gpool = torch.randn(B, co8, Hh, Wh, 8, **b16)
    ws_bytes = lib.dincae_conv3x3_wgrad_bf16_workspace(cinp4, cpad)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    dw = torch.zeros(9, cinp4, cpad, 4, device="cuda")
    db = torch.zeros(cpad, device="cuda")
    g_prev = torch.zeros(B, c0p, Hh, Wh, 8, **b16)
    dx_hi = torch.zeros(B, max(c1p, c0p), H, W, 8, **b16)
    flops = 18.0 * (C0 + C1) * Cout * B * H * W

    if dec:
        fwd = lambda: check(lib.dincae_conv3x3_fwd_bf16(ptr(src0), c0p, 1, ptr(src1), c1p, ptr(wf), ptr(bias), cpad,
                                                        B, H, W, co8, 0.2, ptr(out_full), None, None, None, 0, st))
        c_hi = c1p if name != "dec5" else 0
        if nd_planes * 8 <= 256:
            dgr = lambda: check(lib.dincae_conv3x3_dgrad_bf16(ptr(G), None, None, 0.2, co8, ptr(wd), nd, B, H, W,
                                                              c0p, ptr(src0), 0.2, ptr(g_prev), c_hi,
                                                              ptr(dx_hi) if c_hi else None, None, st))
        else:
            dgr = None
        wgr = lambda: check(lib.dincae_conv3x3_wgrad_bf16(ptr(src0), c0p, 1, ptr(src1), c1p, ptr(G), None, None, 0.2,
                                                          co8, cinp4, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws),
                                                          ws_bytes, st))
    else:
        fwd = lambda: check(lib.dincae_conv3x3_fwd_bf16(ptr(src0), c0p, 0, None, 0, ptr(wf), ptr(bias), cpad,
                                                        B, H, W, co8, 0.2, None, ptr(out_pool), ptr(sign), None, 0, st))
        dgr = lambda: check(lib.dincae_conv3x3_dgrad_bf16(None, ptr(gpool), ptr(sign), 0.2, co8, ptr(wd), nd, B, H, W,
                                                          0, None, 1.0, None, c0p, ptr(dx_hi), None, st))
        wgr = lambda: check(lib.dincae_conv3x3_wgrad_bf16(ptr(src0), c0p, 0, None, 0, None, ptr(gpool), ptr(sign), 0.2,
                                                          co8, cinp4, cpad, B, H, W, ptr(dw), ptr(db), ptr(ws),
                                                          ws_bytes, st))
    import os
    kinds = os.environ.get("CONV_BENCH_KINDS", "fwd,dgrad,wgrad").split(",")
    for label, fn in (("fwd", fwd), ("dgrad", dgr), ("wgrad", wgr)):
        if fn is None or label not in kinds:
            continue
        ms = timeit(fn)
        print("%-5s %-5s B=%d %4dx%-4d %3d+%-3d->%-3d  %8.1f us  %7.1f TFLOP/s" %
              (name, label, B, H, W, C0, C1, Cout, 1e3 * ms, flops / ms / 1e9), flush=True)
