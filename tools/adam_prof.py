"""One launch each of the factored Adam kernels on an 8296 x 8296 matrix (the configs[3] kernels
are 43008 x 8296; a full-set ncu replay of those takes minutes), for ncu:
    ncu --set full -k regex:adam_factored_tma -o rep python tools/adam_prof.py [world]"""
import sys

import torch

sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib  # noqa: E402
from dincae_b200._lib import check, ptr  # noqa: E402

lib = _lib.load()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, k, n = 16, 8296, 8296
st = torch.cuda.current_stream().cuda_stream
w = torch.randn(k, n, device="cuda") * 0.01
m = torch.zeros_like(w)
v = torch.zeros_like(w)
xf = torch.randn(world, B * k, device="cuda").to(torch.bfloat16).float()
zf = torch.randn(world, B * n, device="cuda").to(torch.bfloat16).float()
w8 = torch.zeros(k * n, dtype=torch.bfloat16, device="cuda")
state = torch.tensor([0.9, 0.999, 0.0, 1e-3, 1.0, 0, 0, 0], dtype=torch.float32, device="cuda")
for _ in range(3):
    check(lib.dincae_adam_update_factored(ptr(w), ptr(m), ptr(v), n, ptr(xf), B * k, k, ptr(zf), B * n, n, world, B,
                                          k, n, 1.0, None, 0.9, 0.999, 1e-8, ptr(state), ptr(w8), st))
torch.cuda.synchronize()
print("done", world * B)
