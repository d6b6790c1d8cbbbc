"""Role timeline of CTA 0 of one conv3x3_wgrad_tc launch (library built with -DDINCAE_TC_TIMELINE)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib, layout
from dincae_b200._lib import ptr, check
lib = _lib.load()
lib.dincae_debug_timeline_wg.restype = ctypes.c_int
lib.dincae_debug_timeline_wg.argtypes = [ctypes.c_void_p, ctypes.c_int]
def st(): return torch.cuda.current_stream().cuda_stream
B,H,W,Cin,Cout = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (50,112,112,10,16)
g = torch.Generator().manual_seed(0)
x = torch.randn(B,H,W,Cin, generator=g)
G = torch.randn(B,H,W,Cout, generator=g)
cinp, coutp = (Cin+3)//4, (Cout+3)//4
cpad = layout.round_up(Cout,16)
xd = torch.from_numpy(layout.nhwc_to_p4(x.numpy())).cuda()
Gd = torch.from_numpy(layout.nhwc_to_p4(G.numpy())).cuda()
lib.dincae_set_conv_backend(1)
wsb = lib.dincae_conv3x3_wgrad_workspace(cinp, 0, cpad, B, H, W)
ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
dw = torch.zeros(9, cinp, cpad, 4, device="cuda"); db = torch.zeros(cpad, device="cuda")
def run():
    check(lib.dincae_conv3x3_wgrad(ptr(xd), cinp, 0, None, 0, ptr(Gd), None, None, 0.2, coutp, cpad, B,H,W, ptr(dw), ptr(db), ptr(ws), wsb, st()))
for _ in range(3): run()
torch.cuda.synchronize()
lib.dincae_debug_timeline_wg(None, 1)
run(); torch.cuda.synchronize()
buf = np.zeros((3*1024,2), dtype=np.int64)
lib.dincae_debug_timeline_wg(buf.ctypes.data, 1)
ev = buf[buf[:,1] != 0]
ev = ev[np.argsort(ev[:,0], kind="stable")]
t0 = ev[0,0]
names = {100:"L issued next", 101:"L slot ok", 102:"L committed", 103:"L stage full", 200:"M wait full", 201:"M got full", 202:"M issued"}
lo, hi = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (0, 200)
for t, c in ev[lo:hi]:
    print("%8d  %s%s" % (t - t0, " " * (22 * (c // 100 - 1)), names.get(int(c), str(c))))
print("... total events", len(ev), "span cycles", ev[-1,0]-t0)
