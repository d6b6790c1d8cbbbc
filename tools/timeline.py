"""Role timeline of CTA 0 of one conv3x3_tc launch (library built with -DDINCAE_TC_TIMELINE)."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib, layout
from dincae_b200._lib import ptr, check
lib = _lib.load()
lib.dincae_debug_timeline.restype = ctypes.c_int
lib.dincae_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
def st(): return torch.cuda.current_stream().cuda_stream
B,H,W,Cin,Cout = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (50,112,112,10,16)
g = torch.Generator().manual_seed(0)
x = torch.randn(B,H,W,Cin, generator=g)
k = torch.randn(3,3,Cin,Cout, generator=g)*0.2
cinp, coutp = (Cin+3)//4, (Cout+3)//4
cpad = layout.round_up(Cout,16)
wp = torch.from_numpy(layout.pack_conv(k.numpy(), [Cin], cpad)).cuda()
bias = torch.zeros(cpad, device="cuda")
xd = torch.from_numpy(layout.nhwc_to_p4(x.numpy())).cuda()
lib.dincae_set_conv_backend(1)
op = torch.zeros(B,coutp,(H+1)//2,(W+1)//2,4, device="cuda")
sg = torch.zeros(B,coutp,H,(W+3)//4, dtype=torch.int16, device="cuda")
def run():
    check(lib.dincae_conv3x3_fwd(ptr(xd), cinp, 0, None, 0, ptr(wp), ptr(bias), cpad, B,H,W,coutp, 0.2, None, ptr(op), ptr(sg), st()))
for _ in range(3): run()
torch.cuda.synchronize()
lib.dincae_debug_timeline(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
print("kernel us", e0.elapsed_time(e1)*1e3)
buf = np.zeros((3*1024,2), dtype=np.int64)
lib.dincae_debug_timeline(buf.ctypes.data, 3*1024, 1)
ev = buf[buf[:,1] != 0]
ev = ev[np.argsort(ev[:,0], kind="stable")]
t0 = ev[0,0]
names = {103:"L A-loads issued", 104:"L W stored", 105:"L A stored", 100:"L loads issued", 101:"L slot free", 102:"L stage full", 200:"M wait full", 201:"M got full", 202:"M issued", 300:"E wait acc", 301:"E got acc", 302:"E done"}
lo, hi = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (0, 400)
for t, c in ev[lo:hi]:
    print("%8d  %s%s" % (t - t0, " " * (20 * (c // 100 - 1)), names.get(int(c), str(c))))
print("... total events", len(ev), "span cycles", ev[-1,0]-t0)
