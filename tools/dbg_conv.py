import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from dincae_b200 import _lib, layout
from dincae_b200._lib import ptr, check
lib = _lib.load()
def st(): return torch.cuda.current_stream().cuda_stream
B,H,W,Cin,Cout = [int(a) for a in sys.argv[1:6]] if len(sys.argv) > 5 else (2,12,20,10,16)
g = torch.Generator().manual_seed(0)
x = torch.randn(B,H,W,Cin, generator=g)
k = torch.randn(3,3,Cin,Cout, generator=g)*0.2
cinp, coutp = (Cin+3)//4, (Cout+3)//4
cpad = layout.round_up(Cout,16)
wp = torch.from_numpy(layout.pack_conv(k.numpy(), [Cin], cpad)).cuda()
bias = torch.zeros(cpad, device="cuda")
xd = torch.from_numpy(layout.nhwc_to_p4(x.numpy())).cuda()
outs = []
for be in (0,1):
    lib.dincae_set_conv_backend(be)
    of = torch.zeros(B,coutp,H,W,4, device="cuda")
    op = torch.zeros(B,coutp,(H+1)//2,(W+1)//2,4, device="cuda")
    check(lib.dincae_conv3x3_fwd(ptr(xd), cinp, 0, None, 0, ptr(wp), ptr(bias), cpad, B,H,W,coutp, 0.2, ptr(of), ptr(op), None, st()))
    torch.cuda.synchronize()
    outs.append((of.cpu().numpy(), op.cpu().numpy()))
d = np.abs(outs[0][0]-outs[1][0])   # [B,P,H,W,4]
print("max err full", d.max(), "ref max", np.abs(outs[0][0]).max())
bad = d.max(axis=(1,4)) > 1e-2   # [B,H,W]
for b in range(B):
    print("image", b)
    for y in range(H):
        print("".join("X" if bad[b,y,xx] else "." for xx in range(W)))
dp = np.abs(outs[0][1]-outs[1][1])
print("max err pool", dp.max())
