"""Where the time of DeviceCube.upload goes at config A-full (5266 x 112 x 112): the two host ->
device copies (pinned values + pinned mask) and the two pre-pass kernels, timed one by one."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from dincae_b200 import _lib
from dincae_b200._lib import check, ptr, stream_ptr
from dincae_b200.data import DeviceCube
a = type("A", (), {"ntime": 5266, "size": 112})()
lon, lat, time_, data, mask = bench.make_cube(a, pinned=True)
missing = np.ma.getmaskarray(data)
cube = DeviceCube(lon, lat, time_, [data], [missing], [1.0], 3)
torch.cuda.synchronize()
for k in range(3):
    t0 = time.perf_counter()
    cube.upload([data], [missing], fetch_mean=False)
    torch.cuda.synchronize()
    print("upload ms", 1e3 * (time.perf_counter() - t0))
raw = torch.from_numpy(np.ascontiguousarray(np.ma.getdata(data), dtype=np.float32))
mk = torch.from_numpy(np.ascontiguousarray(missing).view(np.uint8))
print("pinned?", raw.is_pinned(), mk.is_pinned())
lib = _lib.load()


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-12s host %.2f ms, total %.2f ms" % (label, 1e3 * (t1 - t0), 1e3 * (t2 - t0)))


for k in range(2):
    timed("mask copy", lambda: cube.missing[0].copy_(mk, non_blocking=True))
    timed("raw copy", lambda: cube._raw_dev.copy_(raw, non_blocking=True))
    timed("prepare", lambda: check(lib.dincae_prepare_cube(
        ptr(cube._raw_dev), ptr(cube.missing[0]), cube.T, cube.H, cube.W, 1.0, 0, ptr(cube.anom[0]),
        ptr(cube._mean_dev), ptr(cube._never_dev), stream_ptr())))
