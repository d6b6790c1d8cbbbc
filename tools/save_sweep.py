"""Reconstruction sweep to file (api._save_records) vs number of writer threads / slots."""
import os, sys, tempfile, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from dincae_b200 import api
from dincae_b200.data import DeviceCube
from dincae_b200.engine import Plan
from dincae_b200.trainer import Reconstructor
a = type("A", (), {"ntime": 5266, "size": 112})()
B = 50
lon, lat, time_, data, mask = bench.make_cube(a, pinned=True)
cube = DeviceCube(lon, lat, time_, [data], [np.ma.getmaskarray(data)], [1.0], 3)
plan = Plan(112, 112, cube.nvar)
base = Reconstructor(plan, cube, B)
for threads, slots in [(1, 4), (3, 6), (6, 10), (10, 16)]:
    api._SAVE_THREADS = threads
    rec = Reconstructor(plan, cube, B, net=base.net, records=(cube.meandata[0], -9999.), n_out_slots=slots)
    with tempfile.TemporaryDirectory() as d:
        api._save_records(rec, os.path.join(d, "warm.nc"), lon, lat, cube.meandata[0], 2 * B + a.ntime % B, B)
        best = 1e9
        for k in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            api._save_records(rec, os.path.join(d, "data%d.nc" % k), lon, lat, cube.meandata[0], a.ntime, B)
            best = min(best, time.perf_counter() - t0)
        print("threads %2d slots %2d: %.1f ms, %.0f frames/s" % (threads, slots, 1e3 * best, a.ntime / best))
