"""Sharded reconstruction sweep on CPU (gloo, world_size 2): every rank writes the records of
its own contiguous time range into the same NetCDF-3 file with positional writes, rank 0
owns the header; no data is exchanged (reconstruction partitioning).
The device side is replaced by a stand-in that emits recognisable records, so this pins the
host protocol of dincae_b200.api._save_records: batch ranges, offsets, barriers, header."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dincae_b200 import api, ncio

H, W, T, B = 5, 6, 47, 4          # 12 batches, the last one ragged; 6 per rank


class _Done:
    def synchronize(self):
        pass


class FakeRecon:
    """run_batch -> big-endian float32 records [B][2][H][W] with mean_rec = frame number,
    sigma_rec = -frame number (what dincae_head_recon_records would have produced)."""

    def __init__(self, nslots):
        self.out_host = [torch.zeros(B, 2, H, W, dtype=torch.int32) for _ in range(nslots)]
        self.batches = []

    def run_batch(self, frames, slot=0):
        self.batches.append((int(frames[0]), len(frames)))
        rec = np.zeros((B, 2, H, W), dtype=">f4")
        rec[:len(frames), 0] = np.asarray(frames, dtype="f4")[:, None, None]
        rec[:len(frames), 1] = -np.asarray(frames, dtype="f4")[:, None, None]
        self.out_host[slot].numpy().view(np.uint8).reshape(-1)[:] = rec.view(np.uint8).reshape(-1)
        return self.out_host[slot]

    def record_event(self):
        return _Done()


def _worker(rank, world, port, fname, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lon, lat = np.linspace(0, 1, W), np.linspace(2, 3, H)
    md = np.ma.masked_array(np.zeros((1, H, W)), np.zeros((1, H, W), bool))
    recon = FakeRecon(3)
    api._save_records(recon, fname, lon, lat, md, T, B, rank=rank, world=world, barrier=dist.barrier)
    ret[rank] = recon.batches
    dist.destroy_process_group()


def test_shard_batches_partition_the_sweep():
    for test_len, bs, world in [(47, 4, 2), (5266, 50, 8), (3, 50, 4), (100, 10, 3), (1, 1, 1)]:
        nb = -(-test_len // bs)
        got = [api.shard_batches(test_len, bs, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == nb
        assert all(a[1] == b[0] for a, b in zip(got, got[1:])) and all(a <= b for a, b in got)


def test_two_ranks_write_one_file(tmp_path):
    fname = str(tmp_path / "sharded.nc")
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, fname, ret), nprocs=2, join=True)
    assert ret[0] == [(0, 4), (4, 4), (8, 4), (12, 4), (16, 4), (20, 4)]
    assert ret[1] == [(24, 4), (28, 4), (32, 4), (36, 4), (40, 4), (44, 3)]
    out = ncio.read_recon(fname)
    assert out["mean_rec"].shape == (T, H, W) and out["dims"]["time"] is None
    want = np.broadcast_to(np.arange(T, dtype="f4")[:, None, None], (T, H, W))
    assert np.array_equal(np.ma.getdata(out["mean_rec"]), want)
    assert np.array_equal(np.ma.getdata(out["sigma_rec"]), -want)
    # exactly header + fixed variables + T records
    w = ncio.RecordFileWriter(str(tmp_path / "ref.nc"), np.linspace(0, 1, W), np.linspace(2, 3, H),
                              np.ma.masked_array(np.zeros((1, H, W)), np.zeros((1, H, W), bool)))
    assert os.path.getsize(fname) == w.rec_start + T * w.recsize
    w.close()


def test_single_process_sweep_same_protocol(tmp_path):
    """world = 1 (the path `reconstruct` takes on one GPU): same function, no barrier, masked
    pixels of meandata keep whatever the device wrote (it writes the fill value there)."""
    fname = str(tmp_path / "single.nc")
    lon, lat = np.linspace(0, 1, W), np.linspace(2, 3, H)
    md = np.ma.masked_array(np.ones((1, H, W)), np.zeros((1, H, W), bool))
    recon = FakeRecon(2)
    api._save_records(recon, fname, lon, lat, md, T, B)
    assert recon.batches == [(o, min(B, T - o)) for o in range(0, T, B)]
    out = ncio.read_recon(fname)
    want = np.broadcast_to(np.arange(T, dtype="f4")[:, None, None], (T, H, W))
    assert np.array_equal(np.ma.getdata(out["mean_rec"]), want)
    assert np.array_equal(np.ma.getdata(out["meandata"]), np.ones((H, W), "f4"))
