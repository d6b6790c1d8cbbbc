"""CUDA input pipeline vs the pinned numpy restatement: BIT-EXACT (integer indexing, masks,
float32 results of float64 arithmetic).  Reference: DINCAE/__init__.py:155-227."""
import datetime
import glob
import os

import numpy as np
import pytest
import torch

from oracle import datagen_np
from tests.util import T0, dev, from_p4, random_cube, tf32_round

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def exact_inputs():
    """Bit-exact parity is checked with the fp32 conv back end selected; with the tensor-core
    back end the builder rounds xin to TF32 on store (test_tensor_core_backend_rounds_inputs)."""
    from dincae_b200 import _lib
    lib = _lib.load()
    lib.dincae_set_conv_backend(0)
    yield
    lib.dincae_set_conv_backend(1)

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "datagen_*.npz")))


def build(cube, frames, clouds, noise, jitter, B):
    xin = torch.zeros(B, cube.nvar_planes, cube.H, cube.W, 4, device="cuda")
    xtrue = torch.zeros(B, cube.H, cube.W, 2, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    cube.build_batch(dev(np.asarray(frames, dtype=np.int32)),
                     None if clouds is None else dev(np.asarray(clouds, dtype=np.int32)),
                     B, xin, xtrue, cnt, jitter_std=jitter,
                     noise=None if noise is None else dev(noise))
    torch.cuda.synchronize()
    return from_p4(xin, cube.nvar), xtrue.cpu().numpy(), int(cnt.item())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_build_batch_equals_reference_golden(path):
    """Golden frames come from the UNMODIFIED reference generator; the draws it consumed are
    recovered by replaying the seeds through the (pinned) numpy port with hooks."""
    import random
    from dincae_b200.data import DeviceCube
    z = np.load(path, allow_pickle=False)
    time = [T0 + datetime.timedelta(days=int(d)) for d in z["days"]]
    fields = [np.ma.masked_array(d, m) for d, m in zip(z["data"], z["missing"])]
    missing = list(z["missing"])
    nd = len(fields)
    T, H, W = fields[0].shape
    train = bool(z["train"]) or str(z["api"]) == "data_generator"     # quirk Q1
    win = int(z["ntime_win"])
    jitter = [float(j) for j in z["jitter_std"]]
    cube = DeviceCube(z["lon"], z["lat"], time, fields, missing,
                      [float(s) for s in z["obs_err_std"]], win)
    # meandata parity (float64, masked where never observed)
    md = cube.meandata[0]
    assert np.array_equal(np.ma.getmaskarray(md), z["meandata_mask"])
    keep = ~z["meandata_mask"]
    assert np.array_equal(np.ma.getdata(md)[keep], z["meandata"][keep])
    # replay the reference's RNG streams to learn the draws of pass 1
    random.seed(int(z["seeds"][0]))
    np.random.seed(int(z["seeds"][1]))
    clouds, noise = [], []
    for i in range(T):
        if train:
            clouds.append(random.randrange(0, T))
            noise.append(np.stack([np.random.randn(H, W) for _ in range(3 * nd)]))
    frames = np.arange(T)
    xin, xtrue, cnt = build(cube, frames, clouds if train else None,
                            np.stack(noise) if train else None, jitter, T)
    assert np.array_equal(xin, z["xin"])
    assert np.array_equal(xtrue, z["xtrue"])
    assert cnt == int((z["xtrue"][..., 1] != 0).sum())


@pytest.mark.parametrize("nf,win,train", [(1, 3, True), (2, 3, True), (4, 3, False), (1, 5, False)])
def test_build_batch_random_order_and_edges(nf, win, train):
    """Shuffled frame order, clamped neighbours at both ends, ragged sizes."""
    from dincae_b200.data import DeviceCube
    rs = np.random.RandomState(nf * 10 + win)
    T, H, W = 9, 13, 11
    lon, lat, time, fields = random_cube(rs, T, H, W, nf)
    missing = [np.ma.getmaskarray(f) for f in fields]
    oes = list(rs.uniform(0.5, 2.0, nf))
    jit = list(rs.uniform(0.0, 0.2, nf))
    feat = datagen_np.StaticFeatures(lon, lat, time, fields, oes)
    cube = DeviceCube(lon, lat, time, fields, missing, oes, win)
    frames = rs.permutation(T)[:7]
    frames[0], frames[1] = 0, T - 1
    clouds = rs.randint(0, T, size=len(frames))
    noise = rs.standard_normal((len(frames), 3 * nf, H, W))
    want_x, want_t = [], []
    for k, i in enumerate(frames):
        xin = datagen_np.stack_frame(feat, int(i), win)
        if train:
            datagen_np.perturb_frame(xin, missing, int(clouds[k]), list(noise[k]), jit)
        want_x.append(xin)
        want_t.append(feat.x[i, :, :, 0:2])
    xin, xtrue, cnt = build(cube, frames, clouds if train else None, noise if train else None,
                            jit, len(frames))
    assert np.array_equal(xin, np.stack(want_x))
    assert np.array_equal(xtrue, np.stack(want_t))
    assert cnt == int((np.stack(want_t)[..., 1] != 0).sum())


def test_unmasked_input_takes_float32_mean_path():
    from dincae_b200.data import DeviceCube
    rs = np.random.RandomState(5)
    T, H, W = 6, 5, 7
    lon, lat, time, _ = random_cube(rs, T, H, W, 1)
    field = np.ma.asarray((rs.standard_normal((T, H, W)) + 20).astype("f4"))
    assert field.mask is np.ma.nomask
    feat = datagen_np.StaticFeatures(lon, lat, time, [field], [1.0])
    cube = DeviceCube(lon, lat, time, [field], [np.zeros((T, H, W), bool)], [1.0], 3)
    xin, xtrue, _ = build(cube, np.arange(T), None, None, [0.0], T)
    want = np.stack([datagen_np.stack_frame(feat, i, 3) for i in range(T)])
    assert np.array_equal(xin, want)


@pytest.mark.parametrize("T,H,W,nf", [(70, 8, 12, 2), (70, 5, 7, 1), (33, 4, 4, 1)])
def test_prepare_cube_long_time_axis(T, H, W, nf):
    """Time axes longer than the mean kernel's register window (ordered float32 sum across
    several windows + a tail) through both anomaly kernels (vectorised: H*W % 4 == 0)."""
    from dincae_b200.data import DeviceCube
    rs = np.random.RandomState(T + H)
    lon, lat, time, fields = random_cube(rs, T, H, W, nf)
    fields[0][:, 1, 1] = np.ma.masked                      # a second never-observed pixel
    missing = [np.ma.getmaskarray(f) for f in fields]
    oes = list(rs.uniform(0.5, 2.0, nf))
    feat = datagen_np.StaticFeatures(lon, lat, time, fields, oes)
    cube = DeviceCube(lon, lat, time, fields, missing, oes, 3)
    for k in range(nf):
        want = fields[k].mean(axis=0, keepdims=True)
        assert np.array_equal(np.ma.getmaskarray(cube.meandata[k]), np.ma.getmaskarray(want))
        keep = ~np.ma.getmaskarray(want)
        assert np.array_equal(np.ma.getdata(cube.meandata[k])[keep], np.ma.getdata(want)[keep])
    xin, xtrue, _ = build(cube, np.arange(T), None, None, [0.0] * nf, T)
    want = np.stack([datagen_np.stack_frame(feat, i, 3) for i in range(T)])
    assert np.array_equal(xin, want)
    assert np.array_equal(xtrue, feat.x[:, :, :, 0:2])


def test_philox_jitter_statistics_and_determinism():
    """Production noise: N(0, jitter^2) on the three data channels, reproducible, keyed by
    the frame's sequence number (not by its slot in the batch)."""
    from dincae_b200.data import DeviceCube
    rs = np.random.RandomState(7)
    T, H, W = 4, 64, 64
    lon, lat, time, fields = random_cube(rs, T, H, W, 1, frac=0.0)
    fields = [np.ma.masked_array(np.ma.getdata(fields[0]), np.zeros((T, H, W), bool))]
    cube = DeviceCube(lon, lat, time, fields, [np.zeros((T, H, W), bool)], [1.0], 3)
    B = 4

    def run(frames, seqs, seed):
        xin = torch.zeros(B, cube.nvar_planes, H, W, 4, device="cuda")
        xtrue = torch.zeros(B, H, W, 2, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        cube.build_batch(dev(np.asarray(frames, np.int32)), dev(np.zeros(B, np.int32)), B, xin,
                         xtrue, cnt, jitter_std=[0.1], seed=seed,
                         noise_seq=dev(np.asarray(seqs, np.int64)))
        return from_p4(xin, cube.nvar), xtrue.cpu().numpy()

    a, t = run([0, 1, 2, 3], [10, 11, 12, 13], 99)
    b, _ = run([3, 2, 1, 0], [13, 12, 11, 10], 99)
    c, _ = run([0, 1, 2, 3], [10, 11, 12, 13], 100)
    assert np.array_equal(a, b[::-1])
    assert not np.array_equal(a, c)
    d = a[..., 0] - t[..., 0]
    assert abs(d.mean()) < 3e-3 and abs(d.std() - 0.1) < 3e-3
    for ch in (6, 8):
        assert a[..., ch].std() > 0


def test_tensor_core_backend_rounds_inputs():
    """Under conv back end 1 xin is the TF32 rounding (nearest, ties away) of the exact
    result; xtrue and the observed-pixel count are untouched."""
    from dincae_b200 import _lib
    from dincae_b200.data import DeviceCube
    rs = np.random.RandomState(11)
    T, H, W = 7, 9, 13
    lon, lat, time, fields = random_cube(rs, T, H, W, 1)
    missing = [np.ma.getmaskarray(f) for f in fields]
    cube = DeviceCube(lon, lat, time, fields, missing, [1.3], 3)
    frames = [0, 3, 6, 2]
    clouds = [1, 5, 0, 4]
    noise = rs.standard_normal((4, 3, H, W))
    exact = build(cube, frames, clouds, noise, [0.05], 4)
    lib = _lib.load()
    lib.dincae_set_conv_backend(1)
    rounded = build(cube, frames, clouds, noise, [0.05], 4)
    assert np.array_equal(rounded[0], tf32_round(exact[0]))
    assert np.array_equal(rounded[1], exact[1])
    assert rounded[2] == exact[2]


@pytest.mark.parametrize("nf,win,backend", [(4, 3, 1), (1, 3, 1), (2, 3, 0), (1, 5, 1)])
def test_bf16_p8_output_equals_fp32_build_plus_conversion(nf, win, backend):
    """dincae_build_batch_p8 (the bf16 network's input written directly, with or without the fp32
    copy) == dincae_build_batch followed by dincae_p4_to_p8, bit for bit -- explicit noise and the
    Philox path, compile-time and run-time channel decode, planes beyond the stacked channels zero."""
    from dincae_b200 import _lib, layout
    from dincae_b200._lib import check, ptr
    from dincae_b200.data import DeviceCube
    lib = _lib.load()
    lib.dincae_set_conv_backend(backend)
    rs = np.random.RandomState(nf + win)
    T, H, W, B = 6, 12, 20, 5
    lon, lat, time, fields = random_cube(rs, T, H, W, nf)
    missing = [np.ma.getmaskarray(f) for f in fields]
    cube = DeviceCube(lon, lat, time, fields, missing, list(rs.uniform(0.5, 2.0, nf)), win)
    frames = dev(rs.randint(0, T, size=B).astype(np.int32))
    clouds = dev(rs.randint(0, T, size=B).astype(np.int32))
    seq = dev(np.arange(B, dtype=np.int64) + 40)
    planes8 = layout.planes8(cube.nvar) + 1                 # one spare plane: must come back zero
    for noise in (dev(rs.standard_normal((B, 3 * nf, H, W))), None):
        kw = dict(jitter_std=[0.1] * nf, noise=noise, seed=5, noise_seq=seq)
        xin = torch.zeros(B, cube.nvar_planes, H, W, 4, device="cuda")
        xtrue = torch.zeros(B, H, W, 2, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        cube.build_batch(frames, clouds, B, xin, xtrue, cnt, **kw)
        want8 = torch.full((B, planes8, H, W, 8), 7.0, dtype=torch.bfloat16, device="cuda")
        check(lib.dincae_p4_to_p8(ptr(xin), cube.nvar_planes, planes8, B, H, W, ptr(want8),
                                  torch.cuda.current_stream().cuda_stream))
        for with_f32 in (True, False):
            xin2 = torch.zeros_like(xin) if with_f32 else None
            got8 = torch.full_like(want8, 3.0)
            xtrue2, cnt2 = torch.zeros_like(xtrue), torch.zeros_like(cnt)
            cube.build_batch(frames, clouds, B, xin2, xtrue2, cnt2, xin8=got8, planes8=planes8, **kw)
            torch.cuda.synchronize()
            assert torch.equal(got8.view(torch.int16), want8.view(torch.int16))
            assert torch.equal(xtrue2, xtrue) and int(cnt2.item()) == int(cnt.item())
            if with_f32:
                assert torch.equal(xin2, xin)
        assert float(want8.float().abs().max()) > 0 and float(want8[:, -1].float().abs().max()) == 0
