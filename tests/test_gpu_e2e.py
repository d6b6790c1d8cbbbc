"""End-to-end through the reference-facing API on a synthetic SST-like file: same calls as
the reference's own tests (test/test_DINCAE_SST.py:83-131), plus checks of the output file
layout (DINCAE/__init__.py:251-294) that the reference never makes."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_example(tmp_path_factory):
    from dincae_b200 import synthetic
    d = tmp_path_factory.mktemp("data")
    fname = str(d / "sst_small.nc")
    lon, lat, time, data, mask = synthetic.sst_like_cube(T=60, H=40, W=48, seed=1)
    synthetic.write_input_file(fname, "SST", lon, lat, time, data, mask)
    return fname, "SST", (lon, lat, time, data, mask)


def test_load(small_example):
    import DINCAE
    filename, varname, (lon, lat, time, data, mask) = small_example
    lon2, lat2, time2, data2, missing, mask2 = DINCAE.load_gridded_nc(filename, varname)
    assert np.allclose(lon2, lon) and np.allclose(lat2, lat)
    assert time2[0] == time[0] and time2[-1] == time[-1]
    assert np.array_equal(mask2, mask)
    assert np.array_equal(missing, np.ma.getmaskarray(data))
    train_datagen, nvar, train_len, meandata = DINCAE.data_generator(lon2, lat2, time2, data2, missing)
    xin, xtrue = next(train_datagen())
    assert nvar == 10 and train_len == 60
    assert xin.shape == (40, 48, 10) and xtrue.shape == (40, 48, 2) and xin.dtype == np.float32


def test_reconstruct_gridded_nc_all_options(small_example, tmp_path):
    import DINCAE
    filename, varname, (lon, lat, time, data, mask) = small_example
    outdir = str(tmp_path / "temp-result")
    loss = []
    ret = DINCAE.reconstruct_gridded_nc(
        filename, varname, outdir, iseed=12345, epochs=3, save_each=1, tensorboard=True,
        nprefetch=1, nepoch_keep_missing=10, learning_rate_decay_epoch=100,
        truth_uncertain=True, regularization_L2_beta=0.001, loss=loss, batch_size=16)
    assert ret is None                                  # reference quirk Q7
    assert len(loss) == 3 * 4 and all(np.isfinite(loss))
    assert loss[-1] < 48                                # bound of test_DINCAE_SST.py:107
    files = sorted(glob.glob(os.path.join(outdir, "data-*.nc")))
    assert len(files) == 3                              # save_each=1: one per epoch, no overwrite
    from dincae_b200 import ncio
    out = ncio.read_recon(files[-1])
    assert out["mean_rec"].shape == (60, 40, 48) and out["sigma_rec"].shape == (60, 40, 48)
    assert out["mean_rec_dtype"] == "float32"
    never = np.ma.getmaskarray(data).all(axis=0)
    assert np.array_equal(np.ma.getmaskarray(out["mean_rec"])[0], never)
    assert np.array_equal(np.ma.getmaskarray(out["meandata"]), never)
    assert np.all(out["sigma_rec"].compressed() > 0)
    assert os.path.isfile(os.path.join(outdir, "model-001.ckpt.pt"))
    # TensorBoard event file: the reference's scalars and images (DINCAE/__init__.py:575-584), one
    # entry per optimisation step; 3 images per tag, flipped and masked, at the grid's size
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(os.path.join(outdir, "train"), size_guidance={"images": 0, "scalars": 0})
    acc.Reload()
    tags = acc.Tags()
    assert set(tags["scalars"]) >= {"Validation/cost", "Validation/RMS"}
    assert [ev.step for ev in acc.Scalars("Validation/cost")] == list(range(12))
    assert abs(acc.Scalars("Validation/cost")[-1].value - loss[-1]) < 1e-6 * max(1.0, abs(loss[-1]))
    for name in ("m_rec", "m_true", "sigma2_rec"):
        for j in range(3):
            evs = acc.Images("Validation/%s/image/%d" % (name, j))
            assert len(evs) == 12 and evs[0].width == 48 and evs[0].height == 40


def test_reconstruct_gridded_files_learns(small_example, tmp_path):
    import DINCAE
    filename, varname, (lon, lat, time, data, mask) = small_example
    outdir = str(tmp_path / "temp-result2")
    loss = []
    fname = DINCAE.reconstruct_gridded_files(
        [{"filename": filename, "varname": varname}], outdir, iseed=12345, epochs=40,
        save_each=0, loss=loss, batch_size=20)
    assert fname is not None and os.path.isfile(fname)
    assert loss[-1] < 8                                 # bound of test_DINCAE_SST.py:131
    assert np.mean(loss[-3:]) < np.mean(loss[:3])
    from dincae_b200 import ncio
    out = ncio.read_recon(fname)
    truth = np.ma.getdata(data)
    valid = ~np.ma.getmaskarray(data)
    err = (np.ma.getdata(out["mean_rec"]) - truth)[valid]
    assert np.sqrt(np.mean(err ** 2)) < 2.5             # reconstructs observed pixels (sigma of field ~4)


def test_device_records_path_writes_the_same_file_as_the_savesample_hook(small_example, tmp_path):
    """Stock savesample + identity transform = records emitted by the device and written by
    threads; any other hook gets (m_rec, sigma2_rec) on the host like in the reference.  Same
    seed, same (bit-deterministic) training: the two files must hold identical values."""
    import DINCAE
    from dincae_b200 import ncio
    filename, varname, _ = small_example
    calls = []

    def hook(fname, m_rec, s2_rec, *args, **kw):
        calls.append((m_rec.shape, args[4], args[5]))            # e, ii
        return DINCAE.savesample(fname, m_rec, s2_rec, *args, **kw)

    out = []
    for k, kw in enumerate([{}, {"savesample": hook}]):
        fname = DINCAE.reconstruct_gridded_files(
            [{"filename": filename, "varname": varname}], str(tmp_path / ("r%d" % k)), iseed=4321,
            epochs=2, save_each=1, batch_size=16, **kw)
        out.append(ncio.read_recon(fname))
    assert [c[0][0] for c in calls] == [16, 16, 16, 12] * 2      # ragged last batch, 2 sweeps
    a, b = out
    assert a["dims"] == b["dims"] and a["mean_rec"].shape == (60, 40, 48)
    for k in ("lon", "lat", "meandata", "mean_rec", "sigma_rec"):
        assert a[k + "_dims"] == b[k + "_dims"] and a[k + "_dtype"] == b[k + "_dtype"]
        assert np.array_equal(np.ma.getmaskarray(a[k]), np.ma.getmaskarray(b[k])), k
        assert np.array_equal(np.ma.getdata(a[k]).view(np.uint32), np.ma.getdata(b[k]).view(np.uint32)), k


def test_user_generators_take_the_host_feed_path(tmp_path):
    """Arbitrary generator functions (not made by this package) still train (:399-410)."""
    import DINCAE
    rs = np.random.RandomState(0)
    H, W, T = 16, 20, 12
    frames = []
    for i in range(T):
        x = rs.standard_normal((H, W, 10)).astype("f4")
        t = np.zeros((H, W, 2), "f4")
        t[..., 1] = rs.uniform(size=(H, W)) > 0.5
        t[..., 0] = x[..., 0] * t[..., 1]
        frames.append((x, t))

    def gen():
        for f in frames:
            yield f

    loss = []
    meandata = np.ma.masked_array(np.zeros((1, H, W)), np.zeros((1, H, W), bool))
    fname = DINCAE.reconstruct(np.arange(W), np.arange(H), np.ones((H, W), bool), meandata,
                               gen, T, gen, T, str(tmp_path / "o"), epochs=2, batch_size=5,
                               save_each=0, loss=loss, iseed=1)
    assert len(loss) == 2 * 3 and np.isfinite(loss).all()
    from dincae_b200 import ncio
    assert ncio.read_recon(fname)["mean_rec"].shape == (T, H, W)


def test_config_d_shapes_train_and_reconstruct():
    """BASELINE.json configs[3] shapes (512x512, 4 variables -> nvar 28, encoder
    [32,48,72,108,162], 689 M parameters): the layers the tensor-core kernels do not cover
    (image wider than 128 for the weight gradient, more than 96 output channels) take the
    FFMA kernels behind the same entry points; a few steps must run, stay finite and learn."""
    import random
    import torch
    from dincae_b200.data import DeviceCube
    from dincae_b200.engine import Plan
    from dincae_b200.sampler import TrainSampler
    from dincae_b200.trainer import Reconstructor, Trainer
    from tests.util import random_cube
    rs = np.random.RandomState(5)
    T, H, W, B = 6, 512, 512, 2
    lon, lat, time_, fields = random_cube(rs, T, H, W, 4)
    missing = [np.ma.getmaskarray(f) for f in fields]
    cube = DeviceCube(lon, lat, time_, fields, missing, [1.0] * 4, 3)
    assert cube.nvar == 28
    plan = Plan(H, W, cube.nvar, (32, 48, 72, 108, 162), (0.2,), (1, 2, 3, 4, 5))
    assert plan.n_params_tf == 688723466
    tr = Trainer(plan, cube, B, noise_seed=1, jitter_std=[0.05] * 4, use_graph=False)
    tr.set_lr(1e-3)
    random.seed(3)
    sm = TrainSampler(T, B, 4, seed=7)
    losses = []
    for _ in range(3):
        tr.train_step(*sm.next_batch())
        losses += tr.flush_losses()
    assert all(np.isfinite(c) and np.isfinite(r) for c, r in losses)
    rec = Reconstructor(plan, cube, B, net=tr.net, use_graph=False)
    out = rec.run_batch(np.arange(B))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    assert float(out[1].min()) > 0.0                # sigma^2 > 0
    del tr, rec, cube
    torch.cuda.empty_cache()


def test_checkpoint_round_trip_resumes_bit_identically(tmp_path):
    """save_checkpoint / load_checkpoint (the resume path the reference lacks): a trainer
    restored from a checkpoint continues with exactly the same losses as the original."""
    import random
    import torch
    import dincae_b200 as D
    from dincae_b200.data import DeviceCube
    from dincae_b200.engine import Plan
    from dincae_b200.sampler import TrainSampler
    from dincae_b200.trainer import Trainer
    from tests.util import random_cube
    rs = np.random.RandomState(8)
    T, H, W, B = 10, 24, 20, 4
    lon, lat, time_, fields = random_cube(rs, T, H, W, 1)
    missing = [np.ma.getmaskarray(f) for f in fields]
    cube = DeviceCube(lon, lat, time_, fields, missing, [1.0], 3)
    plan = Plan(H, W, cube.nvar)

    def run(tr, batches):
        out = []
        for fb in batches:
            tr.train_step(*fb)
            out += tr.flush_losses()
        return out

    random.seed(1)
    sm = TrainSampler(T, B, 6, seed=2)
    batches = [sm.next_batch() for _ in range(6)]
    a = Trainer(plan, cube, B, noise_seed=3, jitter_std=[0.05], use_graph=False)
    a.set_lr(1e-3)
    run(a, batches[:3])
    path = D.save_checkpoint(str(tmp_path / "model-003.ckpt"), a)
    want = run(a, batches[3:])
    b = Trainer(plan, cube, B, noise_seed=3, jitter_std=[0.05], use_graph=False, init_seed=99)
    b.set_lr(1e-3)
    D.load_checkpoint(path, b)
    assert b.step_count == 3
    got = run(b, batches[3:])
    assert got == want
    assert torch.equal(a.net.params, b.net.params)
