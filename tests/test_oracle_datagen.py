"""The numpy restatement of the input pipeline must equal the reference generator's own
output bit for bit (golden vectors written by oracle/make_golden.py from the unmodified
reference, DINCAE/__init__.py:120-230)."""
import datetime
import glob
import os
import random

import numpy as np
import pytest

from oracle import datagen_np, ref_loader

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "datagen_*.npz")))
T0 = datetime.datetime(1900, 1, 1)


def load_case(path):
    z = np.load(path, allow_pickle=False)
    time = [T0 + datetime.timedelta(days=int(d)) for d in z["days"]]
    data_full = [np.ma.masked_array(d, m) for d, m in zip(z["data"], z["missing"])]
    return z, time, data_full, list(z["missing"])


def run_port(z, time, data_full, missing):
    random.seed(int(z["seeds"][0]))
    np.random.seed(int(z["seeds"][1]))
    if str(z["api"]) == "data_generator":
        return datagen_np.data_generator(
            z["lon"], z["lat"], time, data_full[0], missing[0], train=bool(z["train"]),
            ntime_win=int(z["ntime_win"]), obs_err_std=float(z["obs_err_std"][0]),
            jitter_std=float(z["jitter_std"][0]))
    return datagen_np.data_generator_list(
        z["lon"], z["lat"], time, data_full, missing, train=bool(z["train"]),
        ntime_win=int(z["ntime_win"]), obs_err_std=list(z["obs_err_std"]),
        jitter_std=list(z["jitter_std"]))


def test_golden_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_port_equals_reference_golden(path):
    z, time, data_full, missing = load_case(path)
    gen, nvar, ntime, meandata = run_port(z, time, data_full, missing)
    assert nvar == int(z["nvar"]) and ntime == int(z["ntime"])
    assert str(meandata.dtype) == str(z["meandata_dtype"])
    assert np.array_equal(np.ma.getmaskarray(meandata), z["meandata_mask"])
    keep = ~z["meandata_mask"]
    assert np.array_equal(np.ma.getdata(meandata)[keep], z["meandata"][keep])
    frames = list(gen())
    xin = np.stack([f[0] for f in frames])
    xtrue = np.stack([f[1] for f in frames])
    assert xin.dtype == np.float32 and xtrue.dtype == np.float32
    assert np.array_equal(xin, z["xin"])
    assert np.array_equal(xtrue, z["xtrue"])
    xin2 = np.stack([f[0] for f in gen()])
    assert np.array_equal(xin2, z["xin_pass2"])


def test_single_variable_front_end_ignores_train_flag():
    """Quirk Q1 (DINCAE/__init__.py:126-127): data_generator(train=False) still perturbs."""
    path = [p for p in GOLDEN if "single_testflag" in p][0]
    z, time, data_full, missing = load_case(path)
    assert not bool(z["train"])
    clean = datagen_np.stack_frame(
        datagen_np.StaticFeatures(z["lon"], z["lat"], time, data_full, [1.0]), 0, 3)
    assert not np.array_equal(clean, z["xin"][0])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not on this machine")
def test_port_equals_live_reference_random_shapes():
    ref = ref_loader.load()
    rs = np.random.RandomState(99)
    for trial in range(4):
        T, H, W = rs.randint(3, 9), rs.randint(3, 12), rs.randint(3, 12)
        nf = rs.randint(1, 4)
        time = [T0 + datetime.timedelta(days=int(300 + 20 * i)) for i in range(T)]
        lon, lat = np.sort(rs.uniform(-10, 10, W)), np.sort(rs.uniform(30, 40, H))
        fields = [np.ma.masked_array((rs.standard_normal((T, H, W)) * 4).astype("f4"),
                                     rs.uniform(size=(T, H, W)) < 0.5) for _ in range(nf)]
        missing = [np.ma.getmaskarray(f) for f in fields]
        kw = dict(train=bool(trial % 2 == 0), ntime_win=3,
                  obs_err_std=list(rs.uniform(0.5, 2, nf)), jitter_std=list(rs.uniform(0, 0.2, nf)))
        out = []
        for impl in (ref.data_generator_list, datagen_np.data_generator_list):
            random.seed(trial)
            np.random.seed(trial + 100)
            gen, nvar, ntime, mean = impl(lon, lat, time, fields, missing, **kw)
            fr = list(gen())
            out.append((nvar, ntime, np.ma.filled(mean, -1.), np.stack([f[0] for f in fr]),
                        np.stack([f[1] for f in fr])))
        for a, b in zip(*out):
            assert np.array_equal(a, b)
