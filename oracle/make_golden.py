"""Write tests/golden/datagen_*.npz from the UNMODIFIED reference generator.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For each case the reference's ``data_generator`` / ``data_generator_list``
(DINCAE/__init__.py:120-230) is run with Python ``random`` and ``np.random`` seeded, and
every frame it yields is stored next to the inputs and seeds.  tests/test_oracle_datagen.py
replays the same seeds through ``oracle/datagen_np.py`` and demands bit equality; the GPU
tests then check the CUDA batch builder against that (now pinned) restatement.
"""
import datetime
import os
import random

import numpy as np

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

T0 = datetime.datetime(1900, 1, 1)


def make_field(rs, T, H, W, frac_missing, land=True):
    d = (15 + 3 * rs.standard_normal((T, H, W))).astype("float32")
    m = rs.uniform(size=(T, H, W)) < frac_missing
    if land:
        m[:, 0, :2] = True          # never observed -> masked meandata
        m[:, H - 1, W - 1] = True
    return d, m


CASES = {
    # name: (T, H, W, nfields, train, ntime_win, obs_err_std, jitter_std, day0, seeds, api)
    "single_train": (8, 10, 7, 1, True, 3, [1.0], [0.05], 360, (3, 4), "data_generator"),
    "single_testflag": (6, 5, 9, 1, False, 3, [1.0], [0.05], 10, (5, 6), "data_generator"),
    "multi_train": (7, 6, 8, 3, True, 3, [1.0, 2.0, 0.5], [0.05, 0.1, 0.0], 59, (7, 8), "list"),
    "multi_test": (7, 6, 8, 2, False, 3, [1.0, 1.0], [0.0, 0.0], 200, (9, 10), "list"),
    "win5_test": (9, 4, 5, 1, False, 5, [1.0], [0.0], 0, (11, 12), "list"),
}


def run_case(ref, name):
    T, H, W, nf, train, win, oes, jit, day0, seeds, api = CASES[name]
    rs = np.random.RandomState(sum(map(ord, name)))
    lon = np.linspace(-7, -0.05, W) + 0.01 * rs.standard_normal(W).cumsum()
    lat = np.linspace(33, 39.95, H)
    days = day0 + np.arange(T) * 37            # crosses year boundaries, incl. a leap day
    time = [T0 + datetime.timedelta(days=int(d)) for d in days]
    fields, missings = [], []
    for k in range(nf):
        d, m = make_field(rs, T, H, W, 0.3 + 0.1 * k)
        fields.append(d)
        missings.append(m)
    data_full = [np.ma.masked_array(d, m) for d, m in zip(fields, missings)]

    random.seed(seeds[0])
    np.random.seed(seeds[1])
    if api == "data_generator":
        gen, nvar, ntime, meandata = ref.data_generator(
            lon, lat, time, data_full[0], missings[0], train=train, ntime_win=win,
            obs_err_std=oes[0], jitter_std=jit[0])
    else:
        gen, nvar, ntime, meandata = ref.data_generator_list(
            lon, lat, time, data_full, missings, train=train, ntime_win=win,
            obs_err_std=oes, jitter_std=jit)
    frames = list(gen())
    xin = np.stack([f[0] for f in frames])
    xtrue = np.stack([f[1] for f in frames])
    # a second pass continues the same RNG streams (the train pipeline repeats the generator)
    frames2 = list(gen())
    xin2 = np.stack([f[0] for f in frames2])
    np.savez_compressed(
        os.path.join(OUT, "datagen_%s.npz" % name),
        lon=lon, lat=lat, days=days, data=np.stack(fields), missing=np.stack(missings),
        train=train, ntime_win=win, obs_err_std=np.array(oes), jitter_std=np.array(jit),
        seeds=np.array(seeds), api=api, nvar=nvar, ntime=ntime,
        meandata=np.ma.getdata(meandata), meandata_mask=np.ma.getmaskarray(meandata),
        meandata_dtype=str(meandata.dtype),
        xin=xin, xtrue=xtrue, xin_pass2=xin2)
    print(name, "nvar", nvar, "xin", xin.shape, xin.dtype, "meandata", meandata.dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load()
    for name in CASES:
        run_case(ref, name)


if __name__ == "__main__":
    main()
