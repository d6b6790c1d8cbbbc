"""Synthetic SST-like cubes with land and clouds (bench / test inputs).

There is no network for the AVHRR files the reference's tests download
(test/test_DINCAE_SST.py:19,45), so benchmarks and tests run on seeded synthetic cubes of
the same shape and statistics: a seasonal cycle plus a spatially smooth, AR(1)-in-time
anomaly field, a fixed land blob (~25 % of the pixels, ``mask == 0``) and per-frame cloud
fields leaving 30-80 % of the sea pixels missing.  ``write_input_file`` stores a cube with
the schema of the reference's examples/create_input_file.py:28-52.
"""
import datetime

import numpy as np


def _smooth_field(rs, T, H, W, corr):
    """[T,H,W] unit-variance field, smooth over ~corr pixels (coarse noise, cubic zoom)."""
    from scipy import ndimage
    h = max(2, int(np.ceil(H / corr)) + 3)
    w = max(2, int(np.ceil(W / corr)) + 3)
    coarse = rs.standard_normal((T, h, w)).astype("float32")
    fine = ndimage.zoom(coarse, (1, corr, corr), order=3, mode="nearest", prefilter=False)
    y0 = (fine.shape[1] - H) // 2
    x0 = (fine.shape[2] - W) // 2
    fine = fine[:, y0:y0 + H, x0:x0 + W]
    return fine / max(float(fine.std()), 1e-6)


def sst_like_cube(T=200, H=112, W=112, seed=12345, corr=10, rho=0.9, kind="sst"):
    """Return ``lon, lat, time, data (masked f32 [T,H,W]), mask (bool [H,W])``.

    ``kind``: "sst" (seasonal field with land + clouds), "chl" (log-normal, land + clouds),
    "wind" (complete field, no gaps).
    """
    rs = np.random.RandomState(seed)
    lon = np.linspace(-7, -0.05, W)
    lat = np.linspace(33, 39.95, H)
    t0 = datetime.datetime(1900, 1, 1)
    time = [t0 + datetime.timedelta(days=int(i)) for i in range(T)]
    doy = np.array([t.timetuple().tm_yday for t in time], dtype="float64")

    innov = _smooth_field(rs, T, H, W, corr)
    anom = np.empty_like(innov)
    anom[0] = innov[0]
    s = np.float32(np.sqrt(1 - rho * rho))
    for i in range(1, T):
        anom[i] = np.float32(rho) * anom[i - 1] + s * innov[i]

    phase = (lat - lat.min()) / (lat.max() - lat.min()) * 0.5
    season = 5 * np.sin(2 * np.pi * doy[:, None, None] / 365.25 - phase[None, :, None])
    if kind == "chl":
        field = np.exp(0.5 * anom + 0.2 * season - 1.0)
    elif kind == "wind":
        field = 3 * anom + 0.3 * season
    else:
        field = 18 + season + anom
    field = field.astype("float32")

    yy, xx = np.mgrid[0:H, 0:W]
    blob = ((xx - 0.18 * W) / (0.42 * W)) ** 2 + ((yy - 0.85 * H) / (0.5 * H)) ** 2
    wobble = _smooth_field(rs, 1, H, W, corr)[0]
    land = (blob + 0.15 * wobble) < 0.6
    mask = ~land
    if kind == "wind":
        mask = np.ones((H, W), dtype=bool)
        missing = np.zeros((T, H, W), dtype=bool)
    else:
        cloud = _smooth_field(rs, T, H, W, max(4, corr // 2))
        thresh = rs.uniform(-0.55, 0.85, size=(T, 1, 1)).astype("float32")
        missing = (cloud < thresh) | land[None]
    data = np.ma.masked_array(field, missing)
    return lon, lat, time, data, mask


def write_input_file(filename, varname, lon, lat, time, data, mask, fill_value=-9999.):
    """Write a cube in the layout the reference expects (DINCAE/__init__.py:68-81)."""
    from . import ncio
    t0 = datetime.datetime(1900, 1, 1)
    days = np.array([(t - t0).total_seconds() / 86400. for t in time])
    ncio.write_gridded(filename, varname, lon, lat, days,
                       "days since 1900-01-01 00:00:00", data, mask, fill_value)
