"""numpy restatement of the reference input pipeline (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows DINCAE/__init__.py:120-230 (``data_generator``, ``data_generator_list`` and the
closure ``datagen``), but with every random draw made explicit so that the CUDA batch
builder can be fed the *same* draws:

* ``StaticFeatures``        -- the one-off pre-pass, DINCAE/__init__.py:155-193
* ``stack_frame``           -- time-window stacking of one frame, :198-209
* ``perturb_frame``         -- added clouds + jitter of one frame, :212-225
* ``frame_stream``          -- the generator protocol itself, :196-227

PINNED: ``tests/test_oracle_datagen.py`` checks this file bit-for-bit against arrays
produced by the unmodified reference generator (``tests/golden/datagen_*.npz`` written by
``oracle/make_golden.py``), and, where ``/root/reference`` is present, against the live
reference.
"""
import random as _pyrandom

import numpy as np


def day_of_year(time):
    """tm_yday of each entry (DINCAE/__init__.py:160)."""
    return np.array([t.timetuple().tm_yday for t in time])


def temporal_mean(field):
    """Mean over time of a (masked) [T,H,W] field, keepdims (DINCAE/__init__.py:168).

    numpy's MaskedArray.mean = (float32 sum of the 0-filled data, accumulated slice by
    slice along axis 0) * 1. / (integer count of valid samples) -> float64, masked where
    the count is zero.  An array without a mask takes ndarray.mean (float32 result).
    """
    field = np.ma.asarray(field)
    if field.mask is np.ma.nomask:
        return np.ma.asarray(np.asarray(field).mean(axis=0, keepdims=True))
    valid = ~np.ma.getmaskarray(field)
    total = np.add.reduce(np.where(valid, field.data, field.dtype.type(0)), axis=0, keepdims=True)
    count = valid.sum(axis=0, keepdims=True)
    never = count == 0
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = total * 1. / count
    return np.ma.masked_array(np.where(never, 0., mean), never)


class StaticFeatures:
    """Per-time-step feature cube ``x[T,H,W,2*ndata+4]`` (DINCAE/__init__.py:155-193).

    channel 2i   : anomaly_i / obs_err_std_i**2 (0 where missing)
    channel 2i+1 : (1 - missing_i) / obs_err_std_i**2
    then lon and lat scaled to [-1, 1], cos and sin of 2*pi*dayofyear/365.25.
    """

    def __init__(self, lon, lat, time, data_full, obs_err_std):
        self.ndata = len(data_full)
        self.shape = data_full[0].shape
        T, H, W = self.shape
        self.ntime = T
        nstatic = 2 * self.ndata + 4
        self.meandata = []
        x = np.zeros((T, H, W, nstatic), dtype="float32")
        for i, field in enumerate(data_full):
            if field.shape != self.shape:
                raise ValueError("shape are not coherent")
            mean = temporal_mean(field)
            anomaly = field - mean
            var = obs_err_std[i] ** 2
            x[..., 2 * i] = anomaly.filled(0) / var
            x[..., 2 * i + 1] = (1 - np.ma.getmask(anomaly)) / var
            self.meandata.append(mean)
        doy = day_of_year(time)
        lon = np.asarray(lon)
        lat = np.asarray(lat)
        k = 2 * self.ndata
        x[..., k] = (2 * (lon - lon.min()) / (lon.max() - lon.min()) - 1)[None, None, :]
        x[..., k + 1] = (2 * (lat - lat.min()) / (lat.max() - lat.min()) - 1)[None, :, None]
        x[..., k + 2] = np.cos(2 * np.pi * doy / 365.25)[:, None, None]
        x[..., k + 3] = np.sin(2 * np.pi * doy / 365.25)[:, None, None]
        self.x = x

    def nvar(self, ntime_win):
        return 2 * ntime_win * self.ndata + 4


def window_offsets(ntime_win):
    """Neighbour offsets in channel order: past ... future, centre skipped (:202-206)."""
    half = ntime_win // 2
    return [k - half for k in range(ntime_win) if k != half]


def stack_frame(feat, i, ntime_win):
    """Un-perturbed network input for time index ``i`` (DINCAE/__init__.py:198-209)."""
    T, H, W = feat.shape
    nd2 = 2 * feat.ndata
    xin = np.zeros((H, W, feat.nvar(ntime_win)), dtype="float32")
    xin[..., : nd2 + 4] = feat.x[i]
    at = nd2 + 4
    for nn in window_offsets(ntime_win):
        j = min(T - 1, max(0, i + nn))
        xin[..., at: at + nd2] = feat.x[j, :, :, :nd2]
        at += nd2
    return xin


def perturb_frame(xin, missing, imask, noise, jitter_std):
    """Training-mode perturbation, in place (DINCAE/__init__.py:212-225).

    ``missing[j][imask]`` is laid over the CURRENT-time channels of every variable j, then
    jitter is added to the three data channels the reference hard-wires (current, first
    and second neighbour slot -- this assumes ntime_win == 3, reference quirk Q4).
    ``noise`` is a list of 3*ndata float64 [H,W] standard-normal draws in the reference's
    call order (j-major; current, +2nd+4, +4nd+4).  numpy evaluates ``f32 += f64`` in
    float64 and rounds once.
    """
    ndata = len(missing)
    for j in range(ndata):
        sel = missing[j][imask]
        xin[..., 2 * j][sel] = 0
        xin[..., 2 * j + 1][sel] = 0
    for j in range(ndata):
        xin[..., 2 * j] += jitter_std[j] * noise[3 * j]
        xin[..., 2 * j + 2 * ndata + 4] += jitter_std[j] * noise[3 * j + 1]
        xin[..., 2 * j + 4 * ndata + 4] += jitter_std[j] * noise[3 * j + 2]
    return xin


def frame_stream(feat, missing, train, ntime_win, jitter_std,
                 draw_imask=None, draw_noise=None):
    """Generator function with the reference's protocol (DINCAE/__init__.py:196-227).

    Yields ``(xin[H,W,nvar] f32, xtrue[H,W,2] f32)``.  By default the draws come from
    Python's ``random`` and ``np.random`` exactly as in the reference, so seeding both
    reproduces its stream.
    """
    draw_imask = draw_imask or (lambda n: _pyrandom.randrange(0, n))
    draw_noise = draw_noise or (lambda h, w: np.random.randn(h, w))
    T, H, W = feat.shape

    def datagen():
        for i in range(T):
            xin = stack_frame(feat, i, ntime_win)
            if train:
                imask = draw_imask(T)
                noise = [draw_noise(H, W) for _ in range(3 * feat.ndata)]
                perturb_frame(xin, missing, imask, noise, jitter_std)
            yield xin, feat.x[i, :, :, 0:2]

    return datagen


def data_generator_list(lon, lat, time, data_full, missing, train=True, ntime_win=3,
                        obs_err_std=(1.,), jitter_std=(0.05,)):
    """Same return contract as the reference: (datagen, nvar, ntime, meandata[0])."""
    feat = StaticFeatures(lon, lat, time, data_full, obs_err_std)
    gen = frame_stream(feat, missing, train, ntime_win, list(jitter_std))
    return gen, feat.nvar(ntime_win), feat.ntime, feat.meandata[0]


def data_generator(lon, lat, time, data_full, missing, train=True, ntime_win=3,
                   obs_err_std=1., jitter_std=0.05):
    """Single-variable front end; like the reference it IGNORES ``train`` (quirk Q1,
    DINCAE/__init__.py:126-127)."""
    return data_generator_list(lon, lat, time, [data_full], [missing], train=True,
                               ntime_win=ntime_win, obs_err_std=[obs_err_std],
                               jitter_std=[jitter_std])
