"""Device-resident data cube and the GPU batch builder front end.

Replaces the host feature cube ``x[T,H,W,2*ndata+4]`` of DINCAE/__init__.py:155-193 and the
per-frame closure ``datagen`` (:196-227): only the anomaly field (f32) and the missing
mask (u8) of every variable live in HBM; lon/lat/season channels are tiny vectors and the
stacked network input is produced batch by batch by ``dincae_build_batch``.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def day_of_year(time):
    """tm_yday of every time instance (DINCAE/__init__.py:160)."""
    return np.array([t.timetuple().tm_yday for t in time])


def _same_buffer(a, b):
    """True when two arrays are the same view of the same memory (np.ma hands out a fresh view
    object of a field's mask on every access, so ``is`` does not see it)."""
    if a is b:
        return True
    if not (isinstance(a, np.ndarray) and isinstance(b, np.ndarray)):
        return False
    return (a.dtype == b.dtype and a.shape == b.shape and a.strides == b.strides and
            a.__array_interface__["data"][0] == b.__array_interface__["data"][0])


class DeviceCube:
    def __init__(self, lon, lat, time, data_full, missing, obs_err_std=None, ntime_win=3,
                 device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.DincaeError("dincae_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device(device if device is not None else "cuda")
        self.dev = dev
        nd = len(data_full)
        T, H, W = data_full[0].shape
        self.ndata, self.T, self.H, self.W = nd, T, H, W
        self.ntime_win = int(ntime_win)
        self.nvar = 2 * self.ntime_win * nd + 4
        self.nvar_planes = (self.nvar + 3) // 4
        obs_err_std = [1.0] * nd if obs_err_std is None else list(obs_err_std)
        self.inv_var = [float(np.float32(1 / (float(s) ** 2))) for s in obs_err_std]

        self.anom = torch.empty(nd, T, H, W, dtype=torch.float32, device=dev)
        self.missing = torch.empty(nd, T, H, W, dtype=torch.uint8, device=dev)
        self._raw_dev = torch.empty(T, H, W, dtype=torch.float32, device=dev)
        self._mean_dev = torch.empty(H, W, dtype=torch.float64, device=dev)
        self._never_dev = torch.empty(H, W, dtype=torch.uint8, device=dev)
        self._obs_var = [float(s) ** 2 for s in obs_err_std]
        self.meandata = []
        self.cloud_missing = None
        self.upload(data_full, missing)

        lon = np.asarray(lon, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        doy = day_of_year(time)
        f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dev)
        self.lon_scaled = f32(2 * (lon - lon.min()) / (lon.max() - lon.min()) - 1)
        self.lat_scaled = f32(2 * (lat - lat.min()) / (lat.max() - lat.min()) - 1)
        self.doy_cos = f32(np.cos(2 * np.pi * doy / 365.25))
        self.doy_sin = f32(np.sin(2 * np.pi * doy / 365.25))
        self._inv_var_c = _lib.doubles(self.inv_var)
        self._jitter_cache = {}

    def upload(self, data_full, missing=None, fetch_mean=True):
        """(Re)load the cube from host arrays into the SAME device buffers (pointers captured
        in CUDA graphs stay valid) and run the temporal-mean pre-pass on the device."""
        nd, T, H, W = self.ndata, self.T, self.H, self.W
        if fetch_mean:
            self.meandata = []
        for i in range(nd):
            field = np.ma.asarray(data_full[i])
            if field.shape != (T, H, W):
                raise ValueError("shape are not coherent")     # :171-172
            nomask = field.mask is np.ma.nomask
            miss = np.ma.getmaskarray(field)
            # (same memory first: comparing two T*H*W masks on the host costs tens of milliseconds)
            if missing is not None and not _same_buffer(missing[i], miss) and \
                    not np.array_equal(np.asarray(missing[i], dtype=bool), miss):
                # the reference lays `missing` (not data.mask) over the frame as added clouds
                if self.cloud_missing is None:
                    self.cloud_missing = torch.empty_like(self.missing)
                    for k in range(i):
                        self.cloud_missing[k].copy_(self.missing[k])
                self.cloud_missing[i].copy_(
                    torch.from_numpy(np.ascontiguousarray(missing[i], dtype=bool).view(np.uint8)))
            elif self.cloud_missing is not None:
                self.cloud_missing[i].copy_(torch.from_numpy(np.ascontiguousarray(miss).view(np.uint8)))
            # host -> device: straight from the caller's memory when it is already pinned
            # float32 / bool storage, through one pinned staging copy otherwise
            raw = torch.from_numpy(np.ascontiguousarray(np.ma.getdata(field), dtype=np.float32))
            if not raw.is_pinned():
                raw = raw.pin_memory()
            self._raw_dev.copy_(raw, non_blocking=True)
            mk = torch.from_numpy(np.ascontiguousarray(miss).view(np.uint8))
            self.missing[i].copy_(mk, non_blocking=True)
            check(self.lib.dincae_prepare_cube(
                ptr(self._raw_dev), ptr(self.missing[i]), T, H, W, self._obs_var[i],
                int(nomask), ptr(self.anom[i]), ptr(self._mean_dev), ptr(self._never_dev),
                stream_ptr()), "prepare_cube")
            if fetch_mean:
                mean = self._mean_dev.cpu().numpy().reshape(1, H, W)
                never = self._never_dev.cpu().numpy().astype(bool).reshape(1, H, W)
                if nomask:
                    self.meandata.append(np.ma.asarray(mean.astype(np.float32)))
                else:
                    self.meandata.append(np.ma.masked_array(mean, never))

    def nbytes(self):
        return self.anom.numel() * 4 + self.missing.numel()

    def build_batch(self, frame_idx, cloud_idx, B, xin, xtrue, n_noncloud, jitter_std=None,
                    noise=None, seed=0, noise_seq=None, xin8=None, planes8=0):
        """Fill ``xin`` (P4), ``xtrue`` and add to ``n_noncloud`` for frames
        ``frame_idx[:B]`` (device int32).  ``cloud_idx is None`` = test mode (no added clouds,
        no jitter); otherwise ``jitter_std`` (one value per variable) scales the noise.
        ``xin8`` (bf16 P8, ``planes8`` planes): the bf16 network's input written directly;
        ``xin`` may then be None."""
        key = tuple(float(j) for j in (jitter_std if jitter_std is not None else [0.0] * self.ndata))
        if len(key) != self.ndata:
            raise ValueError("jitter_std needs one entry per variable")
        jit = self._jitter_cache.get(key)
        if jit is None:
            jit = self._jitter_cache[key] = _lib.doubles(key)
        if xin8 is not None:
            check(self.lib.dincae_build_batch_p8(
                ptr(self.anom), ptr(self.missing), ptr(self.cloud_missing), ptr(self.lon_scaled),
                ptr(self.lat_scaled), ptr(self.doy_cos), ptr(self.doy_sin), self._inv_var_c, jit,
                self.ndata, self.T, self.H, self.W, self.ntime_win,
                ptr(frame_idx), ptr(cloud_idx), B, ptr(noise), int(seed) & (2 ** 64 - 1),
                ptr(noise_seq), ptr(xin), self.nvar_planes, ptr(xin8), planes8, ptr(xtrue), ptr(n_noncloud),
                stream_ptr()), "build_batch_p8")
            return
        check(self.lib.dincae_build_batch(
            ptr(self.anom), ptr(self.missing), ptr(self.cloud_missing), ptr(self.lon_scaled), ptr(self.lat_scaled),
            ptr(self.doy_cos), ptr(self.doy_sin), self._inv_var_c, jit,
            self.ndata, self.T, self.H, self.W, self.ntime_win,
            ptr(frame_idx), ptr(cloud_idx), B, ptr(noise), int(seed) & (2 ** 64 - 1),
            ptr(noise_seq), ptr(xin), self.nvar_planes, ptr(xtrue), ptr(n_noncloud),
            stream_ptr()), "build_batch")
