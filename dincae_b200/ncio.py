"""netCDF input/output for the DINCAE file layouts.

The reference reads its input with netCDF4-python (DINCAE/__init__.py:84-92) and writes
``meandata``, ``mean_rec`` and ``sigma_rec`` with it (:251-294).  When ``netCDF4`` is
importable it is used the same way (NETCDF4 files).  This image has neither netCDF4 nor
HDF5, so the fallback backend is ``scipy.io.netcdf_file`` (NetCDF-3, 64-bit offset) with
identical dimensions, variable names, dtypes and ``_FillValue`` attributes; readers such as
xarray / ncdump open both.  ``num2date`` is restated for the CF "<unit> since <date>"
strings the reference relies on (:87).
"""
import datetime
import re

import numpy as np

try:                                    # pragma: no cover - not available in this image
    import netCDF4 as _nc4
except Exception:                       # noqa: BLE001
    _nc4 = None

from scipy.io import netcdf_file as _nc3

_UNIT_SECONDS = {"day": 86400., "days": 86400., "d": 86400.,
                 "hour": 3600., "hours": 3600., "hr": 3600., "h": 3600.,
                 "minute": 60., "minutes": 60., "min": 60.,
                 "second": 1., "seconds": 1., "sec": 1., "s": 1.}


def have_netcdf4():
    return _nc4 is not None


def num2date(values, units):
    """CF time decoding for the (proleptic) Gregorian calendar -> list of datetime."""
    if _nc4 is not None:                # pragma: no cover
        return _nc4.num2date(values, units)
    m = re.match(r"\s*(\w+)\s+since\s+(.+?)\s*$", units)
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError("cannot parse time units %r" % units)
    scale = _UNIT_SECONDS[m.group(1).lower()]
    ref = m.group(2).replace("T", " ").replace("Z", "").strip()
    ref = re.sub(r"\s*(UTC|utc)$", "", ref)
    fmt_try = ["%Y-%m-%d %H:%M:%S.%f", "%Y-%m-%d %H:%M:%S", "%Y-%m-%d %H:%M", "%Y-%m-%d"]
    t0 = None
    for fmt in fmt_try:
        try:
            t0 = datetime.datetime.strptime(ref, fmt)
            break
        except ValueError:
            continue
    if t0 is None:
        parts = [int(p) for p in re.split(r"[-: ]+", ref) if p][:6]
        parts += [1] * (3 - len(parts)) if len(parts) < 3 else []
        t0 = datetime.datetime(*parts)
    vals = np.ma.filled(np.ma.asarray(values, dtype="float64"), np.nan).reshape(-1)
    return [t0 + datetime.timedelta(seconds=float(v) * scale) for v in vals]


def _decode_attr(a):
    return a.decode() if isinstance(a, bytes) else a


def read_gridded(fname, varname):
    """-> lon, lat, time (datetimes), data (masked f32 [T,lat,lon]), mask2d or None.

    Mirrors the reads of DINCAE/__init__.py:84-92: ``_FillValue`` / ``missing_value``
    entries (and NaNs) become masked.
    """
    if _nc4 is not None:                # pragma: no cover
        ds = _nc4.Dataset(fname)
        lon = np.ma.getdata(ds.variables["lon"][:])
        lat = np.ma.getdata(ds.variables["lat"][:])
        time = num2date(ds.variables["time"][:], ds.variables["time"].units)
        data = ds.variables[varname][:, :, :]
        mask = np.ma.getdata(ds.variables["mask"][:, :]) if "mask" in ds.variables else None
        ds.close()
        return lon, lat, time, np.ma.asarray(data), mask
    with _nc3(fname, "r", mmap=False) as ds:
        lon = np.array(ds.variables["lon"][:])
        lat = np.array(ds.variables["lat"][:])
        tv = ds.variables["time"]
        time = num2date(np.array(tv[:]), _decode_attr(tv.units))
        v = ds.variables[varname]
        raw = np.array(v[:, :, :])
        miss = np.zeros(raw.shape, dtype=bool)
        for att in ("_FillValue", "missing_value"):
            if hasattr(v, att):
                miss |= raw == np.asarray(getattr(v, att)).reshape(-1)[0]
        if np.issubdtype(raw.dtype, np.floating):
            miss |= np.isnan(raw)
        scale = getattr(v, "scale_factor", None)
        offset = getattr(v, "add_offset", None)
        if scale is not None or offset is not None:
            raw = raw * (1.0 if scale is None else scale) + (0.0 if offset is None else offset)
        data = np.ma.masked_array(raw, miss if miss.any() else np.ma.nomask)
        mask = np.array(ds.variables["mask"][:, :]) if "mask" in ds.variables else None
    return lon, lat, time, data, mask


def write_gridded(fname, varname, lon, lat, time_values, time_units, data, mask,
                  fill_value=-9999.):
    """Input-file writer with the schema of examples/create_input_file.py:28-52."""
    filled = np.ma.filled(np.ma.asarray(data), fill_value).astype("f4")
    if _nc4 is not None:                # pragma: no cover
        ds = _nc4.Dataset(fname, "w", format="NETCDF4")
        ds.createDimension("lon", len(lon))
        ds.createDimension("lat", len(lat))
        ds.createDimension("time", None)
        ds.createVariable("lon", "f4", ("lon",))[:] = lon
        ds.createVariable("lat", "f4", ("lat",))[:] = lat
        tv = ds.createVariable("time", "f8", ("time",))
        tv.units = time_units
        tv[:] = time_values
        dv = ds.createVariable(varname, "f4", ("time", "lat", "lon"), fill_value=fill_value)
        dv[:, :, :] = np.ma.asarray(data)
        if mask is not None:
            ds.createVariable("mask", "i4", ("lat", "lon"))[:, :] = np.asarray(mask, dtype="i4")
        ds.close()
        return
    with _nc3(fname, "w", version=2) as ds:
        ds.createDimension("time", None)          # NetCDF-3: the record dimension comes first
        ds.createDimension("lon", len(lon))
        ds.createDimension("lat", len(lat))
        ds.createVariable("lon", "f4", ("lon",))[:] = np.asarray(lon, dtype="f4")
        ds.createVariable("lat", "f4", ("lat",))[:] = np.asarray(lat, dtype="f4")
        if mask is not None:
            mv = ds.createVariable("mask", "i4", ("lat", "lon"))
            mv[:, :] = np.asarray(mask, dtype="i4")
        tv = ds.createVariable("time", "f8", ("time",))
        tv.units = time_units
        dv = ds.createVariable(varname, "f4", ("time", "lat", "lon"))
        dv._FillValue = np.float32(fill_value)
        tv[:len(time_values)] = np.asarray(time_values, dtype="f8")
        dv[:len(time_values)] = filled


class ReconWriter:
    """Output file of the reconstruction sweep: same layout as ``savesample``
    (DINCAE/__init__.py:251-278): dims time (unlimited), lon, lat; f4 variables
    lon(lon), lat(lat), meandata(lat,lon), mean_rec(time,lat,lon), sigma_rec(time,lat,lon),
    the last three with _FillValue -9999.  Unlike the reference the file stays open between
    batches (append = slice assignment on the record dimension)."""

    def __init__(self, fname, lon, lat, meandata, fill_value=-9999., append=False):
        self.fname = fname
        self.fill = np.float32(fill_value)
        self.mask = np.ma.getmaskarray(np.ma.asarray(meandata)).reshape(len(lat), len(lon))
        if append:                      # reference: Dataset(fname, 'a') for ii > 0 (:281-283)
            self.ds = _nc4.Dataset(fname, "a") if _nc4 is not None else _nc3(fname, "a")
            self.v_mean = self.ds.variables["mean_rec"]
            self.v_sigma = self.ds.variables["sigma_rec"]
            return
        md = np.ma.filled(np.ma.asarray(meandata), fill_value).reshape(len(lat), len(lon))
        if _nc4 is not None:            # pragma: no cover
            ds = _nc4.Dataset(fname, "w", format="NETCDF4")
            ds.createDimension("time", None)
            ds.createDimension("lon", len(lon))
            ds.createDimension("lat", len(lat))
            ds.createVariable("lon", "f4", ("lon",))[:] = lon
            ds.createVariable("lat", "f4", ("lat",))[:] = lat
            v = ds.createVariable("meandata", "f4", ("lat", "lon"), fill_value=fill_value)
            v[:, :] = np.ma.masked_array(md, self.mask)
            self.v_mean = ds.createVariable("mean_rec", "f4", ("time", "lat", "lon"),
                                            fill_value=fill_value)
            self.v_sigma = ds.createVariable("sigma_rec", "f4", ("time", "lat", "lon"),
                                             fill_value=fill_value)
            self.ds = ds
            return
        ds = _nc3(fname, "w", version=2)
        ds.createDimension("time", None)
        ds.createDimension("lon", len(lon))
        ds.createDimension("lat", len(lat))
        ds.createVariable("lon", "f4", ("lon",))[:] = np.asarray(lon, dtype="f4")
        ds.createVariable("lat", "f4", ("lat",))[:] = np.asarray(lat, dtype="f4")
        v = ds.createVariable("meandata", "f4", ("lat", "lon"))
        v._FillValue = self.fill
        v[:, :] = md.astype("f4")
        self.v_mean = ds.createVariable("mean_rec", "f4", ("time", "lat", "lon"))
        self.v_mean._FillValue = self.fill
        self.v_sigma = ds.createVariable("sigma_rec", "f4", ("time", "lat", "lon"))
        self.v_sigma._FillValue = self.fill
        self.ds = ds

    def write(self, offset, mean_rec, sigma_rec):
        """mean_rec / sigma_rec [n,lat,lon]; entries under meandata.mask get the fill value
        (:288-291)."""
        n = mean_rec.shape[0]
        m = np.broadcast_to(self.mask, mean_rec.shape)
        a = np.where(m, self.fill, np.asarray(mean_rec, dtype="f4"))
        b = np.where(m, self.fill, np.asarray(sigma_rec, dtype="f4"))
        self.v_mean[offset:offset + n] = a
        self.v_sigma[offset:offset + n] = b

    def close(self):
        if self.ds is not None:
            self.ds.close()
            self.ds = None


def read_recon(fname):
    """-> dict(lon, lat, meandata, mean_rec, sigma_rec) as masked arrays (for tests/tools)."""
    out = {}
    if _nc4 is not None:                # pragma: no cover
        ds = _nc4.Dataset(fname)
        for k in ("lon", "lat", "meandata", "mean_rec", "sigma_rec"):
            out[k] = np.ma.asarray(ds.variables[k][:])
        ds.close()
        return out
    with _nc3(fname, "r", mmap=False) as ds:
        out["dims"] = {k: v for k, v in ds.dimensions.items()}
        for k in ("lon", "lat", "meandata", "mean_rec", "sigma_rec"):
            v = ds.variables[k]
            raw = np.array(v[:])
            raw = raw.astype(raw.dtype.newbyteorder("="))
            if hasattr(v, "_FillValue"):
                raw = np.ma.masked_array(raw, raw == np.asarray(v._FillValue).reshape(-1)[0])
            out[k] = np.ma.asarray(raw)
            out[k + "_dtype"] = str(raw.dtype.newbyteorder("="))
            out[k + "_dims"] = v.dimensions
    return out
