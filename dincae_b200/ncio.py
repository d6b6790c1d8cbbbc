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
import os
import re
import struct
import threading

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


# ---------------------------------------------------------------------------------------
# Streaming NetCDF-3 (64-bit offset) writer for the reconstruction file.
#
# scipy.io.netcdf_file keeps every variable in memory and re-allocates the record arrays on
# each append (quadratic in the number of batches: 2.1 s for the 5266 x 112 x 112 cube,
# 60x the time the GPU needs for the forward sweep).  The classic format is simple enough to
# emit directly: a fixed header, the fixed-size variables, then one record per time step
# holding the record variables back to back -- so a batch of n reconstructed frames is ONE
# contiguous byte range of the file, written with a positional write (no seek state: batches
# can be written from several threads), and the record count is patched into the header on
# close.  Format reference: "NetCDF Classic and 64-bit Offset File Formats" (Unidata).
# ---------------------------------------------------------------------------------------
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE, _NC_FLOAT = 10, 11, 12, 5


def _nc_name(name):
    raw = name.encode()
    return struct.pack(">i", len(raw)) + raw + b"\0" * (-len(raw) % 4)


def _nc_var(name, dimids, fill, vsize, begin):
    out = _nc_name(name) + struct.pack(">i", len(dimids)) + b"".join(struct.pack(">i", d) for d in dimids)
    if fill is None:
        out += struct.pack(">ii", 0, 0)
    else:
        out += struct.pack(">ii", _NC_ATTRIBUTE, 1) + _nc_name("_FillValue") + \
            struct.pack(">iif", _NC_FLOAT, 1, fill)
    return out + struct.pack(">iiq", _NC_FLOAT, vsize, begin)


class RecordFileWriter:
    """``outdir/data-*.nc`` with the layout of ``savesample`` (DINCAE/__init__.py:251-278) as a
    NetCDF-3 64-bit-offset file written in streaming fashion.  Dimensions time (unlimited),
    lon, lat; float32 variables lon(lon), lat(lat), meandata(lat,lon), mean_rec(time,lat,lon),
    sigma_rec(time,lat,lon), the last three with ``_FillValue``."""

    def __init__(self, fname, lon, lat, meandata, fill_value=-9999., append=False, expect_records=0):
        self.fname = fname
        self.nlon, self.nlat = len(lon), len(lat)
        self.fill = np.float32(fill_value)
        self.mask = np.ma.getmaskarray(np.ma.asarray(meandata)).reshape(self.nlat, self.nlon)
        self.vsize = self.nlat * self.nlon * 4
        self.recsize = 2 * self.vsize
        header = self._header(0)
        self.rec_start = len(header) + 4 * self.nlon + 4 * self.nlat + self.vsize
        self.numrecs = 0
        self._lock = threading.Lock()
        if append:
            # reference: Dataset(fname, 'a') for ii > 0 (:281-283); only files this class wrote
            with open(fname, "rb") as f:
                head = f.read(len(header))
            if len(head) != len(header) or head[:4] != header[:4] or head[8:] != header[8:]:
                raise ValueError("%s was not written by RecordFileWriter with this grid" % fname)
            self.numrecs = struct.unpack(">i", head[4:8])[0]
            self.fd = os.open(fname, os.O_RDWR)
            return
        md = np.ma.filled(np.ma.asarray(meandata), fill_value).reshape(self.nlat, self.nlon)
        self.fd = os.open(fname, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        fixed = header + np.asarray(lon, dtype=">f4").tobytes() + np.asarray(lat, dtype=">f4").tobytes() + \
            md.astype(">f4").tobytes()
        assert len(fixed) == self.rec_start
        if expect_records > 0:
            # blocks allocated up front: the buffered-write path is the bound of the sweep
            # (~5 GB/s whatever the thread count) and allocation is ~15 % of it
            try:
                os.posix_fallocate(self.fd, 0, self.rec_start + expect_records * self.recsize)
            except (OSError, AttributeError):
                pass
        os.pwrite(self.fd, fixed, 0)

    def _header(self, numrecs):
        nlon, nlat, vs = self.nlon, self.nlat, self.vsize
        fill = float(self.fill)
        dims = struct.pack(">ii", _NC_DIMENSION, 3) + _nc_name("time") + struct.pack(">i", 0) + \
            _nc_name("lon") + struct.pack(">i", nlon) + _nc_name("lat") + struct.pack(">i", nlat)
        # dimension ids: time 0, lon 1, lat 2.  `begin` offsets depend on the header length,
        # which does not depend on their values: build once with zeros, then for real.
        def variables(b_lon, b_lat, b_md, b_rec):
            return struct.pack(">ii", _NC_VARIABLE, 5) + \
                _nc_var("lon", [1], None, 4 * nlon, b_lon) + _nc_var("lat", [2], None, 4 * nlat, b_lat) + \
                _nc_var("meandata", [2, 1], fill, vs, b_md) + \
                _nc_var("mean_rec", [0, 2, 1], fill, vs, b_rec) + \
                _nc_var("sigma_rec", [0, 2, 1], fill, vs, b_rec + vs)
        pre = b"CDF\x02" + struct.pack(">i", numrecs) + dims + struct.pack(">ii", 0, 0)
        hlen = len(pre) + len(variables(0, 0, 0, 0))
        b_lon = hlen
        b_lat = b_lon + 4 * nlon
        b_md = b_lat + 4 * nlat
        return pre + variables(b_lon, b_lat, b_md, b_md + vs)

    def write_records(self, offset, records, n=None):
        """``records``: the byte image of ``n`` records (big-endian float32
        [n][2][lat][lon], already masked with the fill value), e.g. the pinned host buffer
        filled by ``dincae_head_recon_records``.  Thread-safe (positional write)."""
        buf = memoryview(records).cast("B")
        if n is None:
            n = len(buf) // self.recsize
        buf = buf[:n * self.recsize]
        pos = self.rec_start + offset * self.recsize
        while len(buf):
            k = os.pwrite(self.fd, buf, pos)
            buf = buf[k:]
            pos += k
        with self._lock:
            self.numrecs = max(self.numrecs, offset + n)

    def write(self, offset, mean_rec, sigma_rec):
        """mean_rec / sigma_rec [n,lat,lon] (host byte order); entries under meandata.mask get
        the fill value (:288-291)."""
        n = mean_rec.shape[0]
        rec = np.empty((n, 2, self.nlat, self.nlon), dtype=">f4")
        rec[:, 0] = mean_rec
        rec[:, 1] = sigma_rec
        rec[:, :, self.mask] = self.fill
        self.write_records(offset, rec.reshape(-1).view(np.uint8), n)

    def sync(self):
        """Make the file valid as it stands (record count in the header, no trailing
        preallocation): for callers that append batch by batch and may never call close()."""
        if self.fd is not None:
            os.pwrite(self.fd, struct.pack(">i", self.numrecs), 4)
            os.ftruncate(self.fd, self.rec_start + self.numrecs * self.recsize)

    def close(self, finalize=True):
        """``finalize=False``: just drop the descriptor -- for the ranks of a sharded sweep that
        wrote their record ranges into a file whose header another rank owns."""
        if self.fd is not None:
            if finalize:
                os.pwrite(self.fd, struct.pack(">i", self.numrecs), 4)
                os.ftruncate(self.fd, self.rec_start + self.numrecs * self.recsize)
            os.close(self.fd)
            self.fd = None

    ds = property(lambda self: self.fd)


class ReconWriter:
    """Output file of the reconstruction sweep: same layout as ``savesample``
    (DINCAE/__init__.py:251-278): dims time (unlimited), lon, lat; f4 variables
    lon(lon), lat(lat), meandata(lat,lon), mean_rec(time,lat,lon), sigma_rec(time,lat,lon),
    the last three with _FillValue -9999.  Unlike the reference the file stays open between
    batches.  NETCDF4 through netCDF4-python when it is installed (like the reference);
    otherwise NetCDF-3 through ``RecordFileWriter``."""

    def __new__(cls, fname, lon, lat, meandata, fill_value=-9999., append=False, expect_records=0):
        if _nc4 is None:
            return RecordFileWriter(fname, lon, lat, meandata, fill_value, append, expect_records)
        return super().__new__(cls)

    def __init__(self, fname, lon, lat, meandata, fill_value=-9999., append=False,
                 expect_records=0):  # pragma: no cover
        import threading
        self._nc4_lock = threading.Lock()
        self.fname = fname
        self.fill = np.float32(fill_value)
        self.mask = np.ma.getmaskarray(np.ma.asarray(meandata)).reshape(len(lat), len(lon))
        if append:                      # reference: Dataset(fname, 'a') for ii > 0 (:281-283)
            self.ds = _nc4.Dataset(fname, "a")
            self.v_mean = self.ds.variables["mean_rec"]
            self.v_sigma = self.ds.variables["sigma_rec"]
            return
        md = np.ma.filled(np.ma.asarray(meandata), fill_value).reshape(len(lat), len(lon))
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

    def write(self, offset, mean_rec, sigma_rec):  # pragma: no cover
        """mean_rec / sigma_rec [n,lat,lon]; entries under meandata.mask get the fill value
        (:288-291).  netCDF-C / HDF5 are not thread-safe and netCDF4-python drops the GIL inside
        nc_put_vara: writes from the writer threads are serialised here."""
        n = mean_rec.shape[0]
        m = np.broadcast_to(self.mask, mean_rec.shape)
        with self._nc4_lock:
            self.v_mean[offset:offset + n] = np.ma.masked_array(np.asarray(mean_rec, dtype="f4"), m)
            self.v_sigma[offset:offset + n] = np.ma.masked_array(np.asarray(sigma_rec, dtype="f4"), m)

    def sync(self):  # pragma: no cover
        with self._nc4_lock:
            if self.ds is not None:
                self.ds.sync()

    def write_records(self, offset, records, n=None):  # pragma: no cover
        rec = np.frombuffer(memoryview(records).cast("B"), dtype=">f4").reshape(
            -1, 2, self.mask.shape[0], self.mask.shape[1])
        n = rec.shape[0] if n is None else n
        m = np.broadcast_to(self.mask, (n,) + self.mask.shape)
        with self._nc4_lock:
            self.v_mean[offset:offset + n] = np.ma.masked_array(rec[:n, 0].astype("f4"), m)
            self.v_sigma[offset:offset + n] = np.ma.masked_array(rec[:n, 1].astype("f4"), m)

    def close(self):  # pragma: no cover
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
