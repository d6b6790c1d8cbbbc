"""Host-side logic that needs no GPU: layouts, parameter plan, sampler, netCDF I/O, API
surface (compared with the reference's own signatures where the reference tree exists)."""
import datetime
import inspect
import os

import numpy as np
import pytest

from dincae_b200 import api, layout, ncio, sampler
from dincae_b200.engine import Plan
from oracle import model_torch as M
from oracle import ref_loader


def test_p4_round_trip_and_padding():
    rs = np.random.RandomState(0)
    x = rs.standard_normal((2, 5, 7, 10)).astype("f4")
    p = layout.nhwc_to_p4(x)
    assert p.shape == (2, 3, 5, 7, 4)
    assert np.array_equal(layout.p4_to_nhwc(p, 10), x)
    assert np.all(p[:, 2, :, :, 2:] == 0)


def test_conv_pack_round_trip_with_split_sources():
    rs = np.random.RandomState(1)
    k = rs.standard_normal((3, 3, 54 + 36, 36)).astype("f4")
    wp = layout.pack_conv(k, [54, 36], 48)
    assert wp.shape == (9, 14 + 9, 48, 4)
    assert np.array_equal(layout.unpack_conv(wp, [54, 36], 36), k)
    # channel 54 of TF (first skip channel) lands in plane 14, lane 0
    assert wp[4, 14, 7, 0] == k[1, 1, 54, 7]
    assert np.all(wp[:, 13, :, 2:] == 0) and np.all(wp[:, :, 36:, :] == 0)


def test_plan_matches_reference_architecture():
    plan = Plan(112, 112)
    arch = M.Arch(112, 112)
    assert plan.n_params_tf == 2881367
    assert [e["name"] for e in plan.entries] == [n for n, _ in arch.param_shapes()]
    assert plan.dense_units == [529, 2646]
    assert plan.sizes == [(112, 112), (56, 56), (28, 28), (14, 14), (7, 7)]
    params = [p.numpy() for p in M.init_params(arch, 0)]
    back = plan.unpack(plan.pack(params))
    assert all(np.array_equal(a, b) for a, b in zip(back, params))
    # kernels first (L2 region), biases after
    offs_w = [e["offset"] for e in plan.entries if e["kind"].endswith("_w")]
    offs_b = [e["offset"] for e in plan.entries if e["kind"].endswith("_b")]
    assert max(offs_w) < plan.n_decay <= min(offs_b)


def test_plan_odd_sizes_and_config_d():
    p = Plan(30, 22)
    assert p.sizes[-1] == (2, 2)
    d = Plan(512, 512, 28, [32, 48, 72, 108, 162], skipconnections=[1, 2, 3, 4, 5])
    assert d.n_params_tf == 688723466                  # SURVEY.md section 8d (skip on all 5 layers)
    assert Plan(512, 512, 28, [32, 48, 72, 108, 162]).n_params_tf == 688723466 - 9 * 28 * 2
    assert d.dense_units == [8294, 41472]


def test_sampler_shuffle_semantics():
    s = sampler.TrainSampler(100, 10, shuffle_buffer_size=45, seed=0)
    seen = []
    for _ in range(30):
        f, c, q = s.next_batch()
        assert f.dtype == np.int32 and c.dtype == np.int32 and q.dtype == np.int64
        assert len(f) == 10 and (c >= 0).all() and (c < 100).all()
        seen.extend(q.tolist())
    # every generated element is emitted at most once; a frame can lag at most by the buffer
    assert len(set(seen)) == len(seen)
    assert max(seen) < 300 + 45
    first = np.array(seen[:10])
    assert first.max() < 45 + 10                       # local shuffle only (reservoir of 45)
    assert list(first) != sorted(first)
    # frame index = sequence number modulo ntime (the stream repeats)
    s2 = sampler.TrainSampler(7, 5, shuffle_buffer_size=3, seed=1)
    f, c, q = s2.next_batch(20)
    assert np.array_equal(f, (q % 7).astype(np.int32))


def test_sampler_cloud_draws_follow_python_random_in_generator_order():
    import random
    random.seed(3)
    want = [random.randrange(0, 50) for _ in range(60)]
    random.seed(3)
    s = sampler.TrainSampler(50, 5, shuffle_buffer_size=10, seed=0)
    got = {}
    for _ in range(10):
        f, c, q = s.next_batch()
        for ci, qi in zip(c, q):
            got[int(qi)] = int(ci)
    for qi, ci in got.items():
        assert ci == want[qi]


def test_sequential_batches_partial_tail():
    bs = list(sampler.sequential_batches(5266, 50))
    assert len(bs) == 106 and len(bs[-1]) == 16 and bs[-1][-1] == 5265


def test_num2date_and_dayofyear():
    t = ncio.num2date(np.array([0, 59, 60, 365.5]), "days since 1900-01-01 00:00:00")
    assert t[0] == datetime.datetime(1900, 1, 1) and t[2] == datetime.datetime(1900, 3, 2)
    assert t[3] == datetime.datetime(1901, 1, 1, 12)
    t = ncio.num2date([24], "hours since 2000-02-28")
    assert t[0] == datetime.datetime(2000, 2, 29)


def test_input_file_round_trip(tmp_path):
    from dincae_b200 import synthetic
    lon, lat, time, data, mask = synthetic.sst_like_cube(T=12, H=10, W=14, seed=3)
    f = str(tmp_path / "in.nc")
    synthetic.write_input_file(f, "SST", lon, lat, time, data, mask)
    lon2, lat2, time2, data2, missing, mask2 = api.load_gridded_nc(f, "SST")
    assert np.allclose(lon2, lon, atol=1e-5) and time2 == time
    assert np.array_equal(np.ma.getmaskarray(data2), np.ma.getmaskarray(data))
    assert np.array_equal(np.ma.getdata(data2)[~missing], np.ma.getdata(data)[~missing])
    assert np.array_equal(mask2, mask) and mask2.dtype == bool
    frac = np.ma.getmaskarray(data)[:, mask].mean()
    assert 0.3 < frac < 0.8


def test_recon_writer_layout_matches_savesample(tmp_path):
    """dims/vars/dtypes/fill of DINCAE/__init__.py:251-278; mask = meandata.mask (:288-291)."""
    H, W = 6, 5
    lon, lat = np.linspace(0, 1, W), np.linspace(10, 11, H)
    md = np.ma.masked_array(np.full((1, H, W), 15.0), np.zeros((1, H, W), bool))
    md.mask[0, 0, 0] = True
    f = str(tmp_path / "out.nc")
    m1 = np.random.RandomState(0).standard_normal((4, H, W)).astype("f4")
    s1 = np.abs(m1) + 1
    api.savesample(f, m1[:3], s1[:3], md, lon, lat, 0, 0, 0)
    api.savesample(f, m1[3:], s1[3:], md, lon, lat, 0, 1, 3)
    api._close_writer(f)
    out = ncio.read_recon(f)
    assert out["mean_rec_dims"] == ("time", "lat", "lon") and out["meandata_dims"] == ("lat", "lon")
    assert out["dims"]["time"] is None and out["dims"]["lon"] == W and out["dims"]["lat"] == H
    assert out["mean_rec_dtype"] == "float32" and out["lon_dtype"] == "float32"
    assert out["mean_rec"].shape == (4, H, W)
    assert out["mean_rec"].mask[:, 0, 0].all() and out["mean_rec"].mask.sum() == 4
    assert np.allclose(out["mean_rec"][:, 1:, :].data, (m1 + 15.0)[:, 1:, :], atol=1e-6)
    assert np.allclose(out["sigma_rec"][:, 1:, :].data, np.sqrt(s1)[:, 1:, :], atol=1e-6)
    # append to a closed file (reference re-opens in 'a' mode for ii > 0)
    api.savesample(f, m1[:1], s1[:1], md, lon, lat, 0, 2, 4)
    api._close_writer(f)
    assert ncio.read_recon(f)["mean_rec"].shape == (5, H, W)


def test_output_names_never_collide(tmp_path):
    names = {api._output_name(str(tmp_path)) for _ in range(5)}
    assert len(names) == 5
    assert all(os.path.basename(n).startswith("data-") and n.endswith(".nc") for n in names)


REF_KWARGS = dict(
    resize_method=None, epochs=1000, batch_size=30, save_each=10, save_model_each=500,
    skipconnections=[1, 2, 3, 4], dropout_rate_train=0.3, tensorboard=False, truth_uncertain=False,
    shuffle_buffer_size=45, nvar=10, enc_nfilter_internal=[16, 24, 36, 54], frac_dense_layer=[0.2],
    clip_grad=5.0, regularization_L2_beta=0, learning_rate=1e-3, iseed=None, nprefetch=0,
    nepoch_keep_missing=0)


def test_reconstruct_signature_is_the_reference_signature():
    sig = inspect.signature(api.reconstruct)
    names = list(sig.parameters)
    assert names[:9] == ["lon", "lat", "mask", "meandata", "train_datagen", "train_len",
                         "test_datagen", "test_len", "outdir"]
    for k, v in REF_KWARGS.items():
        assert k in sig.parameters, k
        if v is not None:
            assert sig.parameters[k].default == v, k
    assert sig.parameters["learning_rate_decay_epoch"].default == np.inf
    assert sig.parameters["savesample"].default is api.savesample
    assert sig.parameters["transfun"].default == (api.identity, api.identity)
    assert api.__all__ == ["reconstruct", "load_gridded_nc", "data_generator", "reconstruct_gridded_nc"]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not on this machine")
def test_public_signatures_equal_live_reference():
    ref = ref_loader.load()
    for name in ("reconstruct", "reconstruct_gridded_nc", "reconstruct_gridded_files",
                 "data_generator", "data_generator_list", "savesample", "load_gridded_nc"):
        a = inspect.signature(getattr(ref, name))
        b = inspect.signature(getattr(api, name))
        assert list(a.parameters) == list(b.parameters), name
        for p in a.parameters:
            da, db = a.parameters[p].default, b.parameters[p].default
            if p in ("resize_method", "savesample", "transfun"):
                continue
            assert (da is inspect._empty and db is inspect._empty) or da == db, (name, p)
    assert ref.__all__ == api.__all__


def test_dincae_shim_exports():
    import DINCAE
    for n in ("reconstruct", "load_gridded_nc", "data_generator", "reconstruct_gridded_nc",
              "reconstruct_gridded_files", "data_generator_list", "savesample", "identity"):
        assert callable(getattr(DINCAE, n))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dincae_b200._lib import DincaeError
    from dincae_b200.engine import Network
    with pytest.raises(DincaeError):
        Network(Plan(16, 16), 2)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d in ("dincae_b200", "DINCAE"):
        for fn in os.listdir(os.path.join(root, d)):
            if fn.endswith(".py"):
                src = open(os.path.join(root, d, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_ensemble_oracle_definitions():
    """numpy oracle of the ensemble averaging: masked members are skipped; 'mixture' adds the
    spread of the member means (law of total variance)."""
    from oracle import ensemble_np
    m1 = np.ma.masked_array(np.array([1.0, 2.0, 5.0], dtype="f4"), [False, False, True])
    m2 = np.ma.masked_array(np.array([3.0, 2.0, 7.0], dtype="f4"), [False, True, True])
    s1 = np.ma.masked_array(np.array([1.0, 2.0, 1.0], dtype="f4"), [False, False, True])
    s2 = np.ma.masked_array(np.array([1.0, 4.0, 1.0], dtype="f4"), [False, True, True])
    mu, sg = ensemble_np.average([m1, m2], [s1, s2], "mean_var")
    assert np.allclose(np.ma.getdata(mu)[:2], [2.0, 2.0]) and np.ma.getmaskarray(mu)[2]
    assert np.allclose(np.ma.getdata(sg)[:2], [1.0, 2.0])
    mu, sg = ensemble_np.average([m1, m2], [s1, s2], "mixture")
    assert np.allclose(np.ma.getdata(sg)[:2], [np.sqrt(1.0 + 1.0), 2.0])


def test_same_buffer_sees_through_fresh_mask_views():
    """np.ma returns a new view object of a field's mask on every access; DeviceCube.upload must
    recognise it as the caller's `missing` without comparing T*H*W booleans on the host."""
    from dincae_b200.data import _same_buffer
    d = np.zeros((5, 4, 3), "f4")
    m = np.zeros((5, 4, 3), bool)
    data = np.ma.masked_array(d, m)
    assert np.ma.getmaskarray(np.ma.asarray(data)) is not np.ma.getmaskarray(data)
    assert _same_buffer(np.ma.getmaskarray(np.ma.asarray(data)), np.ma.getmaskarray(data))
    assert not _same_buffer(m.copy(), m)
    assert not _same_buffer(m[1:], m[:-1])
    assert not _same_buffer(m.view(np.uint8), m)
    assert not _same_buffer([True], m)


def test_record_file_writer_is_a_valid_netcdf3_file(tmp_path):
    """The streaming writer's bytes against an independent reader (scipy.io.netcdf_file):
    dimensions, variables, dtypes, fill attributes, record interleaving, out-of-order and
    threaded batch writes, append mode (DINCAE/__init__.py:251-294 layout)."""
    from concurrent.futures import ThreadPoolExecutor
    from scipy.io import netcdf_file
    from dincae_b200 import ncio
    rs = np.random.RandomState(0)
    H, W, T, B = 7, 10, 23, 5
    lon, lat = np.linspace(-7, 0, W), np.linspace(33, 40, H)
    md = np.ma.masked_array(rs.rand(1, H, W), rs.rand(1, H, W) < 0.3)
    M = rs.standard_normal((T, H, W)).astype("f4")
    S = rs.rand(T, H, W).astype("f4")
    fname = str(tmp_path / "o.nc")
    w = ncio.RecordFileWriter(fname, lon, lat, md, expect_records=T + 9)      # over-estimate: trimmed on close
    batches = [(o, min(B, T - o)) for o in range(0, T, B)]
    with ThreadPoolExecutor(3) as pool:                       # any order, several threads
        for o, n in reversed(batches[:-1]):
            pool.submit(w.write, o, M[o:o + n], S[o:o + n])
    w.close()
    assert os.path.getsize(fname) == w.rec_start + 20 * w.recsize
    with netcdf_file(fname, "r", mmap=False) as ds:
        assert ds.version_byte == 2
        assert list(ds.dimensions.items()) == [("time", None), ("lon", W), ("lat", H)]
        assert ds.variables["mean_rec"].shape == (20, H, W)
    o, n = batches[-1]                                        # re-open like savesample(ii > 0) does
    w = ncio.RecordFileWriter(fname, lon, lat, md, append=True)
    rec = np.empty((n, 2, H, W), ">f4")
    rec[:, 0], rec[:, 1] = M[o:], S[o:]
    rec[:, :, np.ma.getmaskarray(md)[0]] = -9999.0
    w.write_records(o, rec.view(np.uint8).reshape(-1))
    w.close()
    with netcdf_file(fname, "r", mmap=False) as ds:
        for k in ("meandata", "mean_rec", "sigma_rec"):
            v = ds.variables[k]
            assert v.typecode() == "f" and v._FillValue == np.float32(-9999.0)
        assert ds.variables["mean_rec"].dimensions == ("time", "lat", "lon")
        assert ds.variables["meandata"].dimensions == ("lat", "lon")
        assert np.allclose(ds.variables["lon"][:], lon) and np.allclose(ds.variables["lat"][:], lat)
        mask = np.ma.getmaskarray(md)[0]
        for name, want in (("mean_rec", M), ("sigma_rec", S)):
            got = np.array(ds.variables[name][:])
            assert got.shape == (T, H, W)
            assert np.array_equal(got[:, ~mask], want[:, ~mask])
            assert np.all(got[:, mask] == -9999.0)
        got = np.array(ds.variables["meandata"][:])
        assert np.array_equal(got[~mask], np.ma.getdata(md)[0][~mask].astype("f4"))
        assert np.all(got[mask] == -9999.0)
    with pytest.raises(ValueError):                           # another grid: not appendable
        ncio.RecordFileWriter(fname, lon[:-1], lat, md[:, :, :-1], append=True)


def test_sampler_cloud_draws_are_randrange_draws():
    """The sampler inlines random.randrange's rejection loop; it must consume and produce the
    same stream as the call the reference makes (DINCAE/__init__.py:214)."""
    import random
    from dincae_b200 import sampler
    for T in (1, 2, 7, 64, 100, 5266):
        r1, r2 = random.Random(5), random.Random(5)
        s = sampler.TrainSampler(T, 6, shuffle_buffer_size=1, seed=0, py_random=r1)
        got = np.concatenate([s.next_batch()[1] for _ in range(20)])
        want = [r2.randrange(0, T) for _ in range(len(got) + 1)]       # +1: the refill after the last pick
        assert list(got) == want[:len(got)]
        assert r1.random() == r2.random()                               # streams stay aligned
