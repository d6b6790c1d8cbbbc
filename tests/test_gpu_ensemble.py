"""Ensemble averaging of saved reconstructions vs the numpy oracle:
fill values are skipped per member, sums are carried in float64 on the device."""
import numpy as np
import pytest

from dincae_b200 import ensemble, ncio
from oracle import ensemble_np

pytestmark = pytest.mark.gpu


def members(rs, K, shape, frac):
    ms, ss = [], []
    never = rs.uniform(size=shape[1:]) < 0.1            # masked in every member (land)
    for _ in range(K):
        m = (rs.standard_normal(shape) * 2 + 15).astype("f4")
        s = (rs.uniform(0.1, 1.5, size=shape)).astype("f4")
        bad = (rs.uniform(size=shape) < frac) | never
        ms.append(np.ma.masked_array(m, bad))
        ss.append(np.ma.masked_array(s, bad))
    return ms, ss, never


@pytest.mark.parametrize("mode", ["mean_var", "mixture"])
@pytest.mark.parametrize("K,shape,frac", [(1, (3, 5, 7), 0.0), (4, (6, 17, 13), 0.3), (9, (2, 112, 112), 0.05)])
def test_accumulator_matches_oracle(mode, K, shape, frac):
    rs = np.random.RandomState(K * 7 + shape[1])
    ms, ss, never = members(rs, K, shape, frac)
    acc = ensemble.EnsembleAccumulator(int(np.prod(shape)))
    for m, s in zip(ms, ss):
        acc.add(m, s)
    mean, sigma = acc.result(mode)
    mean = mean.cpu().numpy().reshape(shape)
    sigma = sigma.cpu().numpy().reshape(shape)
    want_m, want_s = ensemble_np.average(ms, ss, mode)
    gone = np.ma.getmaskarray(want_m)
    assert np.array_equal(mean == ensemble.FILL, gone)
    assert np.array_equal(sigma == ensemble.FILL, gone)
    keep = ~gone
    # float64 sums rounded once to float32 on both sides
    assert np.allclose(mean[keep], np.ma.getdata(want_m)[keep], rtol=1e-6, atol=0)
    assert np.allclose(sigma[keep], np.ma.getdata(want_s)[keep], rtol=1e-6, atol=0)


def test_average_reconstruction_files(tmp_path):
    rs = np.random.RandomState(3)
    T, H, W = 5, 9, 11
    lon, lat = np.linspace(-7, -1, W), np.linspace(33, 39, H)
    ms, ss, never = members(rs, 3, (T, H, W), 0.0)
    meandata = np.ma.masked_array(np.full((H, W), 18.0, dtype="f4"), never)
    names = []
    for k, (m, s) in enumerate(zip(ms, ss)):
        fn = str(tmp_path / ("data-%d.nc" % k))
        w = ncio.ReconWriter(fn, lon, lat, meandata)
        w.write(0, np.ma.getdata(m), np.ma.getdata(s))      # writer masks meandata.mask itself
        w.close()
        names.append(fn)
    out = str(tmp_path / "ensemble.nc")
    mean, sigma = ensemble.average_reconstructions(names, outfile=out, sigma_mode="mixture")
    want_m, want_s = ensemble_np.average(ms, ss, "mixture")
    assert np.array_equal(np.ma.getmaskarray(mean), np.broadcast_to(never, (T, H, W)))
    keep = ~np.ma.getmaskarray(mean)
    assert np.allclose(np.ma.getdata(mean)[keep], np.ma.getdata(want_m)[keep], rtol=1e-6)
    assert np.allclose(np.ma.getdata(sigma)[keep], np.ma.getdata(want_s)[keep], rtol=1e-6)
    back = ncio.read_recon(out)
    assert back["mean_rec"].shape == (T, H, W)
    assert np.array_equal(np.ma.getmaskarray(back["mean_rec"]), np.broadcast_to(never, (T, H, W)))
    assert np.allclose(np.ma.getdata(back["mean_rec"])[keep], np.ma.getdata(mean)[keep])
    assert np.array_equal(np.ma.getmaskarray(back["meandata"]), never)
