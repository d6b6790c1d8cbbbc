"""Shared helpers for the GPU parity tests (oracle = torch CPU / numpy restatements)."""
import datetime

import numpy as np
import torch

from dincae_b200 import layout

T0 = datetime.datetime(1900, 1, 1)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def p4(x_nhwc):
    return dev(layout.nhwc_to_p4(np.asarray(x_nhwc, dtype=np.float32)))


def from_p4(t, C):
    return layout.p4_to_nhwc(t.detach().cpu().numpy(), C)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def random_cube(rs, T, H, W, nf, frac=0.4):
    time = [T0 + datetime.timedelta(days=int(300 + 3 * i)) for i in range(T)]
    lon = np.sort(rs.uniform(-10, 10, W))
    lat = np.sort(rs.uniform(30, 40, H))
    fields = []
    for k in range(nf):
        d = (rs.standard_normal((T, H, W)) * 3 + 10).astype("f4")
        m = rs.uniform(size=(T, H, W)) < frac
        m[:, 0, 0] = True
        fields.append(np.ma.masked_array(d, m))
    return lon, lat, time, fields


def tf32_round(a):
    """numpy/torch float32 -> nearest TF32-representable float32 (10 mantissa bits, ties away
    from zero) == the device's to_tf32() / cvt.rna.tf32.f32."""
    is_t = isinstance(a, torch.Tensor)
    x = a.detach().cpu().numpy() if is_t else np.asarray(a)
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return torch.from_numpy(r.copy()) if is_t else r
