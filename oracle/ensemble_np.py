"""TEST INFRASTRUCTURE ONLY (oracle): numpy statement of the ensemble averaging of saved
reconstructions.  The reference has no code for this step (README.md:90 only recommends
it), so the definition is ours: member-wise masked means in float64."""
import numpy as np


def average(members_mean, members_sigma, sigma_mode="mean_var"):
    """members_*: lists of masked float32 arrays of equal shape -> (mean, sigma) masked f32."""
    m = np.ma.stack([np.ma.asarray(a, dtype=np.float64) for a in members_mean])
    s = np.ma.stack([np.ma.asarray(a, dtype=np.float64) for a in members_sigma])
    bad = np.ma.getmaskarray(m) | np.ma.getmaskarray(s)
    m = np.ma.masked_array(np.ma.getdata(m), bad)
    s = np.ma.masked_array(np.ma.getdata(s), bad)
    mu = m.mean(axis=0)
    var = (s ** 2).mean(axis=0)
    if sigma_mode == "mixture":
        var = var + np.maximum((m ** 2).mean(axis=0) - mu ** 2, 0.0)
    return mu.astype(np.float32), np.sqrt(var).astype(np.float32)
