"""Ensemble averaging of the reconstructions saved during training (SURVEY.md section 8f,
rank 1).  The reference writes ``data-*.nc`` every ``save_each`` epochs
(DINCAE/__init__.py:671-689) and its README (README.md:90) advises averaging them; it has no
code for that step.  Here the members are streamed through two element-wise CUDA kernels
(double-precision running sums on the device, fill values skipped) and the result is written
with the same file layout as ``savesample`` (meandata, mean_rec, sigma_rec).

    sigma_mode="mean_var"  sigma = sqrt(mean over members of sigma_rec**2)
    sigma_mode="mixture"   sigma = sqrt(mean sigma_rec**2 + variance of the members' mean_rec)
"""
import numpy as np
import torch

from . import _lib, ncio
from ._lib import check, ptr, stream_ptr

FILL = np.float32(-9999.0)


class EnsembleAccumulator:
    """Running ensemble statistics on the device for fields of ``n`` elements."""

    def __init__(self, n, device=None, fill=FILL):
        if not torch.cuda.is_available():
            raise _lib.DincaeError("dincae_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.n = int(n)
        self.fill = float(fill)
        dev = torch.device(device if device is not None else "cuda")
        self.acc = torch.zeros(3, self.n, dtype=torch.float64, device=dev)
        self.count = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.members = 0

    def add(self, mean_rec, sigma_rec):
        """mean_rec / sigma_rec: float32 arrays (numpy, masked numpy or torch) of n elements."""
        m = self._dev(mean_rec)
        s = self._dev(sigma_rec)
        check(self.lib.dincae_ensemble_accumulate(ptr(m), ptr(s), self.fill, self.n, ptr(self.acc),
                                                  ptr(self.count), stream_ptr()), "ensemble_accumulate")
        self.members += 1

    def result(self, sigma_mode="mean_var"):
        """-> (mean, sigma) float32 torch tensors of n elements (fill where no member is valid)."""
        if sigma_mode not in ("mean_var", "mixture"):
            raise ValueError("sigma_mode must be 'mean_var' or 'mixture'")
        mean = torch.empty(self.n, dtype=torch.float32, device=self.acc.device)
        sigma = torch.empty_like(mean)
        check(self.lib.dincae_ensemble_finalize(ptr(self.acc), ptr(self.count), self.n,
                                                int(sigma_mode == "mixture"), self.fill, ptr(mean),
                                                ptr(sigma), stream_ptr()), "ensemble_finalize")
        return mean, sigma

    def _dev(self, a):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.acc.device, dtype=torch.float32).reshape(-1)
        else:
            a = np.ma.filled(np.ma.asarray(a, dtype=np.float32), self.fill)
            t = torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).to(self.acc.device)
        if t.numel() != self.n:
            raise ValueError("member has %d elements, expected %d" % (t.numel(), self.n))
        return t.contiguous()


def average_reconstructions(filenames, outfile=None, sigma_mode="mean_var"):
    """Average the ``mean_rec`` / ``sigma_rec`` of several reconstruction files (same grid and
    number of time steps).  Returns (mean_rec, sigma_rec) as masked float32 arrays
    [time, lat, lon]; with ``outfile`` the result is also written in the savesample layout."""
    filenames = list(filenames)
    if not filenames:
        raise ValueError("no reconstruction files given")
    first = ncio.read_recon(filenames[0])
    shape = first["mean_rec"].shape
    acc = EnsembleAccumulator(int(np.prod(shape)))
    for k, fn in enumerate(filenames):
        d = first if k == 0 else ncio.read_recon(fn)
        if d["mean_rec"].shape != shape:
            raise ValueError("%s: shape %s differs from %s" % (fn, d["mean_rec"].shape, shape))
        acc.add(d["mean_rec"], d["sigma_rec"])
    mean, sigma = acc.result(sigma_mode)
    mean = mean.cpu().numpy().reshape(shape)
    sigma = sigma.cpu().numpy().reshape(shape)
    masked = mean == FILL
    if outfile is not None:
        w = ncio.ReconWriter(outfile, first["lon"], first["lat"], first["meandata"], fill_value=float(FILL))
        # entries without any valid member keep the fill value; the writer also masks the
        # pixels that are never observed (meandata.mask), like savesample does
        w.write(0, mean, sigma)
        w.close()
    return np.ma.masked_array(mean, masked), np.ma.masked_array(sigma, masked)
