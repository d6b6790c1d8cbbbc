"""Import the UNMODIFIED reference package with its two missing dependencies stubbed.

Test infrastructure only (see oracle/__init__.py).  Works only where
``/root/reference`` exists (the build container); ``available()`` says so.

The reference's only import-time use of TensorFlow is the default argument
``tf.image.ResizeMethod.NEAREST_NEIGHBOR`` (DINCAE/__init__.py:306), so a MagicMock
module is enough to let ``import DINCAE`` succeed; ``data_generator`` /
``data_generator_list`` / ``datagen`` (DINCAE/__init__.py:120-230) are pure numpy and
then run as published.
"""
import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = "/root/reference"
_cached = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "DINCAE", "__init__.py"))


def load():
    """Return the reference ``DINCAE`` module (loaded under a private module name)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "netCDF4")}
    try:
        sys.modules["tensorflow"] = MagicMock()
        fake_nc = types.ModuleType("netCDF4")
        fake_nc.Dataset = None
        fake_nc.num2date = None
        sys.modules["netCDF4"] = fake_nc
        spec = importlib.util.spec_from_file_location(
            "_dincae_reference", os.path.join(REFERENCE_ROOT, "DINCAE", "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached = mod
    return mod
