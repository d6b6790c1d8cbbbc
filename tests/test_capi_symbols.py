"""The C-ABI library must load and export every symbol include/dincae_b200.h declares
(no compute calls here: this runs without a GPU)."""
import ctypes
import os
import re

from dincae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dincae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dincae_[a-z0-9_]+)\s*\(", text)))


def test_library_built_and_loads():
    from dincae_b200 import build
    build.build()
    lib = _lib.load()
    assert b"sm_100a" in lib.dincae_version()
    assert lib.dincae_error_string(-3) == b"workspace too small"


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms


def test_argument_counts_match_header():
    text = open(os.path.join(ROOT, "include", "dincae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        body = m.group(1).strip()
        n = 0 if body in ("", "void") else body.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_sizing_helpers_run_on_cpu():
    lib = _lib.load()
    assert lib.dincae_adam_workspace(1000) >= 16 + 8
    assert lib.dincae_head_loss_workspace(50, 112, 112) > 0
    assert lib.dincae_dense_workspace(50, 2744, 532) >= 50 * 532 * 4
    assert lib.dincae_conv3x3_wgrad_workspace(3, 0, 16, 50, 112, 112) > 0


def test_kernels_are_compiled_for_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
