"""CPU-side checks of the drop-in boundary: the libraries load and export every symbol
include/*.h declares; without a GPU the computing entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpth?_[a-z0-9_]+)\s*\(", text)))


def test_libtpt_exports_every_declared_symbol(tpt):
    lib = tpt.lib()
    names = declared("tpt.h")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libtpt.so does not export " + n
    assert lib.tpt_abi_version() == 1


def test_libtpt_host_exports_every_declared_symbol(tpt):
    h = tpt.host()
    names = [n for n in declared("tpt_host.h") if n.startswith("tpth_")]
    assert len(names) == 6
    for n in names:
        assert hasattr(h, n), "libtpt_host.so does not export " + n


def test_no_oracle_symbols_in_product(tpt):
    """The product must not link or embed the checker."""
    import subprocess
    for so in (tpt.LIBTPT, tpt.LIBHOST):
        out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
        assert "orc_" not in out and "ref_scene" not in out
        ldd = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "liboracle" not in ldd and "libtptref" not in ldd


def test_invalid_arguments_are_reported(tpt):
    lib = tpt.lib()
    handle = C.c_void_p()
    assert lib.tpt_scene_create(None, 0, C.byref(handle)) == 1
    assert b"null" in lib.tpt_last_error()
    d = tpt.SceneDesc()
    assert lib.tpt_scene_create(C.byref(d), 0, C.byref(handle)) == 1
    assert lib.tpt_scene_destroy(None) == 0


def test_unknown_scene_and_missing_models(tpt, tmp_path):
    with pytest.raises(tpt.TptError):
        tpt.HostScene("no-such-scene", 16, 16)
    with pytest.raises(tpt.TptError):
        tpt.HostScene("standard", 16, 16, models_dir=str(tmp_path / "nothing"))


@pytest.mark.skipif("__import__('tpt_b200').device_count() > 0")
def test_no_cpu_fallback_without_a_device(tpt):
    """On a box without a GPU the product refuses to compute instead of falling back."""
    hs = tpt.HostScene("standard", 32, 32)
    with pytest.raises(tpt.TptError, match="no CUDA device"):
        tpt.Scene(hs.desc)
    with pytest.raises(tpt.TptError, match="no CUDA device"):
        tpt.rng(1, 4)


@pytest.mark.gpu
def test_read_bandwidth_probe(tpt):
    """tpt_probe_read_bandwidth (the L2 / HBM ceilings bench.py reports): an L2-resident buffer streams faster
    than one far larger than the L2, and both are in the range a B200 can have."""
    l2 = tpt.probe_read_bandwidth(48 << 20, 20)
    hbm = tpt.probe_read_bandwidth(2 << 30, 2)
    assert 1000.0 < hbm < 9000.0, hbm
    assert l2 > hbm, (l2, hbm)
