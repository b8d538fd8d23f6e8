import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_ready():
    """Is there a CUDA device?  (A device WITHOUT the built library is not a reason to skip: the tests then fail loudly.)"""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as e:  # pragma: no cover
        return False, "torch is not importable: %s" % e
    return True, ""


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a CPU box skips the gpu-marked tests (they would all stop with the library's
    "no CUDA device" error); on a GPU box nothing is skipped -- a missing library fails there -- and `-m gpu`
    without a device reports every test as skipped, never as passed."""
    if not any("gpu" in it.keywords for it in items):
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def tables():
    g = os.path.join(ROOT, "tests", "golden")
    return np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))


def norm_rel(a, b):
    """max|a-b| / max|b| -- the 1e-5 'single precision' gate of BASELINE.md sec. 4."""
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def physical(O, name, m=0):
    """Physical particles of image m in file order, read out of the oracle's *buffered* layout
    (rows cum(0,j,k)+1 .. cum(nt,j,k) of every tile, buffer_density.f90:119-141)."""
    arr = getattr(O, name)(m)
    cum = O.cum(m)
    b = O.ncb
    out = []
    for tz in range(O.nnt):
        for ty in range(O.nnt):
            for tx in range(O.nnt):
                c = cum[tz, ty, tx]
                last = c[b:b + O.nt, b:b + O.nt, b + O.nt - 1]   # cum(nt,j,k)
                first = c[b:b + O.nt, b:b + O.nt, b - 1]          # cum(0,j,k)
                for k in range(O.nt):
                    for j in range(O.nt):
                        out.append(arr[first[k, j]:last[k, j]])
    return np.concatenate(out)
