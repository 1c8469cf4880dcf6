import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
HOSTEMU = os.path.join(ROOT, "tests", "hostemu", "libbauklank_stretch_hostemu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the shared objects exist (no-op when they are up to date)."""
    import __graft_entry__ as ge
    ge.build()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        meta = json.load(f)
    sub = np.load(os.path.join(GOLDEN, "subsamples.npz"))
    return meta, sub


def sha(y):
    return hashlib.sha256(np.ascontiguousarray(y, np.float32).tobytes()).hexdigest()


def assert_matches_golden(name, y, golden, exact=True):
    """Bit-exact: sha256 of the f32 bytes equals the reference's.  On mismatch report where, using the subsample."""
    meta, sub = golden
    m = meta[name]
    assert list(y.shape) == m["shape"], (name, y.shape, m["shape"])
    if sha(y) == m["sha256"]:
        return
    ref = sub[name]
    got = np.ascontiguousarray(y[:, ::97])
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    bad = np.argwhere(got.view(np.uint32) != ref.view(np.uint32))
    first = (int(bad[0][0]), int(bad[0][1]) * 97) if len(bad) else None
    msg = "%s differs from the reference: first differing subsample (ch, n)=%s, max|err| on subsample %.3g" % (name, first, d.max())
    if exact:
        raise AssertionError(msg)
    assert d.max() <= 1e-4, msg
