import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def restate():
    from oracle.ref import Restate
    return Restate()


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r
    if not r.have_ref():
        pytest.skip("oracle/_ref/libplade_ref.so not available")
    return r.Ref()


@pytest.fixture(scope="session")
def ctx():
    import plade_b200
    c = plade_b200.Context()      # raises if the CUDA library or a device is missing: no fallback
    yield c
    c.close()


@pytest.fixture(scope="session")
def poly_pair():
    return dict(np.load(os.path.join(GOLDEN, "polyhedron_pair.npz")))


@pytest.fixture(scope="session")
def poly_stages():
    return dict(np.load(os.path.join(GOLDEN, "polyhedron_stages.npz")))


@pytest.fixture(scope="session")
def synth_stages():
    return dict(np.load(os.path.join(GOLDEN, "synth_small_stages.npz")))


_SYNTH_SMALL = {}


def synth_small_tgt():
    """the target cloud of tests/golden/synth_small_stages.npz (tests/golden/make_golden.py: make_pair(200000, 20, seed 11))"""
    if "tgt" not in _SYNTH_SMALL:
        from plade_b200.synth import make_pair
        _SYNTH_SMALL["tgt"] = make_pair(n_points=200000, n_planes=20, seed=11)[0]
    return _SYNTH_SMALL["tgt"]
