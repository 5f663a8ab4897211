import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mm():
    return ge.load_package()


@pytest.fixture(scope="session")
def synth():
    return ge.load_synth()


@pytest.fixture(scope="session")
def orc():
    o = ge.load_oracle()
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx(mm):
    """GPU context; fails loudly (no CPU fallback) if the extension or the device is missing."""
    c = mm.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def scene(synth):
    """Shared synthetic scene: VLP-16 + Horizon scans at a known pose and the feature map."""
    T_true = synth.make_T(synth.rot_z(0.3), np.array([-3.0, -1.0, 0.2]))
    vx, vring, vs = synth.vlp16_scan(T_true, seed=1001)
    hx, hline, hs = synth.horizon_scan(T_true, 24000, seed=1002)
    ms, mc = synth.feature_map(100_000, 5_000, seed=1002)
    return dict(T_true=T_true, vlp=(vx, vring, vs), hori=(hx, hline, hs), map_surf=ms, map_corner=mc)
