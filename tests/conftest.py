import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def synth_frames():
    """Frames 1, 2 and 3 of synthetic stream 0 (grey 480x640 uint8), their depth maps and poses."""
    from ygz_slam_b200 import synth
    out = [synth.stream_frame(k) for k in (1, 2, 3)]
    return out


@pytest.fixture(scope="session")
def ctx3():
    """GPU context with the reference's default 3-level pyramid."""
    from ygz_slam_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx8():
    """GPU context with the 8-level pyramid of BASELINE config C2."""
    from ygz_slam_b200 import Context
    c = Context(0, n_levels=8)
    yield c
    c.close()
