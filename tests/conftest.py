import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def synth_dir(tmp_path_factory):
    """Small synthetic PP families shared by the tests (deterministic, see locarna_b200/synth.py)."""
    import numpy as np
    from locarna_b200 import synth

    d = tmp_path_factory.mktemp("synth")
    fam = {
        "cfg2": synth.make_family(str(d / "cfg2"), 2, 2, 120),
        "cfg3": synth.make_family(str(d / "cfg3"), 3, 8, lambda rng: int(np.clip(round(rng.normal(100, 15)), 60, 140)), related=True),
        "short": synth.make_family(str(d / "short"), 7, 6, lambda rng: int(rng.randint(20, 60))),
    }
    return fam
