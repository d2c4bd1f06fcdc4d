import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def weights_file(tmp_path_factory):
    """Seeded synthetic weights (oracle/weights.py) written once per session; shared by oracle and engine."""
    from oracle import weights
    p = os.environ.get("DV_WEIGHTS")
    if p and os.path.exists(p):
        return p
    d = tmp_path_factory.mktemp("w")
    path = str(d / "synth.dvw")
    weights.save_weights(path, weights.synth_all())
    return path


@pytest.fixture(scope="session")
def all_weights(weights_file):
    from oracle import weights
    return weights.load_weights(weights_file)
