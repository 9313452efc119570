import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


GOLDEN_CASES = ["joint48k_128", "joint48k_64", "indep48k_128", "joint44k_default", "indep48k_64", "wav_harps44k",
                "wav_speech44k"]      # the last two: cuts from the reference's own training WAVs (oracle/make_golden.py)
SWITCHED_CASES = ["switched48k_128", "switched44k_default"]
