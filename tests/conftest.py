import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference mounted (build container only)")
    config.addinivalue_line("markers", "slow: larger transforms (still seconds on a B200)")


def golden_full_names():
    return sorted(f[len("full_"):-len(".npz")] for f in os.listdir(GOLDEN_DIR)
                  if f.startswith("full_") and f.endswith(".npz"))


def load_golden_full(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"full_{name}.npz"))
    g = {k: z[k] for k in z.files}
    g["error"] = json.loads(str(g["error"]))
    g["sample_rate_in"] = int(g["sample_rate_in"])
    g["lpm"] = int(g["lpm"])
    if "start_frame" in g:
        g["start_frame"] = int(g["start_frame"])
    return g


def load_digests():
    with open(os.path.join(GOLDEN_DIR, "digests.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def digests():
    return load_digests()
