import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds")


@pytest.fixture(scope="session")
def blob_bytes():
    # Byte-for-byte copy of the reference's bundled `blob` (ASCII "0x" + hex, 262146 bytes),
    # which src/commit.rs:30 feeds raw through include_bytes!.
    with open(os.path.join(GOLDEN_DIR, "blob"), "rb") as f:
        data = f.read()
    assert len(data) == 262146
    return data


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(GOLDEN_DIR, "vectors.json")) as f:
        return json.load(f)
