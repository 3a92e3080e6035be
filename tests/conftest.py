import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    return _oracle.Oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.json")
    with open(path) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def ref():
    import _oracle
    if not _oracle.Ref.available():
        pytest.skip("oracle/_ref/libbtlref.so not built (reference tree absent)")
    return _oracle.Ref()
