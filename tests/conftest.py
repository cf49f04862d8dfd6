import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(datasets.GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def data_dir(tmp_path_factory):
    d = os.environ.get("MGTA_TEST_DATA") or str(tmp_path_factory.mktemp("mgta_data"))
    os.makedirs(d, exist_ok=True)
    return d


_READ_CACHE = {}


@pytest.fixture(scope="session")
def read_lib(data_dir, golden):
    """read_lib(name) -> (prefix, reads dict as loaded by the ORACLE's own numpy loader); md5-checked."""
    from oracle import oracle as O

    def get(name):
        if name not in _READ_CACHE:
            prefix = datasets.materialise(name, data_dir)
            assert datasets.md5(prefix + ".bin") == golden["datasets"][name], \
                "dataset %s drifted from the one the goldens were made on" % name
            _READ_CACHE[name] = (prefix, O.load_read_lib(prefix))
        return _READ_CACHE[name]
    return get
