"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every entry point
include/mgta_cuda.h declares (and the ctypes binding knows all of them), and refuses to work without a device
instead of falling back to anything."""
import ctypes
import os
import re

import pytest

from megagta_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    txt = open(os.path.join(ROOT, "include", "mgta_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgta_[a-z0-9_]+)\s*\(", txt)) - {"mgta_bucket_sink"})


def test_library_exports_every_declared_entry_point():
    lib = ctypes.CDLL(cabi.LIB_PATH)
    names = declared()
    assert len(names) == 36
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    assert sorted(cabi.EXPORTS) == declared()


def test_no_device_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cabi.MgtaError) as e:
        cabi.Context(31, 2)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(cabi.MgtaError):                            # the index builder has no CPU path either
        cabi.Sdbg(31)
