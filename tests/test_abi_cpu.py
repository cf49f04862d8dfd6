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


def test_ctypes_structs_have_the_layout_of_the_header(tmp_path):
    """sizeof / offsetof of every struct that crosses the boundary, as a C compiler sees include/mgta_cuda.h, against the
    ctypes mirror in cabi.py (a silent mismatch would shift every field after it)"""
    import subprocess
    structs = {"mgta_opts": cabi.Opts, "mgta_stage_stats": cabi.StageStats, "mgta_collective": cabi.Collective,
               "mgta_sdbg_header_t": cabi.SdbgHeader}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mgta_cuda.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for field, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field, name, field))
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (name, field)]) == getattr(cls, field).offset, (name, field)
    assert cabi.COLL_ALL_REDUCE_SUM_U64 == 4 and cabi.SDBG_TIP_MAJOR == 23 and cabi.SDBG_TIP_SEQ == 8


def test_every_entry_point_with_arguments_has_its_argument_types_declared():
    """without argtypes ctypes passes Python ints as C int: a 64-bit count or pointer would be truncated silently"""
    lib = cabi.load()
    no_args = {"mgta_abi_version", "mgta_tools_last_error"}
    missing = [n for n in cabi.EXPORTS if n not in no_args and getattr(lib, n).argtypes is None]
    assert not missing, missing
    assert lib.mgta_words_per_key(1, 31) == 3 and lib.mgta_words_per_key(2, 31) == 3 and lib.mgta_words_per_key(2, 99) == 7


class _StubLib:
    """stands in for libmgta_cuda.so in the two tests below: delivers two batches to the sink like mgta_stage2 does"""

    def __init__(self, fail_rc=-1):
        self.seen_rc = None

    def mgta_last_error(self, h):
        return b"sink aborted"

    def mgta_stage2(self, h, cb, user, totals):
        import numpy as np
        data = (ctypes.c_ubyte * 8)(*range(8))
        meta = np.zeros((3, 3), dtype=np.int64)
        meta[:, 0] = [1, 0, 3]
        ptr = ctypes.cast(data, ctypes.c_void_p).value
        for b0, b1, off, n in ((0, 2, 0, 2), (2, 3, 2, 6)):
            m = np.ascontiguousarray(meta[b0:b1])
            rc = cb(None, b0, b1, ptr + off, n, m.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
            if rc != 0:
                self.seen_rc = rc
                return -1
        return 0


def _stub_context(lib):
    ctx = cabi.Context.__new__(cabi.Context)
    ctx.lib, ctx.h = lib, ctypes.c_void_p(1)
    return ctx


def test_python_sink_collects_the_deliveries():
    stream, meta, totals = _stub_context(_StubLib()).stage2()
    assert stream == bytes(range(8)) and meta[:3, 0].tolist() == [1, 0, 3] and int(meta.sum()) == 4
    n, meta, _ = _stub_context(_StubLib()).stage2(collect="count")
    assert n == 8 and meta[:3, 0].tolist() == [1, 0, 3]


def test_python_sink_failure_aborts_the_stage_and_surfaces(monkeypatch):
    """an exception inside the ctypes callback must not be swallowed (ctypes would print it and return 0: a silently
    truncated stream): the sink reports failure to the library and the exception is raised from stage2()"""
    lib = _StubLib()
    calls = []

    def broken(ptr, size):
        calls.append(size)
        raise MemoryError("no room for the delivery")

    monkeypatch.setattr(cabi.ctypes, "string_at", broken)
    with pytest.raises(MemoryError):
        _stub_context(lib).stage2()
    assert lib.seen_rc == -1 and calls == [2]                       # the library saw the abort after the first delivery


class _StubSharded(_StubLib):
    def mgta_sharded_begin(self, h, stage, cb, user):
        self.cb = cb
        return 0

    def mgta_sharded_step(self, h, c):                               # one step: the stage-2 deliveries, then "finished" (op stays 0)
        return self.mgta_stage2(h, self.cb, None, None)

    def mgta_sharded_result(self, h, ec, totals):
        return 0


def test_python_sharded_sink_collects_and_surfaces_failures(monkeypatch):
    ctx = _stub_context(_StubSharded())
    stream, meta, totals = ctx.sharded(2, run_collective=None)
    assert stream == bytes(range(8)) and meta[:3, 0].tolist() == [1, 0, 3]
    n, meta, _ = ctx.sharded(2, run_collective=None, collect="count")
    assert n == 8

    def broken(ptr, size):
        raise MemoryError("no room for the delivery")

    monkeypatch.setattr(cabi.ctypes, "string_at", broken)
    with pytest.raises(MemoryError):
        ctx.sharded(2, run_collective=None)
