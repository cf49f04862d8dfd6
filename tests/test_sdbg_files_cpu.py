"""CPU: Sdbg.from_files puts the bucket ranges of <p>.sdbg.<i> back into bucket order and hands them over in deliveries
that tile [0, 65536) -- checked on files written by the unmodified reference with 4 writer threads (its buckets are
scattered over the files), with the device calls replaced by a recorder."""
import numpy as np
import pytest

from megagta_b200 import cabi, sdbg_io
from oracle import oracle as O


class Recorder(cabi.Sdbg):
    def __init__(self):
        self.calls, self.data = [], []

    def append(self, b0, b1, data, meta):
        self.calls.append((b0, b1, int(np.asarray(meta)[:, 0].sum())))
        self.data.append(bytes(data))


def test_from_files_delivers_the_bucket_ordered_stream(read_lib, tmp_path):
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    prefix, _ = read_lib("smoke")
    out = str(tmp_path / "ref")
    O.run_ref_buildgraph(prefix, out, 31, 2, threads=4)
    hdr, stream, meta = sdbg_io.canonical(out)
    for delivery in (1 << 12, 1 << 29):
        r = Recorder()
        assert r.from_files(out, delivery_bytes=delivery)["num_threads"] == hdr["num_threads"]
        assert r.calls[0][0] == 0 and r.calls[-1][1] == 65536
        assert all(a[1] == b[0] for a, b in zip(r.calls[:-1], r.calls[1:]))
        assert b"".join(r.data) == stream and sum(c[2] for c in r.calls) == hdr["total_size"]
    assert len(r.calls) == 1
