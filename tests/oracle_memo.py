"""Session memo of the oracle's answers for the CPU logic tests: four test families ask for the same (dataset, k, m) graphs.
Arrays are handed out as copies (O.mercy extends is_solid in place)."""
import hashlib

import numpy as np

from oracle import oracle as O

_S1, _S2 = {}, {}


def _rd_key(rd):
    return (id(rd["seq"]), int(rd["n_reads"]))          # read sets are cached for the session (conftest.read_lib)


def stage1(rd, k, m, need_mercy=False):
    key = _rd_key(rd) + (k, m, bool(need_mercy))
    if key not in _S1:
        _S1[key] = O.stage1(rd, k, m, need_mercy)
    solid, ec, cands = _S1[key]
    return solid.copy(), ec.copy(), cands.copy()


def stage2(rd, k, m, is_solid):
    tag = b"" if is_solid is None else hashlib.blake2b(np.ascontiguousarray(is_solid).tobytes(), digest_size=16).digest()
    key = _rd_key(rd) + (k, m, tag)
    if key not in _S2:
        _S2[key] = O.stage2(rd, k, m, is_solid)
    stream, meta, totals = _S2[key]
    return stream, meta.copy(), totals.copy()
