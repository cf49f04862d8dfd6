"""GPU parity at the size SURVEY.md Appendix C pins with the unmodified reference: 5M x 150 bp metagenome reads (`gen_bin`,
seed 20261017) at k = 21 / 31 / 41 / 61 (1-, 2-, 2- and 3-/4-word keys), min-count 2.  The reads are regenerated in memory;
the md5 of the `.bin` bytes is checked first so a drifted RNG stream fails loudly.  Compared: total_size, num_tips,
large_multi, stream length, bucket-ordered stream hash and per-bucket table hash (cx1_read2sdbg_s2.cpp:742-835 output as
SdbgWriter lays it out, sdbg_multi_io.h:83-112,160-187)."""
import json
import os

import numpy as np
import pytest

import datasets
from megagta_b200 import cabi, sdbg_io

pytestmark = pytest.mark.gpu

with open(os.path.join(datasets.GOLDEN_DIR, "appendix_c.json")) as f:
    APPENDIX_C = json.load(f)


@pytest.fixture(scope="module")
def reads_5m():
    seq, start, md5 = datasets.metagenome_in_memory(APPENDIX_C["n_reads"], APPENDIX_C["read_len"], APPENDIX_C["seed"])
    assert md5 == APPENDIX_C["bin_md5"], "the 5M x 150 read set drifted from the one Appendix C was made on"
    return seq, start


@pytest.mark.parametrize("k", [31, 21, 41, 61])
def test_gpu_matches_appendix_c_reference_hashes(reads_5m, k):
    seq, start = reads_5m
    g = APPENDIX_C["cases"][str(k)]
    with cabi.Context(k, APPENDIX_C["min_count"]) as ctx:
        ctx.set_reads(seq, start, max_len=APPENDIX_C["read_len"])
        ctx.stage1()
        stream, meta, totals = ctx.stage2()
    assert int(meta[:, 0].sum()) == g["total_size"]
    assert int(meta[:, 1].sum()) == g["num_tips"]
    assert int(meta[:, 2].sum()) == g["large_multi"]
    assert len(stream) == g["stream_bytes"]
    assert sdbg_io.stream_hash(stream) == g["stream_hash"]
    assert sdbg_io.meta_hash(meta) == g["meta_hash"]
    assert int(np.sum(totals[:9])) == g["total_size"]
