"""The in-memory read generator of the large GPU parity tests equals the file path the goldens were made through."""
import numpy as np

import datasets
from megagta_b200 import synth
from oracle import oracle as O


def test_metagenome_in_memory_equals_the_written_read_library(tmp_path):
    n, L = 1_050_000, 32                                   # two RNG chunks, the second one ragged
    kw = dict(n_genomes=8, glen=(20_000, 50_000))
    prefix = str(tmp_path / "m")
    synth.write_metagenome(prefix, n, L, seed=5, **kw)
    rd = O.load_read_lib(prefix)
    seq, start, md5 = datasets.metagenome_in_memory(n, L, seed=5, **kw)
    assert md5 == datasets.md5(prefix + ".bin")
    assert np.array_equal(start, rd["start"])
    assert np.array_equal(seq, rd["seq"])
