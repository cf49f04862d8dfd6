"""Stage 2 with and without the node pass (node_kernels.cuh): the same records, a third of the items.  The switch is
read at context creation (MGTA_NODE_PASS, the A/B hook of mgta_ctx_create)."""
import os

import numpy as np
import pytest

from megagta_b200 import cabi

pytestmark = pytest.mark.gpu


def build(rd, k, m, node_pass, **kw):
    os.environ["MGTA_NODE_PASS"] = "1" if node_pass else "0"
    try:
        with cabi.Context(k, m, **kw) as ctx:
            ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
            ec = ctx.stage1() if m > 1 else None
            stream, meta, totals = ctx.stage2()
            again = ctx.stage2()                       # a second stage 2 on the same edges must not add the tips twice
            return dict(ec=ec, stream=stream, meta=meta, totals=totals, again=again, stats=ctx.stats(2))
    finally:
        os.environ.pop("MGTA_NODE_PASS", None)


@pytest.mark.parametrize("ds,k,m,kw", [("smoke", 31, 2, {}), ("meta200k", 31, 2, {}), ("meta200k", 61, 2, {}), ("adversarial", 32, 1, {}),
                                       ("adversarial", 16, 2, {}), ("smoke", 99, 2, {}), ("xander", 44, 2, {}),
                                       ("meta200k", 31, 2, {"sort_items_cap": 96}), ("meta200k", 21, 3, {"hbm_budget_bytes": 40 << 20})])
def test_node_pass_gives_the_same_records_from_fewer_items(read_lib, ds, k, m, kw):
    _, rd = read_lib(ds)
    a = build(rd, k, m, True, **kw)
    b = build(rd, k, m, False, **kw)
    assert a["stream"] == b["stream"]
    assert np.array_equal(a["meta"], b["meta"]) and np.array_equal(a["totals"], b["totals"])
    assert a["again"][0] == a["stream"] and np.array_equal(a["again"][1], a["meta"])
    assert b["stats"]["n_tip_items"] == 0 and b["stats"]["n_node_ops"] == 0
    assert a["stats"]["n_node_ops"] > 0
    assert a["stats"]["n_items"] < b["stats"]["n_items"]
    if ds == "meta200k" and not kw:
        assert a["stats"]["n_items"] * 2 < b["stats"]["n_items"]       # one item per record instead of ~3
        assert a["stats"]["n_tip_items"] == 2 * int(a["meta"][:, 1].sum())   # every tip k-mer: one $-in and one $-out item


def test_stage2_output_delivered_in_parts_behind_the_sort(golden, read_lib):
    """With a sink, a batch is sorted and emitted in a few launches and the bytes of one part travel to the host while the
    next part is sorted (from the second delivery on, once the pinned staging buffer exists).  Same bytes either way."""
    g = golden["cases"]["meta1m_k31_m2"]
    _, rd = read_lib(g["dataset"])
    from oracle import oracle as O
    with cabi.Context(g["k"], g["m"]) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        ctx.stage1()
        first = ctx.stage2()
        launches_first = ctx.stats(2)["n_launches"]
        second = ctx.stage2()                              # pipelined delivery
        nbytes, meta3, _ = ctx.stage2(collect="count")
        dev_only = ctx.stage2(collect=False)
        launches_dev = ctx.stats(2)["n_launches"]
    assert O.stream_hash(first[0]) == g["stream_hash"] and O.meta_hash(first[1]) == g["meta_hash"]
    assert second[0] == first[0] and np.array_equal(second[1], first[1]) and np.array_equal(second[2], first[2])
    assert nbytes == len(first[0]) and np.array_equal(meta3, first[1])
    assert np.array_equal(dev_only[2], first[2])
    assert launches_first > launches_dev                   # several sort launches with a sink, one without
