"""Generate tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref/megagta_ref,
built by oracle/Makefile from /root/reference/src) on the deterministic datasets of tests/datasets.py.
Run here (the container that has /root/reference):  python tests/golden/make_golden.py [case-filter]

Also packs the reference's only in-tree fixture for this path
(share/RDPTools/Xander_assembler/testdata/test_reads.fa) into tests/golden/xander.bin via the
reference's own `buildlib`.
"""
import hashlib
import json
import os
import re
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
from megagta_b200 import sdbg_io  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = []
for k, m, mercy in [(31, 2, False), (31, 2, True), (21, 2, False), (22, 2, False), (32, 2, False), (41, 2, False),
                    (61, 2, False), (99, 2, False), (31, 1, False), (31, 3, False), (27, 3, True)]:
    CASES.append(("smoke", k, m, mercy))
for k, m, mercy in [(29, 1, False), (44, 1, False), (29, 2, True), (127, 1, False), (126, 2, False), (112, 2, True)]:   # 127 = kMaxK
    CASES.append(("xander", k, m, mercy))
for k, m, mercy in [(31, 2, False), (21, 1, False), (27, 3, False), (30, 2, False), (31, 2, True), (27, 3, True),
                    (48, 2, False), (17, 2, False)]:
    CASES.append(("adversarial", k, m, mercy))
for k, m, mercy in [(21, 1, False), (25, 2, False), (25, 2, True), (10, 2, False), (11, 1, False)]:
    CASES.append(("tiny", k, m, mercy))
for k, m, mercy in [(31, 2, False), (61, 2, False), (21, 3, False)]:
    CASES.append(("meta200k", k, m, mercy))
CASES.append(("meta1m", 31, 2, False))
ASSIST_CASES = [("smoke", 31, 2), ("smoke", 31, 1), ("adversarial", 27, 3), ("meta200k", 61, 2)]   # --assist_seq


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    O.build()
    assert O.have_ref(), "reference binary missing"
    out_path = os.path.join(datasets.GOLDEN_DIR, "golden.json")
    golden = json.load(open(out_path)) if os.path.exists(out_path) else {"datasets": {}, "cases": {}}
    work = tempfile.mkdtemp(prefix="mgta_golden_")
    xfa = "/root/reference/share/RDPTools/Xander_assembler/testdata/test_reads.fa"
    if os.path.exists(xfa) and not os.path.exists(os.path.join(datasets.GOLDEN_DIR, "xander.bin")):
        libf = os.path.join(work, "x.lib")
        open(libf, "w").write("xander\nse %s\n" % xfa)
        O.run_ref_buildlib(libf, os.path.join(work, "xander"))
        shutil.copy(os.path.join(work, "xander.bin"), os.path.join(datasets.GOLDEN_DIR, "xander.bin"))
        shutil.copy(os.path.join(work, "xander.lib_info"), os.path.join(datasets.GOLDEN_DIR, "xander.lib_info"))
    for ds, k, m, mercy in CASES:
        name = "%s_k%d_m%d%s" % (ds, k, m, "_mercy" if mercy else "")
        if flt and flt not in name:
            continue
        prefix = datasets.materialise(ds, os.path.join(work, "data"))
        golden["datasets"][ds] = datasets.md5(prefix + ".bin")
        outp = os.path.join(work, name)
        log = O.run_ref_buildgraph(prefix, outp, k, m, threads=8, need_mercy=mercy)
        hdr, stream, meta = sdbg_io.canonical(outp)
        mm = re.search(r"Number mercy: (\d+)", log)
        nw = re.search(r"\]\s+((?:\d+ ){9})\s*$", log, re.M)
        entry = dict(dataset=ds, k=k, m=m, mercy=mercy, total_size=hdr["total_size"], num_tips=hdr["num_tips"],
                     large_multi=hdr["large_multi"], words_per_tip_label=hdr["words_per_tip_label"],
                     stream_bytes=len(stream), stream_hash=sdbg_io.stream_hash(stream),
                     meta_hash=sdbg_io.meta_hash(meta),
                     num_w=[int(x) for x in nw.group(1).split()] if nw else None,
                     num_mercy=int(mm.group(1)) if mm else None)
        if m > 1:
            entry["counting_sha"] = hashlib.sha256(open(outp + ".counting", "rb").read()).hexdigest()[:16]
        if mercy:
            import numpy as np
            cands = np.concatenate([np.fromfile(f, dtype="<u8") for f in
                                    sorted(os.path.join(work, x) for x in os.listdir(work)
                                           if x.startswith(name + ".mercy_cand."))] or [np.empty(0, "<u8")])
            entry["mercy_cand_n"] = int(len(cands))
            entry["mercy_cand_sha"] = hashlib.sha256(np.sort(cands).tobytes()).hexdigest()[:16]
        golden["cases"][name] = entry
        print(name, entry["total_size"], entry["stream_hash"], entry["meta_hash"], entry.get("num_mercy"), flush=True)
        for f in os.listdir(work):
            if f.startswith(name + "."):
                os.remove(os.path.join(work, f))
    for ds, k, m in ASSIST_CASES:
        name = "%s_k%d_m%d_assist" % (ds, k, m)
        if flt and flt not in name:
            continue
        prefix = datasets.materialise(ds, os.path.join(work, "data"))
        fa = datasets.assist_fasta(ds, os.path.join(work, "data"))
        golden["datasets"][ds + ".assist.fa"] = datasets.md5(fa)
        outp = os.path.join(work, name)
        log = O.run_ref_buildgraph(prefix, outp, k, m, threads=8, assist_seq=fa)
        hdr, stream, meta = sdbg_io.canonical(outp)
        nw = re.search(r"\]\s+((?:\d+ ){9})\s*$", log, re.M)
        entry = dict(dataset=ds, k=k, m=m, mercy=False, assist=True, total_size=hdr["total_size"], num_tips=hdr["num_tips"],
                     large_multi=hdr["large_multi"], words_per_tip_label=hdr["words_per_tip_label"],
                     stream_bytes=len(stream), stream_hash=sdbg_io.stream_hash(stream), meta_hash=sdbg_io.meta_hash(meta),
                     num_w=[int(x) for x in nw.group(1).split()] if nw else None, num_mercy=None)
        if m > 1:
            entry["counting_sha"] = hashlib.sha256(open(outp + ".counting", "rb").read()).hexdigest()[:16]
        golden["cases"][name] = entry
        print(name, entry["total_size"], entry["stream_hash"], entry["meta_hash"], flush=True)
        for f in os.listdir(work):
            if f.startswith(name + "."):
                os.remove(os.path.join(work, f))
    json.dump(golden, open(out_path, "w"), indent=1, sort_keys=True)
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
