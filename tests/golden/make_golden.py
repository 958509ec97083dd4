#!/usr/bin/env python
"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

Inputs: the only PAF fixtures the reference ships for this path
(reference tests/test_data/crush/c4_fragments/*.paf and
tests/test_data/crush/top_flubble_seqwish_minrun.paf) are copied
verbatim as test DATA. Outputs: the oracle's BED / BEDPE / PAF text for a fixed
set of queries on each fixture (golden.json), so the GPU box — which has no
/root/reference — can check both the oracle and the CUDA path against them.
BASELINE config 1 is the first query on short_floor.paf.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

REF = "/root/reference/tests/test_data/crush"
FIXTURES = ["c4_fragments/bounded_multi_bubble.paf", "c4_fragments/duplicated_repeat.paf",
            "c4_fragments/easy_shared_flank.paf", "c4_fragments/short_floor.paf", "top_flubble_seqwish_minrun.paf"]


def queries_for(idx):
    qs = []
    for sid in range(idx.n_seqs):
        L = idx.seq_len(sid)
        qs.append((sid, 0, L))
        if L >= 40:
            qs.append((sid, L // 4, L // 4 + L // 3))
    return qs


def case(idx, sid, s, e, mode, depth, d, fmt, text):
    """Full text is kept for small outputs; larger ones are pinned by sha256 + line count."""
    c = {"seq": idx.seq_name(sid), "start": s, "end": e, "mode": mode, "max_depth": depth, "d": d, "format": fmt,
         "lines": text.count("\n"), "sha256": hashlib.sha256(text.encode()).hexdigest()}
    if len(text) <= 700:
        c["text"] = text
    return c


def main():
    golden = {}
    for rel in FIXTURES:
        name = os.path.basename(rel)
        dst = os.path.join(HERE, name)
        shutil.copyfile(os.path.join(REF, rel), dst)
        idx = O.Index.from_paf(dst)
        cases = []
        for (sid, s, e) in queries_for(idx):
            region = f"{idx.seq_name(sid)}:{s}-{e}"
            for mode, depth in ((O.MODE_QUERY, 1), (O.MODE_BFS, 2), (O.MODE_BFS, 0)):
                for d in (0, 50, -1):
                    p = O.make_params(mode=mode, max_depth=depth, min_transitive_len=0, merge_distance=d)
                    bed = idx.format(idx.perform_query(sid, s, e, p), "bed", region, d)
                    cases.append(case(idx, sid, s, e, mode, depth, d, "bed", bed))
                pc = O.make_params(mode=mode, max_depth=depth, min_transitive_len=0, merge_distance=0, store_cigar=True)
                for fmt in ("bedpe", "paf"):
                    res = idx.perform_query(sid, s, e, pc)
                    res.drop_first()
                    cases.append(case(idx, sid, s, e, mode, depth, 0, fmt, idx.format(res, fmt, region, 0)))
        golden[name] = cases
        print(name, idx.n_seqs, "seqs", idx.n_records, "records", len(cases), "cases")
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
