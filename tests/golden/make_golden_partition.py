#!/usr/bin/env python
"""Regenerates tests/golden/golden_partition.json from the fixture PAFs already in this directory: the oracle's
`partitions.bed` text (reference src/commands/partition.rs:158-712, :1682-1717) for a fixed set of `partition`
parameter sets per fixture, and the bytes of the `.impg` index file (reference src/impg.rs:1655-1720, as restated in
tests/_impg_format.py) of each fixture. The GPU box checks the CUDA path against these without the oracle's help."""
import glob
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _impg_format as F  # noqa: E402
import _oracle as O  # noqa: E402

PARAMS = [dict(window_size=100, merge_distance=0, min_missing_size=0, min_boundary_distance=0),
          dict(window_size=150, merge_distance=30, min_missing_size=40, min_boundary_distance=25, rehome_singletons=False),
          dict(window_size=80, merge_distance=1000, selection_mode="total", max_depth=0, min_transitive_len=0),
          dict(window_size=120, merge_distance=10, transitive_dfs=True, selection_mode="sample", max_depth=3),
          dict(window_size=60, merge_distance=-1, min_missing_size=10, min_boundary_distance=10, min_transitive_len=20)]


def main():
    golden = {}
    for paf in sorted(glob.glob(os.path.join(HERE, "*.paf"))):
        name = os.path.basename(paf)
        idx = O.Index.from_paf(paf)
        cases = []
        for kw in PARAMS:
            out = idx.partition(O.make_partition_params(**kw))
            text = out["bed"]
            c = {"params": kw, "lines": text.count("\n"), "windows": len(out["windows"]), "n_partitions": out["n_partitions"],
                 "sha256": hashlib.sha256(text.encode()).hexdigest()}
            if len(text) <= 1500:
                c["text"] = text
            cases.append(c)
        names, lens, recs = F.parse_paf_like_reference([paf])
        impg = {str(int(b)): hashlib.sha256(F.encode(names, lens, F.entries_by_target(recs, b))).hexdigest() for b in (True, False)}
        golden[name] = {"partition": cases, "impg_sha256": impg}
        print(name, [c["lines"] for c in cases])
    with open(os.path.join(HERE, "golden_partition.json"), "w") as f:
        json.dump(golden, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
