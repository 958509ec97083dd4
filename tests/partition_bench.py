#!/usr/bin/env python
"""partition_bench.py — `impg partition -o bed` on a synthetic pangenome:
wall time of impgx_partition (one masked BFS + BED merge on the device per window, host
bookkeeping between windows), next to the CPU restatement of partition_alignments with the
reference's cost structure (CIGAR text on disk, pread + parse per hit) on the contig-0
sub-world of the same index (contigs never align to one another in the synthetic world, so the
windows, masks and partitions of contig-0 sequences are the same in both).

    python tests/partition_bench.py [--workload c3] [--window 1000000] [-d 10000] [--keep 1.0] [--cpu-budget 30]

Not a bench.py line: partition's unit of work (a window) is not BASELINE.json's metric. It lives under
tests/ because its CPU leg runs the oracle (test infrastructure) as checker and baseline."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench as B  # noqa: E402
import impg_b200 as ix  # noqa: E402


def subset(recs, runs, offs, keep, seed):
    if keep >= 1.0:
        return recs, runs, offs
    rng = np.random.default_rng(seed)
    sel = np.flatnonzero(rng.random(len(recs)) < keep)
    n = np.diff(offs.astype(np.int64))[sel]
    offs2 = np.zeros(len(sel) + 1, np.uint64)
    np.cumsum(n, out=offs2[1:])
    idx = np.repeat(offs[sel].astype(np.int64) - offs2[:-1].astype(np.int64), n) + np.arange(int(offs2[-1]))
    return recs[sel].copy(), runs[idx], offs2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--window", type=int, default=1000000)
    ap.add_argument("-d", "--merge-distance", type=int, default=10000)
    ap.add_argument("--keep", type=float, default=1.0, help="fraction of the alignments kept (patchy homology)")
    ap.add_argument("--cpu-budget", type=float, default=30.0, help="skip the CPU leg with 0")
    ap.add_argument("--max-depth", type=int, default=2)
    args = ap.parse_args()
    g, c, L, a, eq, rev, seed, _ = B.WORKLOADS[args.workload]
    cfg = ix.synth_cfg(g, c, L, a, eq, rev, seed)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    recs, runs, offs = subset(recs, runs, offs, args.keep, seed)
    t0 = time.perf_counter()
    idx = ix.Impg.from_records(recs, runs, offs, lens, names=names)
    build_s = time.perf_counter() - t0
    kw = dict(window_size=args.window, merge_distance=args.merge_distance, max_depth=args.max_depth)
    idx.query_batch_bed(np.array([(0, 0, 5000)], dtype=ix.RANGE_DTYPE), ix.make_params(mode=ix.MODE_BFS))  # warm-up
    idx.partition(ix.make_partition_params(**kw))  # warm-up: scratch arena growth, pinned staging
    t0 = time.perf_counter()
    parts = idx.partition(ix.make_partition_params(**kw))
    gpu_s = time.perf_counter() - t0
    assert parts.partitioned_bp == parts.total_bp == int(lens.sum())
    out = {"workload": args.workload, "alignments": int(len(recs)), "sequences": int(len(lens)), "keep": args.keep,
           "window": args.window, "merge_distance": args.merge_distance, "max_depth": args.max_depth,
           "index_build_s": round(build_s, 3),
           "gpu": {"wall_s": round(gpu_s, 4), "windows": int(parts.n_windows), "partitions": int(parts.n_partitions),
                   "intervals": int(len(parts.start)), "ms_per_window": round(1e3 * gpu_s / max(parts.n_windows, 1), 4),
                   "bp_per_s": round(parts.total_bp / gpu_s, 1)}}
    st = idx.stats()
    out["gpu"]["launches_last_window"] = int(st["kernel_launches"])
    if args.cpu_budget > 0:
        import tempfile

        import _oracle as O
        # the sequences of contig 0 form a closed sub-world (contigs never align to one another): same windows,
        # masks and partitions there as in the full index; renumber them densely (id = genome * contigs + contig)
        keep = np.flatnonzero(recs["target_id"] % cfg.contigs == 0)
        n = np.diff(offs.astype(np.int64))[keep]
        s_offs = np.zeros(len(keep) + 1, np.uint64)
        np.cumsum(n, out=s_offs[1:])
        gi = np.repeat(offs[keep].astype(np.int64) - s_offs[:-1].astype(np.int64), n) + np.arange(int(s_offs[-1]))
        s_recs, s_runs = recs[keep].copy(), runs[gi]
        s_recs["query_id"] //= cfg.contigs
        s_recs["target_id"] //= cfg.contigs
        s_lens = lens[::cfg.contigs].copy()
        s_names = names[::cfg.contigs]
        # reference cost structure: CIGAR text on disk, pread + parse per hit (src/impg.rs:495-552)
        path = os.path.join(os.environ.get("IMPGX_TMP", tempfile.gettempdir()), f"impgx_part_{os.getpid()}.txt")
        o_off, o_len = ix.write_cigar_text(s_runs, s_offs, path)
        orc = O.Index.build(s_recs, np.zeros(1, np.uint32), np.zeros(len(s_recs) + 1, np.uint64), s_lens, names=s_names)
        orc.attach_cigar_file(path, o_off, o_len)
        threads = B.host_threads()
        t0 = time.perf_counter()
        want = orc.partition(O.make_partition_params(**kw), threads=threads)
        cpu_s = time.perf_counter() - t0
        os.unlink(path)
        sub = ix.Impg.from_records(s_recs, s_runs, s_offs, s_lens, names=s_names)
        sub.partition(ix.make_partition_params(**kw))  # warm-up: arena growth
        t0 = time.perf_counter()
        got = sub.partition(ix.make_partition_params(**kw))
        sub_gpu_s = time.perf_counter() - t0
        norm = [(p, s, min(x, y), max(x, y)) for p, s, x, y in want["rows"]]
        nw = max(len(want["windows"]), 1)
        out["contig0_subworld"] = {"cpu_wall_s": round(cpu_s, 3), "cpu_threads": threads, "gpu_wall_s": round(sub_gpu_s, 4),
                                   "windows": len(want["windows"]), "bit_exact": got.rows() == norm,
                                   "cpu_ms_per_window": round(1e3 * cpu_s / nw, 4),
                                   "gpu_ms_per_window": round(1e3 * sub_gpu_s / nw, 4),
                                   "launches_last_window": int(sub.stats()["kernel_launches"]),
                                   "cpu_kind": "oracle port, reference cost structure (pread + CIGAR parse per hit)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
