#!/usr/bin/env python
"""refine_bench.py — `impg refine` on a synthetic pangenome: wall time of impgx_refine (the
flank search of every locus batched by phase, <= 4 device batches per call) next to the CPU restatement of
run_refine (loci in parallel like the reference's par_iter, CIGARs pre-decoded in RAM — what the reference's
populate_cigar_cache leaves in its cache) on the contig-0 sub-world of the same index, records compared.

    python tests/refine_bench.py [--workload c3] [--loci 2000] [--span 1000] [--max-extension 0.5]

Not a bench.py line: refine's unit of work (a locus) is not BASELINE.json's metric. It lives under tests/
because its CPU leg runs the oracle (test infrastructure) as checker and baseline."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--loci", type=int, default=2000)
    ap.add_argument("--span", type=int, default=1000)
    ap.add_argument("--max-extension", type=float, default=0.5)
    ap.add_argument("--step", type=int, default=1000)
    ap.add_argument("--support-level", type=int, default=0)
    ap.add_argument("--transitive", type=int, default=0)
    ap.add_argument("--cpu-loci", type=int, default=0, help="loci of the CPU leg (default: all)")
    args = ap.parse_args()
    g, c, L, a, eq, rev, seed, _ = B.WORKLOADS[args.workload]
    cfg = ix.synth_cfg(g, c, L, a, eq, rev, seed)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    idx = ix.Impg.from_records(recs, runs, offs, lens, names=names)
    # loci on contig 0 of every genome: the sub-world the CPU leg holds
    rng = np.random.default_rng(5)
    loci = np.zeros(args.loci, ix.RANGE_DTYPE)
    loci["target_id"] = rng.integers(0, g, args.loci) * c
    ln = rng.integers(2000, 20000, args.loci)
    st = rng.integers(0, L - 20000, args.loci)
    loci["start"], loci["end"] = st, st + ln
    rp = ix.make_refine_params(span_bp=args.span, max_extension=args.max_extension, extension_step=args.step,
                               support_level=args.support_level, transitive=args.transitive)
    idx.refine(loci[:16], rp)  # warm-up
    t0 = time.perf_counter()
    got = idx.refine(loci, rp)
    gpu_s = time.perf_counter() - t0
    records = got[0] if isinstance(got, tuple) else got
    out = {"workload": args.workload, "alignments": int(len(recs)), "loci": args.loci, "span_bp": args.span,
           "max_extension": args.max_extension, "extension_step": args.step, "transitive": args.transitive,
           "gpu": {"wall_s": round(gpu_s, 4), "loci_per_s": round(args.loci / gpu_s, 1),
                   "ms_per_locus": round(1e3 * gpu_s / args.loci, 4)}}
    if isinstance(got, tuple) and len(got) > 1:
        out["gpu"]["candidates_batches"] = [int(x) for x in got[1]]
    import _oracle as O
    s_recs, s_runs, s_offs, s_lens, s_names = ix.synth_generate_contig(cfg, 0)
    orc = O.Index.build(s_recs, s_runs, s_offs, s_lens, names=s_names)
    n_cpu = args.cpu_loci or args.loci
    threads = B.host_threads()
    secs, want = orc.refine_parallel(loci[:n_cpu], rp, threads)
    keys = ("refined_start", "refined_end", "original_start", "original_end", "applied_left_extension",
            "applied_right_extension", "support_count", "original_support_count")
    have = np.array([[r[k] for k in keys] for r in records[:n_cpu]], np.int64)
    out["cpu"] = {"wall_s": round(secs, 4), "threads": threads, "loci": n_cpu, "loci_per_s": round(n_cpu / secs, 1),
                  "ms_per_locus": round(1e3 * secs / n_cpu, 4), "records_equal": bool((have == want).all()),
                  "kind": "oracle port of run_refine, loci in parallel, CIGARs in RAM, contig-0 sub-world"}
    out["speedup"] = round((args.loci / gpu_s) / (n_cpu / secs), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
