#!/usr/bin/env python
"""How much do two concurrent batch calls on ONE index overlap on the device? (the liftover is bound by DRAM sector
fetches, the bucket merge by shared memory and barriers). Prints the time of one call, of two calls back to back and of
two calls from two host threads."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import impg_b200 as ix
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "c4p"
cfg, rows = bench.workload_cfg(ix, name)
recs, runs, offs, lens, names = ix.synth_generate(cfg)
idx = ix.Impg.from_records(recs, runs, offs, lens, names=names)
del runs
bed = ix.synth_bed(cfg, rows, seed=2)
p = bench.mode_params(ix, name)
halves = [bed[: len(bed) // 2], bed[len(bed) // 2:]]
for _ in range(2):
    idx.query_batch_bed(bed, p)
    for h in halves:
        idx.query_batch_bed(h, p)


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def concurrent():
    ts = [threading.Thread(target=lambda h=h: idx.query_batch_bed(h, p)) for h in halves]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


print(f"{name}: whole batch {timed(lambda: idx.query_batch_bed(bed, p)):.1f} ms; halves back to back "
      f"{timed(lambda: [idx.query_batch_bed(h, p) for h in halves]):.1f} ms; halves from two threads {timed(concurrent):.1f} ms")
