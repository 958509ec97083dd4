"""Latency of single-row calls (the shape partition / refine drive the index with through the ImpgIndex
trait: one window per call, src/commands/partition.rs:359-391): `python tools/latency_single_row.py`."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, impg_b200 as ix
cfg = ix.synth_cfg(50, 8, 2500000, 51, 100, 100, 1)
recs, runs, offs, lens, names = ix.synth_generate(cfg)
gpu = ix.Impg.from_records(recs, runs, offs, lens)
bed = ix.synth_bed(cfg, 400, seed=2)
mask = ix.mask_csr({s: [(1000, 50000)] for s in range(len(lens))}, len(lens))
for label, p in (("bfs d2 raw", ix.make_params(mode=ix.MODE_BFS, max_depth=2)),
                 ("bfs d2 raw masked", ix.make_params(mode=ix.MODE_BFS, max_depth=2, masked_regions=mask)),
                 ("bfs d2 bed", ix.make_params(mode=ix.MODE_BFS, max_depth=2, merge_distance=1000)),
                 ("query d1 raw", ix.make_params(mode=ix.MODE_QUERY))):
    for k in range(20):
        (gpu.query_batch_bed if "bed" in label else gpu.query_batch)(bed[k:k + 1], p)
    t0 = time.perf_counter()
    n = 200
    for k in range(n):
        (gpu.query_batch_bed if "bed" in label else gpu.query_batch)(bed[k:k + 1], p)
    dt = (time.perf_counter() - t0) / n
    print(f"{label}: {dt*1e3:.2f} ms per single-row call, launches {gpu.stats()['kernel_launches']}")
