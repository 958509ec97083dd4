#!/usr/bin/env python
"""bench.py — ranges projected/sec for `impg query -b <BED> -x -m 2 -o bed` on
synthetic all-vs-all alignments (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3|c2|c4|tiny]

A step = one pass of the hot path over one batch of BED rows (stab, liftover,
BFS fold/frontier, BED merges). `value` times it with the rows already in HBM
and the merged rows left in HBM; `e2e` times the reference-facing C-ABI call
with HOST buffers (H2D of the rows, D2H of the merged rows inside the timed
region). `--impl reference` times the reference's CPU algorithm (the oracle
port with the reference's cost structure: per-hit pread + CIGAR text parse,
rows serial, threads inside a BFS level) on a bounded sample of the same rows.
Multi-GPU (weak scaling: N x the rows of one GPU), two layouts measured in the
same run on the same rows:
  * rows over index replicas — rows are independent, each rank holds a replica
    and owns its batch of rows, no data-path collective (the headline `value`
    unless --parallelism targets);
  * index sharded by target sequence — every rank owns 1/N of the
    sequences, the batch of all N x rows is one collective call, lifted hits and
    frontier ranges are exchanged over NCCL between hops (`target_sharded`).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# torchrun exports OMP_NUM_THREADS=1; the host-side legs (BED text of a batch, CPU reference) split the
# box's cores over the ranks instead (set before any OpenMP runtime is loaded)
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

import numpy as np  # noqa: E402

# stdout carries exactly one JSON line: NCCL's banner ("NCCL version ...", printed at NCCL_DEBUG=VERSION/WARN)
# and any other NCCL log go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "ranges projected/sec (batch -b query, -x depth 2)"
UNIT = "ranges/s"

WORKLOADS = {
    # name: (genomes, contigs, contig_len, tiles, eq_mean, rev_permille, seed, bed_rows)
    "tiny": (8, 2, 200000, 10, 100, 100, 1, 512),
    "c2": (50, 8, 2500000, 51, 100, 100, 1, 10000),   # BASELINE configs[1]: depth 1
    "c3": (50, 8, 2500000, 51, 100, 100, 1, 10000),   # BASELINE configs[2]: -x -m 2
    "c4": (200, 8, 2500000, 63, 200, 100, 1, 100000),  # BASELINE configs[3] on one GPU
    "c4p": (200, 8, 2500000, 63, 200, 100, 1, 5000),   # the c4 index with a short BED (profiling: two row batches)
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not
    throttle the CPU reference: it is given every core, like rayon would take)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def gen_workload(ix, name, rank):
    g, c, L, a, eq, rev, seed, rows = WORKLOADS[name]
    cfg = ix.synth_cfg(g, c, L, a, eq, rev, seed)
    t0 = time.time()
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    from impg_b200 import dist as D

    bed = ix.synth_bed(cfg, rows, seed=D.rank_seed(2, rank))
    return cfg, recs, runs, offs, lens, names, bed, time.time() - t0


def mode_params(ix_or_O, name, is_oracle=False):
    depth1 = name == "c2"
    if is_oracle:
        return ix_or_O.make_params(mode=ix_or_O.MODE_QUERY if depth1 else ix_or_O.MODE_BFS, max_depth=2,
                                   min_transitive_len=101, min_dist=10, merge_distance=1000, merge_strands=True)
    return ix_or_O.make_params(mode=ix_or_O.MODE_QUERY if depth1 else ix_or_O.MODE_BFS, max_depth=2,
                               min_transitive_len=101, min_distance_between_ranges=10, merge_distance=1000,
                               merge_strands=True)


def contig_subworld(cfg, recs, runs, offs, bed, contig=0):
    """The alignments and BED rows of ONE contig. The synthetic world only aligns contig c of
    one genome with contig c of another, so the trees, hits and transitive closure of a row on
    contig c are identical in this sub-world and in the full index: the CPU reference can be
    timed on it without holding the CIGAR text of all 20 M alignments (c4)."""
    C_ = cfg.contigs
    keep = np.nonzero(recs["target_id"] % C_ == contig)[0]
    nr = np.diff(offs.astype(np.int64))[keep]
    sub_offs = np.zeros(len(keep) + 1, np.uint64)
    np.cumsum(nr, out=sub_offs[1:])
    sub_runs = np.empty(int(sub_offs[-1]), np.uint32)
    A = cfg.tiles  # alignments of one (pair, contig) are consecutive: copy them block-wise
    starts = keep[::A]
    pos = 0
    for a in starts:
        lo, hi = int(offs[a]), int(offs[a + A])
        sub_runs[pos:pos + hi - lo] = runs[lo:hi]
        pos += hi - lo
    assert pos == len(sub_runs)
    rows = bed[bed["target_id"] % C_ == contig]
    return recs[keep], sub_runs, sub_offs, rows


def cpu_reference_setup(name, recs, runs, offs, lens, names, ix):
    """Oracle index with the reference's cost structure: CIGARs stay as TEXT in a
    file, every hit preads + parses its whole CIGAR (reference src/impg.rs:495-552)."""
    import _oracle as O

    tmpdir = os.environ.get("IMPGX_TMP", tempfile.gettempdir())
    path = os.path.join(tmpdir, f"impgx_cigars_{name}_{os.getpid()}.txt")
    t0 = time.time()
    o_off, o_len = ix.write_cigar_text(runs, offs, path)
    orc = O.Index.build(recs, np.zeros(1, np.uint32), np.zeros(len(recs) + 1, np.uint64), lens, names=names)
    orc.attach_cigar_file(path, o_off, o_len)
    return O, orc, path, time.time() - t0


def cpu_reference_time(O, orc, bed, name, budget_s, threads):
    """Times the reference driver on a bounded row sample sized for ~budget_s."""
    p = mode_params(O, name, is_oracle=True)
    n_probe = min(4, len(bed))
    orc.run_batch(bed[:1], p, threads=threads, fmt="bed")  # warm the page cache / allocator
    t, nres, _, _ = orc.run_batch(bed[:n_probe], p, threads=threads, fmt="bed")
    per_row = max(t / n_probe, 1e-6)
    n = int(max(n_probe, min(len(bed), budget_s / per_row)))
    t, nres, nbytes, csum = orc.run_batch(bed[:n], p, threads=threads, fmt="bed")
    return n / t, n, t, nres


def run_shard_only(args, ix, torch, dist, name, rank, local_rank, world, config):
    """N > 1, index sharded by target sequence, no replica anywhere: rank r generates and
    uploads only the alignments its sequences walk; the batch of world x rows BED rows is one
    collective call per step."""
    from impg_b200 import dist as D

    g, c, L, a, eq, rev, seed, rows = WORKLOADS[name]
    cfg = ix.synth_cfg(g, c, L, a, eq, rev, seed)
    t0 = time.time()
    recs, runs, offs, lens, names, owner = ix.synth_generate_shard(cfg, world, rank)
    gen_s = time.time() - t0
    t0 = time.time()
    shard = ix.Impg.from_records_shard(recs, runs, offs, lens, owner, rank, world, device=local_rank)
    build_s = time.time() - t0
    del runs
    comm = D.nccl_comm(rank, world, local_rank)
    p = mode_params(ix, name)
    gbed = np.concatenate([ix.synth_bed(cfg, rows, seed=D.rank_seed(2, r)) for r in range(world)])
    bed_bytes = torch.from_numpy(gbed.view(np.uint8).copy())
    d_bed = bed_bytes.cuda()
    h_bed_np = np.frombuffer(bed_bytes.pin_memory().numpy(), dtype=ix.RANGE_DTYPE)
    stream = torch.cuda.current_stream()
    n = len(gbed)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        shard.query_batch_bed_sharded_device(comm, d_bed.data_ptr(), n, p, stream.cuda_stream)
    shard.query_batch_bed_sharded(comm, h_bed_np, p)
    barrier()
    if rank == 0:
        sampler.start()
    tr0 = comm.traffic()
    acc = {k: 0 for k in ("kernel_launches", "lift_bytes", "liftovers", "lift_launches", "lift_touched_bytes",
                          "lift_window_runs", "lift_ms", "stab_ms", "fold_ms", "merge_ms", "exchange_ms")}
    merged = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        r = shard.query_batch_bed_sharded_device(comm, d_bed.data_ptr(), n, p, stream.cuda_stream)
        st = shard.stats()
        for k in acc:
            acc[k] += st[k]
        merged = st["merged"]
        del r
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    tr1 = comm.traffic()
    h2d = d2h = 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        r = shard.query_batch_bed_sharded(comm, h_bed_np, p)
        st = shard.stats()
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]
        del r
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    barrier()
    if rank == 0:
        sampler.stop()
    dev_ms, e2e_ms = D.max_over_ranks([dev_ms, e2e_ms], device="cuda")
    merged_total = D.gather_row_counts(merged, device="cuda")
    lift_total = D.gather_row_counts(acc["liftovers"], device="cuda")
    sent_total = D.gather_row_counts(tr1["bytes_sent"] - tr0["bytes_sent"], device="cuda")
    bytes_max = D.max_over_ranks([float(shard.device_bytes)], device="cuda")[0]
    if rank != 0:
        dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    lift_ms = acc["lift_ms"]
    achieved = (acc["lift_bytes"] / 1e9) / (lift_ms / 1e3) if lift_ms > 0 else 0.0
    config["parallelism"] = (f"index sharded by target sequence over {world} GPUs (no replica), one collective batch of "
                             f"{world} x {rows} rows per step, NCCL hit / frontier exchange between hops")
    line = {"metric": METRIC, "value": n * args.steps / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": n * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "note": "bytes of rank 0; every rank copies the rows in and its own share of the BED rows out"},
            "gpu_launches": int(acc["kernel_launches"]),
            "roofline": {"kernel": "k_liftover_ends", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "rank": 0,
                         "algorithmic_bytes_per_launch": acc["lift_bytes"] / max(1, acc["lift_launches"]),
                         "r_ov_mean": acc["lift_window_runs"] / max(1, acc["liftovers"]),
                         "avg_launch_ms": lift_ms / max(1, acc["lift_launches"]),
                         "liftovers_per_step_all_ranks": lift_total // args.steps,
                         "step_share": {"liftover_ms": lift_ms / args.steps, "stab_ms": acc["stab_ms"] / args.steps,
                                        "fold_ms": acc["fold_ms"] / args.steps, "merge_ms": acc["merge_ms"] / args.steps,
                                        "exchange_ms_within_merge_and_fold": acc["exchange_ms"] / args.steps,
                                        "step_ms": dev_ms / args.steps}},
            "cpu_baseline": None, "clocks": sampler.summary(), "bed_rows_out_per_step": int(merged_total),
            "target_sharded": {"n_shards": world, "transport": "nccl", "rows_per_step": n,
                               "exchange_bytes_per_step": int(sent_total) // args.steps,
                               "exchanges_per_step": (tr1["exchanges"] - tr0["exchanges"]) // args.steps,
                               "shard_device_bytes_max": int(bytes_max)},
            "setup": {"generate_s": gen_s, "index_build_s": build_s, "index_device_bytes": shard.device_bytes}}
    emit(line)
    dist.destroy_process_group()
    return 0


_JSON_FD = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line. Libraries print there too (NCCL's version banner,
    for one), so the process's fd 1 is pointed at stderr and the JSON line goes to a private
    duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="impgx", choices=["impgx", "reference"])
    ap.add_argument("--workload", default=os.environ.get("IMPGX_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parallelism", default=os.environ.get("IMPGX_BENCH_PARALLELISM", "rows"),
                    choices=["rows", "targets"], help="which multi-GPU layout is the headline value at N > 1")
    ap.add_argument("--virtual-shards", type=int, default=int(os.environ.get("IMPGX_BENCH_VIRTUAL_SHARDS", "0")),
                    help="N = 1 only: also time the target-sharded path with this many virtual ranks on the one GPU")
    ap.add_argument("--shard-only", action="store_true",
                    help="N > 1: skip the replica leg; every rank generates and holds only its shard (for indexes "
                         "whose full copy per rank would not fit the host, e.g. c4 on 8 GPUs)")
    ap.add_argument("--profile", action="store_true",
                    help="for ncu: exactly --warmup + --steps device-resident steps, no e2e / cpu legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    g, c, L, a, eq, rev, seed, rows = WORKLOADS[name]
    config = {"workload": f"{name}: synthetic {g}-genome all-vs-all PAF, {g * (g - 1) * c * a} alignments, "
                          f"{rows}-row BED per GPU, " + ("depth 1" if name == "c2" else "-x -m 2") + ", -d 1000 -o bed",
              "genomes": g, "contigs": c, "contig_len": L, "alignments": g * (g - 1) * c * a, "bed_rows_per_gpu": rows,
              "parallelism": f"rows sharded over {world} index replica(s)", "cache": "inputs larger than L2 "
              "(index run stream >> 126 MB; every step re-reads it from HBM)"}
    import impg_b200 as ix

    # ------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg, recs, runs, offs, lens, names, bed, gen_s = gen_workload(ix, name, 0)
        if name.startswith("c4"):
            recs, runs, offs, bed = contig_subworld(cfg, recs, runs, offs, bed)
        O, orc, path, setup_s = cpu_reference_setup(name, recs, runs, offs, lens, names, ix)
        threads = host_threads()
        try:
            per_step_budget = max(2.0, min(args.cpu_budget, 150.0 / max(1, args.steps + args.warmup)))
            vals, n_used = [], 0
            for i in range(args.warmup + args.steps):
                v, n_used, t, nres = cpu_reference_time(O, orc, bed, name, per_step_budget, threads)
                if i >= args.warmup:
                    vals.append((v, t))
            value = float(np.mean([v for v, _ in vals])) if vals else 0.0
            ms = float(np.mean([t for _, t in vals]) * 1e3) if vals else 0.0
        finally:
            try:
                os.unlink(path)
            except OSError:
                pass
        sample = f"first {n_used} of {rows} BED rows per step (reference driver: rows serial, {threads} threads inside a BFS level, per-hit pread+parse of CIGAR text)"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------ GPU arm
    import torch

    if not torch.cuda.is_available() or ix.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libimpgx has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.shard_only and world > 1:
        return run_shard_only(args, ix, torch, dist, name, rank, local_rank, world, config)

    cfg, recs, runs, offs, lens, names, bed, gen_s = gen_workload(ix, name, rank)
    t0 = time.time()
    idx = ix.Impg.from_records(recs, runs, offs, lens, names=names, device=local_rank)
    build_s = time.time() - t0
    p = mode_params(ix, name)
    n = len(bed)

    # rows resident in HBM (value) and in pinned host memory (e2e)
    bed_bytes = torch.from_numpy(bed.view(np.uint8).copy())
    d_bed = bed_bytes.cuda()
    h_bed = bed_bytes.pin_memory()
    h_bed_np = np.frombuffer(h_bed.numpy(), dtype=ix.RANGE_DTYPE)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        r = idx.query_batch_bed_device(d_bed.data_ptr(), n, p, stream.cuda_stream)
        st = idx.stats()
        return r, st

    def step_host():
        r = idx.query_batch_bed(h_bed_np, p)
        st = idx.stats()
        return r, st

    if args.profile:
        for _ in range(args.warmup + args.steps):
            step_device()
        torch.cuda.synchronize()
        emit({"profile_run": True, "workload": name, "steps": args.steps, "warmup": args.warmup,
              "stats_last_step": idx.stats()})
        return 0

    sampler = ClockSampler(local_rank)
    for _ in range(max(args.warmup, 3)):
        step_device()
    for _ in range(1):
        step_host()

    # ---- value: device-resident timing (CUDA events on the launching stream)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = lift_bytes = liftovers = lift_launches = merged = touched = window_runs = 0
    lift_ms = stab_ms = fold_ms = merge_ms = 0.0
    e0.record(stream)
    for _ in range(args.steps):
        r, st = step_device()
        launches += st["kernel_launches"]
        lift_bytes += st["lift_bytes"]
        liftovers += st["liftovers"]
        lift_launches += st["lift_launches"]
        touched += st["lift_touched_bytes"]
        window_runs += st["lift_window_runs"]
        lift_ms += st["lift_ms"]; stab_ms += st["stab_ms"]; fold_ms += st["fold_ms"]; merge_ms += st["merge_ms"]
        merged = st["merged"]
        del r
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region
    barrier()
    h2d = d2h = 0
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    w0 = time.perf_counter()
    for _ in range(args.steps):
        r, st = step_host()
        h2d += st["h2d_bytes"]; d2h += st["d2h_bytes"]
        del r
    f1.record(stream)
    torch.cuda.synchronize()
    host_wall_ms = (time.perf_counter() - w0) * 1e3
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), host_wall_ms)
    if rank == 0:
        sampler.stop()

    # ---- e2e with the BED text (SURVEY.md 8d "text formatting included"): the same call plus
    # impgx_format_bed_batch over every row — what `impgx-query -b ... -o bed` writes to stdout
    names_arr = (C.c_char_p * n)(*[f"r{k}".encode() for k in range(n)])
    text_bytes = 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        r, st = step_host()
        text_bytes += idx.format_bed_batch(r, names_arr, length_only=True)
        del r
    torch.cuda.synchronize()
    text_ms = (time.perf_counter() - w0) * 1e3
    barrier()

    from impg_b200 import dist as D

    dev_ms, e2e_ms, text_ms = D.max_over_ranks([dev_ms, e2e_ms, text_ms], device="cuda")  # slowest rank defines the step
    total_rows = D.gather_row_counts(n, device="cuda") * args.steps
    value = total_rows / (dev_ms / 1e3)
    e2e_value = total_rows / (e2e_ms / 1e3)

    # ---- index sharded by target sequence: the same N x rows as ONE collective batch
    sharded = None
    n_shards = world if world > 1 else args.virtual_shards
    if n_shards > 1:
        merged_rows_mode = D.gather_row_counts(merged, device="cuda")
        gbed = np.concatenate([ix.synth_bed(cfg, rows, seed=D.rank_seed(2, r)) for r in range(world)])
        d_gbed = torch.from_numpy(gbed.view(np.uint8).copy()).cuda()
        owner = ix.assign_owners(recs, offs, len(lens), n_shards)
        t0 = time.time()
        if world > 1:
            comm = D.nccl_comm(rank, world, local_rank)
            shard = ix.Impg.from_records_shard(recs, runs, offs, lens, owner, rank, world, device=local_rank)
            sh_stats = lambda: [shard.stats()]
            sh_bytes = shard.device_bytes
            traffic = lambda: comm.traffic()

            def step_sharded():
                return [shard.query_batch_bed_sharded_device(comm, d_gbed.data_ptr(), len(gbed), p, stream.cuda_stream)]
        else:
            sh = ix.ShardedImpg.from_records(recs, runs, offs, lens, [local_rank] * n_shards)
            sh_stats = sh.stats
            sh_bytes = max(x.device_bytes for x in sh.shards)

            def traffic():
                t = [c.traffic() for c in sh.comms]
                return {k: sum(x[k] for x in t) for k in t[0]}

            def step_sharded():
                return sh._collective(lambda r: sh.shards[r].query_batch_bed_sharded_device(
                    sh.comms[r], d_gbed.data_ptr(), len(gbed), p, 0))
        shard_build_s = time.time() - t0
        for _ in range(max(args.warmup, 3)):
            step_sharded()
        barrier()
        tr0 = traffic()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sh_merged = sh_lift = sh_launches = 0
        sh_lift_ms = sh_merge_ms = sh_fold_ms = sh_stab_ms = sh_exch_ms = 0.0
        w0 = time.perf_counter()
        g0.record(stream)
        for _ in range(args.steps):
            rs = step_sharded()
            sts = sh_stats()
            sh_merged = sum(x["merged"] for x in sts)
            sh_lift += sum(x["liftovers"] for x in sts)
            sh_launches += sum(x["kernel_launches"] for x in sts)
            sh_lift_ms += max(x["lift_ms"] for x in sts); sh_merge_ms += max(x["merge_ms"] for x in sts)
            sh_fold_ms += max(x["fold_ms"] for x in sts); sh_stab_ms += max(x["stab_ms"] for x in sts)
            sh_exch_ms += max(x["exchange_ms"] for x in sts)
            del rs
        g1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - w0) * 1e3
        barrier()
        tr1 = traffic()
        # virtual ranks run on their own streams: the host wall clock (all threads joined) is the time
        sh_ms = g0.elapsed_time(g1) if world > 1 else wall_ms
        (sh_ms,) = D.max_over_ranks([sh_ms], device="cuda")
        sh_merged_total = D.gather_row_counts(sh_merged, device="cuda")
        sent = D.gather_row_counts(tr1["bytes_sent"] - tr0["bytes_sent"], device="cuda") if world > 1 \
            else tr1["bytes_sent"] - tr0["bytes_sent"]
        sharded = {"value": len(gbed) * args.steps / (sh_ms / 1e3), "unit": UNIT, "ms_per_step": sh_ms / args.steps,
                   "n_shards": n_shards, "transport": "nccl" if world > 1 else "in-process (virtual ranks on one GPU)",
                   "rows_per_step": len(gbed), "exchange_bytes_per_step": int(sent) // args.steps,
                   "exchanges_per_step": (tr1["exchanges"] - tr0["exchanges"]) // args.steps,
                   "bed_rows_out_per_step": int(sh_merged_total),
                   "bed_rows_out_match_rows_mode": bool(sh_merged_total == merged_rows_mode),
                   "liftovers_per_step_this_rank": sh_lift // args.steps, "shard_device_bytes": int(sh_bytes),
                   "shard_build_s": shard_build_s, "gpu_launches": int(sh_launches),
                   "step_share_slowest_rank": {"liftover_ms": sh_lift_ms / args.steps, "stab_ms": sh_stab_ms / args.steps,
                                               "fold_ms": sh_fold_ms / args.steps, "merge_ms": sh_merge_ms / args.steps,
                                               "exchange_ms_within_merge_and_fold": sh_exch_ms / args.steps}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    achieved = (lift_bytes / 1e9) / (lift_ms / 1e3) if lift_ms > 0 else 0.0
    roofline = {"kernel": "k_liftover_ends", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": lift_bytes / max(1, lift_launches),
                "algorithmic_bytes_formula": "SURVEY.md 8(d): per liftover 32 (entry) + 16 (2 checkpoints) + 4*r_ov "
                                             "(runs intersecting the request) + 24 (hit out)",
                "r_ov_mean": window_runs / max(1, liftovers),
                "touched_bytes_per_launch": touched / max(1, lift_launches),
                "touched_GBps": (touched / 1e9) / (lift_ms / 1e3) if lift_ms > 0 else 0.0,
                "avg_launch_ms": lift_ms / max(1, lift_launches), "liftovers_per_step": liftovers / args.steps,
                "step_share": {"liftover_ms": lift_ms / args.steps, "stab_ms": stab_ms / args.steps,
                               "fold_ms": fold_ms / args.steps, "merge_ms": merge_ms / args.steps,
                               "step_ms": dev_ms / args.steps}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("k_liftover_dram_bytes_per_launch")
        except Exception:
            pass

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        c_recs, c_runs, c_offs, c_bed, note = recs, runs, offs, bed, ""
        if name.startswith("c4"):
            c_recs, c_runs, c_offs, c_bed = contig_subworld(cfg, recs, runs, offs, bed)
            note = " on contig 0 (per-row work identical to the full index: alignments never cross contigs)"
        O, orc, path, setup_s = cpu_reference_setup(name, c_recs, c_runs, c_offs, lens, names, ix)
        try:
            threads = host_threads()
            v, n_used, tsec, nres = cpu_reference_time(O, orc, c_bed, name, args.cpu_budget, threads)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"first {n_used} of {rows} BED rows{note}, {tsec:.1f} s (oracle port with the reference's "
                                      "cost structure: rows serial, threads inside a BFS level, per-hit pread+parse)"}
            # parity at full size on a row sample: the oracle's merged BED rows vs the C-ABI's, bit for bit
            # (outside every timed region; the oracle is the checker, never the thing measured)
            k = min(16, len(c_bed))
            want, woffs = orc.query_batch(c_bed[:k], mode_params(O, name, is_oracle=True), bed_merge=True)
            wc, gc = want.columns(), idx.query_batch_bed(c_bed[:k], p).columns()
            exact = gc["row_offsets"].tolist() == woffs.tolist() and all(
                (gc[c] == wc[c]).all() for c in ("q_id", "q_first", "q_last"))
            cpu_baseline["parity_sample"] = {"rows": int(k), "bed_rows": int(len(wc["q_id"])), "bit_exact": bool(exact)}
            if not exact:
                raise SystemExit("bench.py: the CUDA path and the oracle disagree on the parity sample")
        finally:
            try:
                os.unlink(path)
            except OSError:
                pass

    rows_mode = {"value": value, "unit": UNIT, "ms_per_step": dev_ms / args.steps}
    if args.parallelism == "targets" and sharded is not None and world > 1:
        value, dev_ms = sharded["value"], sharded["ms_per_step"] * args.steps
        launches = sharded["gpu_launches"]
        config["parallelism"] = (f"index sharded by target sequence over {world} GPUs, one collective batch of "
                                 f"{world} x {rows} rows, NCCL hit / frontier exchange between hops")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps},
            "e2e_text": {"value": total_rows / (text_ms / 1e3), "unit": UNIT, "ms_per_step": text_ms / args.steps,
                         "text_bytes_per_step": int(text_bytes // args.steps),
                         "what": "e2e plus the BED text of every row (impgx_format_bed_batch on the host cores)"},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(), "bed_rows_out_per_step": int(merged), "rows_over_replicas": rows_mode,
            "target_sharded": sharded,
            "setup": {"generate_s": gen_s, "index_build_s": build_s, "index_device_bytes": idx.device_bytes}}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
