#!/usr/bin/env python
"""bench.py — ranges projected/sec for `impg query -b <BED> -x -m 2 -d 1000 -o bed` on
synthetic all-vs-all alignments (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4|c3|c2|tiny]

Default workload: c4 = BASELINE configs[3], the config the metric is quoted on (200 genomes,
20.06 M alignments, 100,000-row BED, -x -m 2); it fits one B200 (41 GB index).

A step = one pass of the hot path over one batch of BED rows (stab, liftover, BFS fold / frontier,
BED merges). `value` times it with the rows already in HBM and the merged rows left in HBM; `e2e`
times the reference-facing C-ABI call with HOST buffers (H2D of the rows, D2H of the merged rows
inside the timed region). `--impl reference` times the reference's CPU algorithm (the oracle port
with the reference's cost structure: per-hit pread + CIGAR text parse, rows serial, threads inside
a BFS level) on a bounded sample of the same rows.

N = 1: one GPU holds the whole index.
N > 1 (weak scaling: N x the rows of one GPU per step), headline layout = what north_star names:
  * index sharded by target sequence — every rank owns 1/N of the sequences and
    generates / holds only its shard, the batch of all N x rows is ONE collective call per step,
    lifted hits and frontier ranges are exchanged over NCCL between hops. A row sample of the
    sharded result is gathered to rank 0 and compared bit for bit with the oracle
    (`target_sharded.parity_sample`).
  * `--parallelism rows` (or `--with-replicas`, default for indexes below 8 GB) additionally times
    rows over index replicas (no data-path collective) and reports it as `rows_over_replicas`.
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
PARITY_ROWS = 64

WORKLOADS = {
    # name: (genomes, contigs, contig_len, tiles, eq_mean, rev_permille, seed, bed_rows)
    "tiny": (8, 2, 200000, 10, 100, 100, 1, 512),
    "c2": (50, 8, 2500000, 51, 100, 100, 1, 10000),   # BASELINE configs[1]: depth 1
    "c3": (50, 8, 2500000, 51, 100, 100, 1, 10000),   # BASELINE configs[2]: -x -m 2
    "c4": (200, 8, 2500000, 63, 200, 100, 1, 100000),  # BASELINE configs[3]: the config the metric is quoted on
    "c4p": (200, 8, 2500000, 63, 200, 100, 1, 5000),   # the c4 index with a short BED (profiling: two row batches)
    "c4d1": (200, 8, 2500000, 63, 200, 100, 1, 100000),  # BASELINE configs[3] at depth 1 ("run both")
    # BASELINE configs[4], sparsified as SURVEY.md 8d allows: 500 genomes, every genome aligned against k = 50 others
    # (12.6 M alignments, ~100 neighbours per genome in the bidirectional index), 125,000 rows per GPU = 1 M rows on
    # 8 GPUs, -x -m 3
    "c5": (500, 8, 2500000, 63, 400, 100, 1, 125000),
    "c5p": (500, 8, 2500000, 63, 400, 100, 1, 4000),   # the c5 index with a short BED (one-GPU check of the depth-3 path)
}
PARTNERS = {"c5": 50, "c5p": 50}
DEPTH = {"c2": 1, "c4d1": 1, "c5": 3, "c5p": 3}
# workloads whose CPU reference / parity oracle runs on the alignments of contig 0 only (per-row work is
# identical to the full index: the synthetic world never aligns across contigs)
SUBWORLD = {"c4", "c4p", "c4d1", "c5", "c5p"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """DRAM bytes per lifted hit / per merged box from the committed ncu --set full captures of THIS workload
    (profiles/traffic.json, keyed by workload), or None: the roofline's `traffic` is never borrowed
    from another workload's capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get("c4" if name in ("c4", "c4p") else name)
    except Exception:
        return None


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


def workload_cfg(ix, name):
    g, c, L, a, eq, rev, seed, rows = WORKLOADS[name]
    return ix.synth_cfg(g, c, L, a, eq, rev, seed, partners=PARTNERS.get(name, 0)), rows


def mode_params(ix_or_O, name, is_oracle=False):
    depth = DEPTH.get(name, 2)
    depth1 = depth == 1
    if is_oracle:
        return ix_or_O.make_params(mode=ix_or_O.MODE_QUERY if depth1 else ix_or_O.MODE_BFS, max_depth=depth,
                                   min_transitive_len=101, min_dist=10, merge_distance=1000, merge_strands=True)
    return ix_or_O.make_params(mode=ix_or_O.MODE_QUERY if depth1 else ix_or_O.MODE_BFS, max_depth=depth,
                               min_transitive_len=101, min_distance_between_ranges=10, merge_distance=1000,
                               merge_strands=True)


class CpuWorld:
    """What the CPU legs (reference arm, cpu_baseline, parity samples) run on: the whole synthetic
    world, or for the c4 index the alignments of contig 0 — the synthetic world only aligns contig c
    of one genome with contig c of another, so the trees, hits and transitive closure of a row on
    contig 0 are identical in this sub-world and in the full index, and the CPU side need not hold
    the CIGAR text of all 20 M alignments. Built without a GPU (oracle + generator only)."""

    def __init__(self, ix, name, full=None):
        import _oracle as O

        self.O, self.ix, self.name = O, ix, name
        cfg, _ = workload_cfg(ix, name)
        self.cfg = cfg
        self.sub = name in SUBWORLD
        if self.sub:
            self.recs, self.runs, self.offs, self.lens, self.names = ix.synth_generate_contig(cfg, 0)
            self.note = " on contig 0 (per-row work identical to the full index: alignments never cross contigs)"
        else:
            self.recs, self.runs, self.offs, self.lens, self.names = full if full is not None else ix.synth_generate(cfg)[:5]
            self.note = ""
        self._faithful = self._ram = None
        self.path = None

    def rows_of(self, bed):
        return bed[bed["target_id"] % self.cfg.contigs == 0] if self.sub else bed

    def faithful(self):
        """Oracle index with the reference's cost structure: CIGARs stay as TEXT in a file, every hit
        preads + parses its whole CIGAR (reference src/impg.rs:495-552)."""
        if self._faithful is None:
            tmpdir = os.environ.get("IMPGX_TMP", tempfile.gettempdir())
            self.path = os.path.join(tmpdir, f"impgx_cigars_{self.name}_{os.getpid()}.txt")
            o_off, o_len = self.ix.write_cigar_text(self.runs, self.offs, self.path)
            orc = self.O.Index.build(self.recs, np.zeros(1, np.uint32), np.zeros(len(self.recs) + 1, np.uint64), self.lens,
                                     names=self.names)
            orc.attach_cigar_file(self.path, o_off, o_len)
            self._faithful = orc
        return self._faithful

    def in_ram(self):
        """Oracle index with the CIGARs pre-decoded in RAM (parity checks, the "CPU-batched" baseline)."""
        if self._ram is None:
            self._ram = self.O.Index.build(self.recs, self.runs, self.offs, self.lens, names=self.names)
        return self._ram

    def close(self):
        if self.path:
            try:
                os.unlink(self.path)
            except OSError:
                pass
            self.path = None

    def time_reference(self, bed, budget_s, threads):
        """The reference driver (rows serial, threads inside a BFS level) on a row sample sized for ~budget_s."""
        orc, p = self.faithful(), mode_params(self.O, self.name, is_oracle=True)
        n_probe = min(4, len(bed))
        orc.run_batch(bed[:1], p, threads=threads, fmt="bed")  # warm the page cache / allocator
        t, _, _, _ = orc.run_batch(bed[:n_probe], p, threads=threads, fmt="bed")
        per_row = max(t / n_probe, 1e-6)
        n = int(max(n_probe, min(len(bed), budget_s / per_row)))
        t, nres, _, _ = orc.run_batch(bed[:n], p, threads=threads, fmt="bed")
        return n / t, n, t, nres

    def time_batched(self, bed, budget_s, threads):
        """The "CPU-batched" second baseline: rows in parallel, CIGARs pre-decoded in RAM."""
        orc, p = self.in_ram(), mode_params(self.O, self.name, is_oracle=True)
        n_probe = min(max(threads, 4), len(bed))
        t, _, _ = orc.run_batch_rows_parallel(bed[:n_probe], p, threads=threads, fmt="bed")
        per_row = max(t / n_probe, 1e-6)
        n = int(max(n_probe, min(len(bed), budget_s / per_row)))
        t, nres, _ = orc.run_batch_rows_parallel(bed[:n], p, threads=threads, fmt="bed")
        return n / t, n, t

    def parity(self, sample_rows, got_cols):
        """The oracle's merged BED rows of `sample_rows` vs the CUDA path's, bit for bit."""
        want, woffs = self.in_ram().query_batch(sample_rows, mode_params(self.O, self.name, is_oracle=True), bed_merge=True)
        wc = want.columns()
        exact = got_cols["row_offsets"].tolist() == woffs.tolist() and all(
            (np.asarray(got_cols[c]) == wc[c]).all() for c in ("q_id", "q_first", "q_last"))
        return {"rows": int(len(sample_rows)), "bed_rows": int(len(wc["q_id"])), "bit_exact": bool(exact),
                "rows_spread": "evenly over the BED rows of the oracle's world" + self.note}


def spread_sample(rows, k):
    if len(rows) <= k:
        return rows
    idx = np.unique(np.linspace(0, len(rows) - 1, k).astype(np.int64))
    return rows[idx]


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


STAT_SUMS = ("kernel_launches", "lift_bytes", "liftovers", "lift_launches", "lift_touched_bytes", "lift_window_runs",
             "lift_ms", "stab_ms", "fold_ms", "merge_ms", "exchange_ms", "merge_boxes", "merge_kernel_ms", "exchange_bytes")


def acc_stats(acc, st):
    for k in STAT_SUMS:
        acc[k] = acc.get(k, 0) + st.get(k, 0)


def rooflines(name, acc, steps, step_ms, peak, peak_src, extra_share=None):
    """`roofline` (the liftover kernel, north_star's target) and `roofline_merge` (the BED segment merge)."""
    lift_ms = acc["lift_ms"]
    n_l = max(1, acc["lift_launches"])
    achieved = (acc["lift_bytes"] / 1e9) / (lift_ms / 1e3) if lift_ms > 0 else 0.0
    tr = ncu_traffic(name) or {}
    hits_per_launch = acc["liftovers"] / n_l
    share = {"liftover_ms": lift_ms / steps, "stab_ms": acc["stab_ms"] / steps, "fold_ms": acc["fold_ms"] / steps,
             "merge_ms": acc["merge_ms"] / steps, "step_ms": step_ms}
    if extra_share:
        share.update(extra_share)
    lift = {"kernel": "k_liftover_ends", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": (tr["liftover_dram_bytes_per_hit"] * hits_per_launch) if "liftover_dram_bytes_per_hit" in tr else None,
            "traffic_source": tr.get("source"),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": acc["lift_bytes"] / n_l,
            "algorithmic_bytes_formula": "SURVEY.md 8(d): per liftover 32 (entry) + 16 (2 checkpoints) + 4*r_ov "
                                         "(runs intersecting the request) + 24 (hit out)",
            "r_ov_mean": acc["lift_window_runs"] / max(1, acc["liftovers"]),
            "touched_bytes_per_launch": acc["lift_touched_bytes"] / n_l,
            "touched_GBps": (acc["lift_touched_bytes"] / 1e9) / (lift_ms / 1e3) if lift_ms > 0 else 0.0,
            "ncu_dram_GBps": (tr["liftover_dram_bytes_per_hit"] * acc["liftovers"] / 1e9) / (lift_ms / 1e3)
            if "liftover_dram_bytes_per_hit" in tr and lift_ms > 0 else None,
            "avg_launch_ms": lift_ms / n_l, "liftovers_per_step": acc["liftovers"] / steps, "step_share": share}
    boxes = acc.get("merge_boxes", 0)
    mk_ms = acc.get("merge_kernel_ms", 0.0)
    merge = None
    if boxes and acc["merge_ms"] > 0:
        bpb = 64  # one 32-byte box record written by the liftover epilogue and read once by the segment merge
        a_phase = (boxes * bpb / 1e9) / (acc["merge_ms"] / 1e3)
        merge = {"kernel": "k_merge_buckets", "bound": "hbm", "unit": "GB/s", "peak": peak,
                 "algorithmic_bytes_per_box": bpb, "boxes_per_step": boxes / steps,
                 "phase_ms_per_step": acc["merge_ms"] / steps, "achieved_phase": a_phase, "frac_phase": a_phase / peak,
                 "kernel_ms_per_step": mk_ms / steps if mk_ms else None,
                 "achieved": (boxes * bpb / 1e9) / (mk_ms / 1e3) if mk_ms else None,
                 "frac": ((boxes * bpb / 1e9) / (mk_ms / 1e3)) / peak if mk_ms else None,
                 "traffic": (tr["merge_dram_bytes_per_box"] * boxes / steps) if "merge_dram_bytes_per_box" in tr else None}
    return lift, merge


def reference_arm(args, ix, name, rank, config, rows):
    if rank != 0:
        return 0
    world_ = CpuWorld(ix, name)
    cfg = world_.cfg
    from impg_b200 import dist as D

    bed = world_.rows_of(ix.synth_bed(cfg, rows, seed=D.rank_seed(2, 0)))
    threads = host_threads()
    try:
        per_step_budget = max(2.0, min(args.cpu_budget, 150.0 / max(1, args.steps + args.warmup)))
        vals, n_used = [], 0
        for i in range(args.warmup + args.steps):
            v, n_used, t, nres = world_.time_reference(bed, per_step_budget, threads)
            if i >= args.warmup:
                vals.append((v, t))
        value = float(np.mean([v for v, _ in vals])) if vals else 0.0
        ms = float(np.mean([t for _, t in vals]) * 1e3) if vals else 0.0
    finally:
        world_.close()
    sample = (f"first {n_used} of the {len(bed)} BED rows{world_.note} per step (reference driver: rows serial, {threads} threads "
              "inside a BFS level, per-hit pread+parse of CIGAR text)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def run_sharded(args, ix, torch, dist, name, rank, local_rank, world, config):
    """N > 1, index sharded by target sequence: rank r generates and uploads only the alignments
    its sequences walk; the batch of world x rows BED rows is one collective call per step."""
    from impg_b200 import dist as D

    cfg, rows = workload_cfg(ix, name)
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
    W = max(args.warmup, 3)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(W):
        shard.query_batch_bed_sharded_device(comm, d_bed.data_ptr(), n, p, stream.cuda_stream)
    shard.query_batch_bed_sharded(comm, h_bed_np, p)
    barrier()
    if rank == 0:
        sampler.start()
    tr0 = comm.traffic()
    acc = {}
    merged = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        r = shard.query_batch_bed_sharded_device(comm, d_bed.data_ptr(), n, p, stream.cuda_stream)
        st = shard.stats()
        acc_stats(acc, st)
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
    slow = D.max_over_ranks([acc["lift_ms"], acc["merge_ms"], acc["fold_ms"], acc["stab_ms"], acc["exchange_ms"]], device="cuda")

    # ---- parity of the NCCL path, visible to the driver: a row sample through the same collective call,
    # every rank's share gathered to rank 0, reassembled and compared bit for bit with the oracle
    parity = None
    cw = None
    if not args.no_cpu_baseline:
        if rank == 0:
            cw = CpuWorld(ix, name)
            sample = spread_sample(cw.rows_of(gbed), PARITY_ROWS)
        else:
            sample = None
        box = [sample]
        dist.broadcast_object_list(box, src=0)
        sample = box[0]
        part = shard.query_batch_bed_sharded(comm, sample, p).columns()
        parts = D.gather_columns({k: np.array(v) for k, v in part.items()})
        if rank == 0:
            parity = cw.parity(sample, ix.merge_shard_columns(parts))
            parity["transport"] = "nccl"
            cw.close()

    # ---- rows over index replicas (extra field): no data-path collective
    replicas = None
    if args.with_replicas:
        replicas = replicas_leg(args, ix, torch, dist, name, rank, local_rank, world, barrier)

    if rank != 0:
        dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    config["parallelism"] = (f"index sharded by target sequence over {world} GPUs (no rank holds a replica), one collective "
                             f"batch of {world} x {rows} rows per step, NCCL hit / frontier / box exchange between hops")
    lift, merge = rooflines(name, acc, args.steps, dev_ms / args.steps, peak, peak_src,
                            {"exchange_ms_within_merge_and_fold": acc["exchange_ms"] / args.steps})
    lift["rank"] = 0
    lift["liftovers_per_step_all_ranks"] = lift_total // args.steps
    line = {"metric": METRIC, "value": n * args.steps / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": n * args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps,
                    "note": "bytes of rank 0; every rank copies the rows in and its own share of the BED rows out"},
            "gpu_launches": int(acc["kernel_launches"]), "roofline": lift, "roofline_merge": merge,
            "cpu_baseline": None, "clocks": sampler.summary(), "bed_rows_out_per_step": int(merged_total),
            "target_sharded": {"n_shards": world, "transport": "nccl", "rows_per_step": n,
                               "exchange_bytes_per_step": int(sent_total) // args.steps,
                               "exchanges_per_step": (tr1["exchanges"] - tr0["exchanges"]) // args.steps,
                               "shard_device_bytes_max": int(bytes_max), "parity_sample": parity,
                               "step_share_slowest_rank": dict(zip(("liftover_ms", "merge_ms", "fold_ms", "stab_ms", "exchange_ms"),
                                                                   [x / args.steps for x in slow]))},
            "rows_over_replicas": replicas,
            "setup": {"generate_s": gen_s, "index_build_s": build_s, "index_device_bytes": shard.device_bytes}}
    emit(line)
    dist.destroy_process_group()
    return 0


def replicas_leg(args, ix, torch, dist, name, rank, local_rank, world, barrier):
    """Rows over index replicas at N > 1: every rank holds the whole index and takes its own rows."""
    from impg_b200 import dist as D

    cfg, rows = workload_cfg(ix, name)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    idx = ix.Impg.from_records(recs, runs, offs, lens, names=names, device=local_rank)
    del runs
    bed = ix.synth_bed(cfg, rows, seed=D.rank_seed(2, rank))
    d_bed = torch.from_numpy(bed.view(np.uint8).copy()).cuda()
    p = mode_params(ix, name)
    stream = torch.cuda.current_stream()
    for _ in range(max(args.warmup, 3)):
        idx.query_batch_bed_device(d_bed.data_ptr(), len(bed), p, stream.cuda_stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        r = idx.query_batch_bed_device(d_bed.data_ptr(), len(bed), p, stream.cuda_stream)
        del r
    e1.record(stream)
    barrier()
    (ms,) = D.max_over_ranks([e0.elapsed_time(e1)], device="cuda")
    total = D.gather_row_counts(len(bed), device="cuda") * args.steps
    return {"value": total / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / args.steps,
            "what": "every rank holds a replica of the index and its own rows; no data-path collective"}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="impgx", choices=["impgx", "reference"])
    ap.add_argument("--workload", default=os.environ.get("IMPGX_BENCH_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (cpu_baseline, parity samples)")
    ap.add_argument("--parallelism", default=os.environ.get("IMPGX_BENCH_PARALLELISM", "targets"),
                    choices=["rows", "targets"], help="N > 1: the layout of the headline value (default: the index "
                    "sharded by target sequence, what north_star names)")
    ap.add_argument("--with-replicas", action="store_true", default=None,
                    help="N > 1, sharded headline: also time rows over index replicas (default for indexes below 8 GB)")
    ap.add_argument("--virtual-shards", type=int, default=int(os.environ.get("IMPGX_BENCH_VIRTUAL_SHARDS", "0")),
                    help="N = 1 only: also time the target-sharded path with this many virtual ranks on the one GPU")
    ap.add_argument("--text-steps", type=int, default=3, help="steps of the e2e_text leg (BED text of every row)")
    ap.add_argument("--profile", action="store_true",
                    help="for ncu: exactly --warmup + --steps device-resident steps, no e2e / cpu legs")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload
    g, c, L, a, eq, rev, seed, rows = WORKLOADS[name]
    if args.with_replicas is None:
        args.with_replicas = name in ("tiny", "c2", "c3")
    k_partners = PARTNERS.get(name, 0) or g - 1
    depth = DEPTH.get(name, 2)
    config = {"workload": f"{name}: synthetic {g}-genome " + ("all-vs-all" if k_partners == g - 1 else f"sparsified (each genome vs {k_partners} others)") +
                          f" PAF, {g * k_partners * c * a} alignments, "
                          f"{rows}-row BED per GPU, " + ("depth 1" if depth == 1 else f"-x -m {depth}") + ", -d 1000 -o bed",
              "genomes": g, "contigs": c, "contig_len": L, "alignments": g * k_partners * c * a, "bed_rows_per_gpu": rows,
              "parallelism": "one GPU holds the whole index", "cache": "inputs larger than L2 "
              "(index run stream >> 126 MB; every step re-reads it from HBM)"}
    import impg_b200 as ix

    if args.impl == "reference":
        return reference_arm(args, ix, name, rank, config, rows)

    # ------------------------------------------------------------ GPU arm
    import torch

    if not torch.cuda.is_available() or ix.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libimpgx has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if args.parallelism == "targets":
            return run_sharded(args, ix, torch, dist, name, rank, local_rank, world, config)
        config["parallelism"] = f"rows sharded over {world} index replicas (no data-path collective)"

    from impg_b200 import dist as D

    cfg, _ = workload_cfg(ix, name)
    t0 = time.time()
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    bed = ix.synth_bed(cfg, rows, seed=D.rank_seed(2, rank))
    gen_s = time.time() - t0
    t0 = time.time()
    idx = ix.Impg.from_records(recs, runs, offs, lens, names=names, device=local_rank)
    build_s = time.time() - t0
    p = mode_params(ix, name)
    n = len(bed)
    W = max(args.warmup, 3)

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
        return r, idx.stats()

    def step_host():
        r = idx.query_batch_bed(h_bed_np, p)
        return r, idx.stats()

    if args.profile:
        for _ in range(args.warmup + args.steps):
            step_device()
        torch.cuda.synchronize()
        emit({"profile_run": True, "workload": name, "steps": args.steps, "warmup": args.warmup,
              "stats_last_step": idx.stats()})
        return 0

    sampler = ClockSampler(local_rank)
    for _ in range(W):
        step_device()
    step_host()

    # ---- value: device-resident timing (CUDA events on the launching stream)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    acc = {}
    merged = 0
    e0.record(stream)
    for _ in range(args.steps):
        r, st = step_device()
        acc_stats(acc, st)
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
    text_steps = max(1, min(args.steps, args.text_steps))
    w0 = time.perf_counter()
    for _ in range(text_steps):
        r, st = step_host()
        text_bytes += idx.format_bed_batch(r, names_arr, length_only=True)
        del r
    torch.cuda.synchronize()
    text_ms = (time.perf_counter() - w0) * 1e3 / text_steps
    barrier()

    dev_ms, e2e_ms, text_ms = D.max_over_ranks([dev_ms, e2e_ms, text_ms], device="cuda")  # slowest rank defines the step
    rows_per_step = D.gather_row_counts(n, device="cuda")
    total_rows = rows_per_step * args.steps
    value = total_rows / (dev_ms / 1e3)
    e2e_value = total_rows / (e2e_ms / 1e3)

    # ---- N = 1 only: the target-sharded path with virtual ranks on the one GPU (in-process transport)
    sharded = None
    if world == 1 and args.virtual_shards > 1:
        sharded = virtual_sharded_leg(args, ix, torch, cfg, recs, runs, offs, lens, bed, p, local_rank, merged)

    if rank != 0:
        dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    lift, merge = rooflines(name, acc, args.steps, dev_ms / args.steps, peak, peak_src)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N=1 only
        if name in SUBWORLD:
            del runs  # the CPU legs run on the contig-0 sub-world, generated on its own
        cw = CpuWorld(ix, name, full=None if name in SUBWORLD else (recs, runs, offs, lens, names))
        try:
            threads = host_threads()
            c_bed = cw.rows_of(bed)
            v, n_used, tsec, nres = cw.time_reference(c_bed, args.cpu_budget, threads)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"first {n_used} of the {len(c_bed)} BED rows{cw.note}, {tsec:.1f} s (oracle port with the "
                                      "reference's cost structure: rows serial, threads inside a BFS level, per-hit pread+parse)"}
            vb, nb, tb = cw.time_batched(c_bed, min(args.cpu_budget, 10.0), threads)
            cpu_baseline["cpu_batched"] = {"value": vb, "unit": UNIT, "cores": threads,
                                           "sample": f"first {nb} rows, {tb:.1f} s",
                                           "what": "second baseline (SURVEY.md 8d): rows in parallel (one thread per row), CIGARs "
                                                   "pre-decoded in RAM; value / this = hardware + kernel gain, this / cpu_baseline = "
                                                   "gain of batching rows and keeping the run stream resident"}
            # parity at full size on a row sample spread over the BED: the oracle's merged BED rows vs the
            # C-ABI's, bit for bit (outside every timed region; the oracle is the checker, never the thing measured)
            sample = spread_sample(c_bed, PARITY_ROWS)
            cpu_baseline["parity_sample"] = cw.parity(sample, idx.query_batch_bed(sample, p).columns())
            if not cpu_baseline["parity_sample"]["bit_exact"]:
                raise SystemExit("bench.py: the CUDA path and the oracle disagree on the parity sample")
        finally:
            cw.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps,
                    "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms / args.steps},
            "e2e_text": {"value": rows_per_step / (text_ms / 1e3), "unit": UNIT, "ms_per_step": text_ms, "steps": text_steps,
                         "text_bytes_per_step": int(text_bytes // text_steps),
                         "what": "e2e plus the BED text of every row (impgx_format_bed_batch on the host cores)"},
            "gpu_launches": int(acc["kernel_launches"]), "roofline": lift, "roofline_merge": merge,
            "cpu_baseline": cpu_baseline, "clocks": sampler.summary(), "bed_rows_out_per_step": int(merged),
            "target_sharded": sharded,
            "setup": {"generate_s": gen_s, "index_build_s": build_s, "index_device_bytes": idx.device_bytes}}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def virtual_sharded_leg(args, ix, torch, cfg, recs, runs, offs, lens, bed, p, local_rank, merged_rows_mode):
    n_shards = args.virtual_shards
    d_gbed = torch.from_numpy(bed.view(np.uint8).copy()).cuda()
    t0 = time.time()
    sh = ix.ShardedImpg.from_records(recs, runs, offs, lens, [local_rank] * n_shards)
    shard_build_s = time.time() - t0

    def traffic():
        t = [c.traffic() for c in sh.comms]
        return {k: sum(x[k] for x in t) for k in t[0]}

    def step():
        return sh._collective(lambda r: sh.shards[r].query_batch_bed_sharded_device(
            sh.comms[r], d_gbed.data_ptr(), len(bed), p, 0))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    tr0 = traffic()
    sh_merged = 0
    accs = [{} for _ in range(n_shards)]
    w0 = time.perf_counter()
    for _ in range(args.steps):
        rs = step()
        sts = sh.stats()
        sh_merged = sum(x["merged"] for x in sts)
        for a, s in zip(accs, sts):
            acc_stats(a, s)
        del rs
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1e3
    tr1 = traffic()
    # parity of the sharded path on the same GPU: a row sample against the unsharded CUDA path is covered by
    # tests/test_gpu_sharded.py; here the row counts must agree
    return {"value": len(bed) * args.steps / (wall_ms / 1e3), "unit": UNIT, "ms_per_step": wall_ms / args.steps,
            "n_shards": n_shards, "transport": "in-process (virtual ranks on one GPU, host wall clock)",
            "rows_per_step": len(bed), "exchange_bytes_per_step": int(tr1["bytes_sent"] - tr0["bytes_sent"]) // args.steps,
            "exchanges_per_step": (tr1["exchanges"] - tr0["exchanges"]) // args.steps,
            "bed_rows_out_per_step": int(sh_merged), "bed_rows_out_match_unsharded": bool(sh_merged == merged_rows_mode),
            "shard_build_s": shard_build_s,
            "step_share_slowest_rank": {k: max(a.get(k, 0) for a in accs) / args.steps
                                        for k in ("lift_ms", "stab_ms", "fold_ms", "merge_ms", "exchange_ms")}}


if __name__ == "__main__":
    sys.exit(main())
