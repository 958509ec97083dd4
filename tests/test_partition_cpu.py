"""partition on the CPU: the oracle's restatement of partition_alignments
(reference src/commands/partition.rs:158-712) against the reference's own scenario, and the
product's HOST logic (the window / mask / missing bookkeeping of libimpgx's partitioner stepper)
driven by the oracle's masked transitive queries — no device involved. The device path
(impgx_partition) is compared with the same oracle in test_gpu_partition.py."""
import os
import tempfile

import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix


def small_world(seed=7, genomes=5, contigs=2, tiles=6, contig_len=60000, eq_mean=40, rev=300):
    cfg = ix.synth_cfg(genomes, contigs, contig_len, tiles, eq_mean, rev, seed)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    oidx = O.Index.build(recs, runs, offs, lens, names=names)
    return oidx, lens, names


def patchy_world(seed, keep=0.35, **kw):
    """A random subset of the alignments: homology is patchy, so windows leave slivers, masks cut
    later windows and the sliver-extension / boundary rules fire."""
    cfg = ix.synth_cfg(kw.get("genomes", 6), kw.get("contigs", 2), kw.get("contig_len", 50000), kw.get("tiles", 9),
                       40, kw.get("rev", 300), seed)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    rng = np.random.default_rng(seed)
    sel = np.flatnonzero(rng.random(len(recs)) < keep)
    runs2 = np.concatenate([runs[int(offs[i]):int(offs[i + 1])] for i in sel])
    offs2 = np.zeros(len(sel) + 1, np.uint64)
    offs2[1:] = np.cumsum([int(offs[i + 1]) - int(offs[i]) for i in sel])
    recs2 = recs[sel].copy()
    return (recs2, runs2, offs2, lens, names), O.Index.build(recs2, runs2, offs2, lens, names=names)


def norm(rows):
    return [(p, s, min(a, b), max(a, b)) for p, s, a, b in rows]


def step_with_oracle(oidx, lens, names, kw, bed_rows=False):
    """The product's stepper, every window answered by the oracle's masked transitive query."""
    pp = ix.make_partition_params(**kw)
    st = ix.Partitioner(lens, pp, names=names)
    mode = O.MODE_DFS if kw.get("transitive_dfs") else O.MODE_BFS
    windows = []
    while True:
        nx = st.next()
        if nx is None:
            break
        (t, s, e), mask = nx
        windows.append((t, s, e))
        qp = O.make_params(mode=mode, max_depth=kw.get("max_depth", 2), min_transitive_len=kw.get("min_transitive_len", 101),
                           min_dist=kw.get("min_distance_between_ranges", 10), masked_regions=mask,
                           merge_distance=kw["merge_distance"], merge_strands=True)
        if bed_rows:  # what impgx_partition feeds for -d >= 0: the rows of output_results_bed's two merges
            res, _ = oidx.query_batch(np.array([(t, s, e)], dtype=O.RANGE_DTYPE), qp, bed_merge=True)
        else:
            res = oidx.perform_query(t, s, e, qp)
        c = res.columns()
        st.feed(c["q_id"], c["q_first"], c["q_last"])
    return st.finish(), windows


def test_reference_scenario_partition_window_separation():
    # tests/test_transitive_integrity.rs:592-646: `partition -d 100000 -w 2000 -o bed` must not merge the windows
    lines = ["A\t10000\t0\t1000\t+\tB\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000=",
             "A\t10000\t5000\t6000\t+\tC\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000="]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "test.paf")
        open(path, "w").write("\n".join(lines) + "\n")
        oidx = O.Index.from_paf(path)
    out = oidx.partition(O.make_partition_params(window_size=2000, merge_distance=100000))
    bed = [l for l in out["bed"].splitlines() if l]
    assert len(bed) >= 2
    assert len({l.split("\t")[3] for l in bed}) >= 2  # several partitions, not one merged window
    # hand-derived first partition: window A:0-2000 reaches B:0-1000 through the first alignment
    a, b = oidx.seq_id("A"), oidx.seq_id("B")
    raw = oidx.partition(O.make_partition_params(window_size=2000, merge_distance=100000, rehome_singletons=False))
    assert norm([r for r in raw["rows"] if r[0] == 0]) == sorted([(0, a, 0, 2000), (0, b, 0, 1000)])
    assert norm([r for r in raw["rows"] if r[0] == 1]) == [(1, a, 2000, 4000)]  # a singleton sliver ...
    assert (0, a, 2000, 4000) in norm(out["rows"])  # ... that the default rehoming gives to its left flank
    assert out["partitioned_bp"] == out["total_bp"] == 20000
    # and the product's host logic walks the same windows to the same partitions
    lens = np.array([oidx.seq_len(i) for i in range(oidx.n_seqs)], np.uint64)
    got, windows = step_with_oracle(oidx, lens, None, dict(window_size=2000, merge_distance=100000))
    assert windows == out["windows"] and got.rows() == norm(out["rows"])


CASES = [
    dict(window_size=20000, merge_distance=1000),
    dict(window_size=20000, merge_distance=0, min_missing_size=0, min_boundary_distance=0),
    dict(window_size=7000, merge_distance=100000, rehome_singletons=False),
    dict(window_size=15000, merge_distance=-1),
    dict(window_size=15000, merge_distance=500, transitive_dfs=True),
    dict(window_size=25000, merge_distance=1000, selection_mode="total"),
    dict(window_size=25000, merge_distance=1000, selection_mode="sample"),
    dict(window_size=25000, merge_distance=1000, selection_mode="haplotype,#", max_depth=1),
    dict(window_size=30000, merge_distance=200, starting_seqs=[3, 0], max_depth=3),
    dict(window_size=100000, merge_distance=1000, max_depth=0, min_transitive_len=2000),
    dict(window_size=9000, merge_distance=50, min_missing_size=12000, min_boundary_distance=8000),
]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_stepper_matches_oracle_partition(k):
    kw = CASES[k]
    oidx, lens, names = small_world(seed=11 + k)
    want = oidx.partition(O.make_partition_params(**kw))
    got, windows = step_with_oracle(oidx, lens, names, kw)
    assert windows == want["windows"]
    assert got.rows() == norm(want["rows"])
    assert (got.n_partitions, got.partitioned_bp, got.total_bp, got.n_windows) == (
        want["n_partitions"], want["partitioned_bp"], want["total_bp"], len(want["windows"]))
    if kw["merge_distance"] >= 0:
        # fed with the BED rows of the window (the device's merged output) instead of the raw result list
        got2, windows2 = step_with_oracle(oidx, lens, names, kw, bed_rows=True)
        assert windows2 == windows and got2.rows() == got.rows()


@pytest.mark.parametrize("k", range(len(CASES)))
def test_stepper_matches_oracle_partition_patchy(k):
    kw = dict(CASES[k])
    kw["window_size"] = max(3000, kw["window_size"] // 3)
    (recs, runs, offs, lens, names), oidx = patchy_world(seed=101 + k, keep=0.2 + 0.05 * (k % 5))
    want = oidx.partition(O.make_partition_params(**kw))
    assert len(want["windows"]) > 3 and want["partitioned_bp"] == want["total_bp"]
    got, windows = step_with_oracle(oidx, lens, names, kw)
    assert windows == want["windows"]
    assert got.rows() == norm(want["rows"])
    assert (got.n_partitions, got.partitioned_bp, got.n_windows) == (want["n_partitions"], want["partitioned_bp"],
                                                                     len(want["windows"]))
    if kw["merge_distance"] >= 0:
        got2, windows2 = step_with_oracle(oidx, lens, names, kw, bed_rows=True)
        assert windows2 == windows and got2.rows() == got.rows()


def test_group_selection_modes_with_other_separators():
    # names without the separator, a multi-character separator, the empty separator (str::split("") semantics)
    (recs, runs, offs, lens, names), _ = patchy_world(seed=42, keep=0.3)
    for sep_names, mode in (([n.replace("#", "::") for n in names], "haplotype,::"),
                            ([n.replace("#", "") for n in names], "sample"),
                            (names, "sample,"), (names, "haplotype,"), (names, "sample,1#")):
        oidx = O.Index.build(recs, runs, offs, lens, names=sep_names)
        kw = dict(window_size=9000, merge_distance=300, selection_mode=mode)
        want = oidx.partition(O.make_partition_params(**kw))
        got, windows = step_with_oracle(oidx, lens, sep_names, kw)
        assert windows == want["windows"] and got.rows() == norm(want["rows"]), mode


def test_partitions_tile_every_sequence_exactly_once():
    oidx, lens, names = small_world(seed=3, genomes=6, rev=500)
    for kw in (dict(window_size=10000, merge_distance=1000), dict(window_size=33333, merge_distance=0, max_depth=1)):
        out = oidx.partition(O.make_partition_params(**kw))
        assert out["partitioned_bp"] == out["total_bp"] == int(lens.sum())
        per_seq = {}
        for _, s, a, b in norm(out["rows"]):
            per_seq.setdefault(s, []).append((a, b))
        assert sorted(per_seq) == list(range(len(lens)))
        for s, iv in per_seq.items():
            iv.sort()
            assert iv[0][0] == 0 and iv[-1][1] == int(lens[s])
            assert all(iv[i][1] == iv[i + 1][0] for i in range(len(iv) - 1))  # disjoint and gap-free


def test_rehoming_only_moves_singletons_next_to_a_flank():
    oidx, lens, names = small_world(seed=5)
    kw = dict(window_size=7000, merge_distance=100)
    a = oidx.partition(O.make_partition_params(rehome_singletons=False, **kw))
    b = oidx.partition(O.make_partition_params(rehome_singletons=True, **kw))
    ia = sorted((s, x, y) for _, s, x, y in norm(a["rows"]))
    ib = sorted((s, x, y) for _, s, x, y in norm(b["rows"]))
    assert ia == ib  # the same intervals, possibly under another partition number
    assert len({r[0] for r in b["rows"]}) <= len({r[0] for r in a["rows"]})


def test_stepper_rejects_bad_arguments():
    lens = np.array([1000, 2000], np.uint64)
    with pytest.raises(ix.ImpgxError):
        ix.Partitioner(lens, ix.make_partition_params(window_size=0, merge_distance=0))
    with pytest.raises(ix.ImpgxError) as e:
        ix.Partitioner(lens, ix.make_partition_params(window_size=100, merge_distance=0, selection_mode="widest"))
    assert "Invalid selection mode" in str(e.value)
    with pytest.raises(ix.ImpgxError):  # sample mode needs names
        ix.Partitioner(lens, ix.make_partition_params(window_size=100, merge_distance=0, selection_mode="sample"))
    st = ix.Partitioner(lens, ix.make_partition_params(window_size=1500, merge_distance=0))
    (t, s, e), _ = st.next()
    assert (t, s, e) == (1, 0, 2000)  # longest; the 500-base tail joins the first window
    with pytest.raises(ix.ImpgxError):
        st.next()  # the window was not fed back
    with pytest.raises(ix.ImpgxError):
        st.feed([7], [0], [10])  # unknown sequence


def test_partition_without_device_fails_loudly():
    if ix.device_count() > 0:
        pytest.skip("a CUDA device is present")
    # there is no CPU fallback behind impgx_partition either
    cfg = ix.synth_cfg(3, 1, 20000, 2, 40, 100, 1)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    with pytest.raises(ix.ImpgxError) as e:
        ix.Impg.from_records(recs, runs, offs, lens, names=names)
    assert e.value.code == ix.E_NO_DEVICE


def test_stepper_matches_oracle_on_random_worlds_and_parameters():
    """A sweep over random worlds / parameter combinations (the host logic of impgx_partition against the
    oracle's literal restatement of partition_alignments)."""
    import random
    rnd = random.Random(2024)
    modes = ["longest", "total", "sample", "haplotype", "sample,#1#", "haplotype,#"]
    for trial in range(24):
        world, oidx = patchy_world(seed=500 + trial, keep=rnd.choice([0.15, 0.3, 0.6, 1.0]), genomes=rnd.choice([3, 5, 7]),
                                   contigs=rnd.choice([1, 2, 3]), contig_len=rnd.choice([20000, 45000]),
                                   tiles=rnd.choice([3, 7, 11]), rev=rnd.choice([0, 200, 700]))
        lens, names = world[3], world[4]
        kw = dict(window_size=rnd.choice([2500, 6000, 15000, 50000]), merge_distance=rnd.choice([-1, 0, 100, 3000, 100000]),
                  selection_mode=rnd.choice(modes), min_missing_size=rnd.choice([0, 500, 3000, 20000]),
                  min_boundary_distance=rnd.choice([0, 700, 3000]), transitive_dfs=rnd.random() < 0.3,
                  max_depth=rnd.choice([0, 1, 2, 3]), min_transitive_len=rnd.choice([0, 101, 1500]),
                  min_distance_between_ranges=rnd.choice([0, 10, 400]), rehome_singletons=rnd.random() < 0.7)
        if rnd.random() < 0.3:
            kw["starting_seqs"] = [rnd.randrange(len(lens)) for _ in range(rnd.randrange(1, 4))]
        want = oidx.partition(O.make_partition_params(**kw))
        got, windows = step_with_oracle(oidx, lens, names, kw)
        assert windows == want["windows"], (trial, kw)
        assert got.rows() == norm(want["rows"]), (trial, kw)
        assert got.partitioned_bp == want["partitioned_bp"] == want["total_bp"], (trial, kw)


def test_stepper_reproduces_the_committed_partition_golden():
    """The product's partition bookkeeping, windows answered by the oracle, on the reference's fixture PAFs against
    tests/golden/golden_partition.json — raw result lists and BED-merged rows (what the device returns) alike."""
    import hashlib
    import json
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    pgold = json.load(open(os.path.join(gold, "golden_partition.json")))
    for fixture in sorted(pgold):
        oidx = O.Index.from_paf(os.path.join(gold, fixture))
        names = [oidx.seq_name(i) for i in range(oidx.n_seqs)]
        lens = np.array([oidx.seq_len(i) for i in range(oidx.n_seqs)], np.uint64)
        for c in pgold[fixture]["partition"]:
            kw = c["params"]
            for bed_rows in ((False, True) if kw["merge_distance"] >= 0 else (False,)):
                got, windows = step_with_oracle(oidx, lens, names, kw, bed_rows=bed_rows)
                text = "".join(f"{names[s]}\t{a}\t{b}\t{p}\n" for p, s, a, b in got.rows())
                assert hashlib.sha256(text.encode()).hexdigest() == c["sha256"], (fixture, kw, bed_rows)
                assert (len(windows), got.n_partitions) == (c["windows"], c["n_partitions"])


def test_partition_with_any_answerer_matches_the_oracle():
    """impg_b200.partition_with — the driver behind ShardedImpg.partition / dist.partition_sharded — on the CPU:
    every window answered by the oracle's merged BED rows (what a sharded index returns after its parts are
    merged); the partitions are the oracle's."""
    oidx, lens, names = small_world(seed=11)

    class Meta:  # what partition_with needs of an index: the sequence table
        n_seqs = len(lens)

        def seq_len(self, i):
            return int(lens[i])

        def seq_name(self, i):
            return names[i]

    asked = []

    def answer(window, gp):
        asked.append(tuple(int(x) for x in window[0]))
        mo = np.ctypeslib.as_array((ix.C.c_uint64 * (len(lens) + 1)).from_address(gp.mask_offsets))
        nm = int(mo[-1])
        mr = np.ctypeslib.as_array((ix.C.c_int32 * max(2 * nm, 2)).from_address(gp.mask_ranges))
        qp = O.make_params(mode=O.MODE_DFS if gp.mode == ix.MODE_DFS else O.MODE_BFS, max_depth=gp.max_depth,
                           min_transitive_len=gp.min_transitive_len, min_dist=gp.min_distance_between_ranges,
                           masked_regions=(mo.copy(), mr.copy()), merge_distance=gp.merge_distance, merge_strands=True)
        res, _ = oidx.query_batch(window, qp, bed_merge=True)
        return res.columns()

    for kw in (dict(window_size=20000, merge_distance=1000), dict(window_size=9000, merge_distance=0, max_depth=3),
               dict(window_size=25000, merge_distance=3000, transitive_dfs=True)):
        asked.clear()
        got = ix.partition_with(Meta(), ix.make_partition_params(**kw), answer)
        want = oidx.partition(O.make_partition_params(**kw))
        assert got.rows() == norm(want["rows"]), kw
        assert asked == [tuple(w) for w in want["windows"]]
    with pytest.raises(ix.ImpgxError) as e:
        ix.partition_with(Meta(), ix.make_partition_params(window_size=20000, merge_distance=-1), answer)
    assert e.value.code == ix.E_UNSUPPORTED
