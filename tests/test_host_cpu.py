"""CPU-only tests of the product's host side: the C-ABI library loads and
exports every symbol include/impgx.h declares, the host index builder matches
the oracle (entry order and coitrees visit order), the synthetic generator is
self-consistent, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "impgx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(impgx_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    L = ix.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert L.impgx_abi_version() == 4


def test_parse_cigar_matches_oracle():
    for s in ["10=5I5D", "", "1M", "250=", "2=8D4=2X3=3D228=11I", "0=3X"]:
        assert list(ix.parse_cigar(s)) == list(O.parse_cigar(s))
    with pytest.raises(ix.ImpgxError):
        ix.parse_cigar("10=5Q")


def small_cfg(seed=7, genomes=5, contigs=2, tiles=6, contig_len=60000, eq_mean=40, rev=300):
    return ix.synth_cfg(genomes, contigs, contig_len, tiles, eq_mean, rev, seed)


def test_synth_records_are_consistent_with_their_cigars():
    cfg = small_cfg()
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    assert len(recs) == 5 * 4 * 2 * 6 and len(names) == 10
    for i in range(len(recs)):
        r = runs[int(offs[i]):int(offs[i + 1])]
        op, ln = r >> 29, r & 0x1FFFFFFF
        t = int(ln[op != 2].sum())
        q = int(ln[op != 3].sum())
        assert t == recs[i]["target_end"] - recs[i]["target_start"]
        assert q == recs[i]["query_end"] - recs[i]["query_start"]
        assert recs[i]["query_end"] <= cfg.contig_len and recs[i]["target_end"] <= cfg.contig_len
        assert op[0] == 0 and op[-1] == 0 and (ln > 0).all()
        assert recs[i]["query_id"] != recs[i]["target_id"]
    assert 0 < recs["strand"].mean() < 1
    # deterministic
    recs2, runs2, *_ = ix.synth_generate(small_cfg())
    assert (recs2 == recs).all() and (runs2 == runs).all()
    bed = ix.synth_bed(cfg, 100)
    assert (bed["start"] < bed["end"]).all() and (bed["end"] <= cfg.contig_len).all()
    assert (bed["end"] - bed["start"] >= 1000).all() and (bed["end"] - bed["start"] <= 10000).all()


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 17, 64, 100, 1000])
def test_visit_ranks_are_a_permutation(n):
    r = np.zeros(n, np.uint32)
    ix.lib().impgx_debug_visit_ranks(C.c_size_t(n), r.ctypes.data_as(C.c_void_p))
    assert sorted(r.tolist()) == list(range(n))
    if n <= 8:  # a whole tree of <= SIMPLE_SUBTREE_CUTOFF nodes is one sorted run
        assert r.tolist() == list(range(n))


def test_host_columns_match_oracle_visit_order():
    """The product's (sorted entry columns, visit_rank) reproduce the oracle's
    coitrees restatement hit-for-hit and in order."""
    cfg = small_cfg(seed=11, genomes=6, tiles=9)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    # add overlapping / duplicate-start / self alignments to stress ties
    extra = recs[:40].copy()
    extra["target_start"] = recs[:40]["target_start"]
    recs2 = np.concatenate([recs, extra])
    offs2 = np.concatenate([offs, offs[-1] + (offs[1:41] - offs[0:40]).cumsum()]).astype(np.uint64)
    runs2 = np.concatenate([runs] + [runs[int(offs[i]):int(offs[i + 1])] for i in range(40)])
    n_seqs = len(lens)
    cols = ix.host_columns(recs2, offs2, n_seqs)
    orc = O.Index.build(recs2, runs2, offs2, lens)
    rng = np.random.default_rng(5)
    for _ in range(300):
        t = int(rng.integers(0, n_seqs))
        s = int(rng.integers(0, cfg.contig_len - 10))
        e = s + int(rng.integers(1, 30000))
        lo, hi = int(cols["tgt_off"][t]), int(cols["tgt_off"][t + 1])
        sel = [i for i in range(lo, hi) if cols["e_start"][i] <= e and cols["e_end"][i] >= s]  # closed test
        sel.sort(key=lambda i: cols["e_vrank"][i])
        got = [(int(cols["e_aln"][i]), int((cols["e_flags"][i] >> 1) & 1)) for i in sel]
        assert got == orc.stab_order(t, s, e)
    # prefix max column
    for t in range(n_seqs):
        lo, hi = int(cols["tgt_off"][t]), int(cols["tgt_off"][t + 1])
        assert (cols["e_pmax"][lo:hi] == np.maximum.accumulate(cols["e_end"][lo:hi])).all()
        assert (np.diff(cols["e_start"][lo:hi]) >= 0).all()


def test_compute_fails_loudly_without_gpu():
    if ix.device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = small_cfg()
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    with pytest.raises(ix.ImpgxError) as e:
        ix.Impg.from_records(recs, runs, offs, lens)
    assert e.value.code == ix.E_NO_DEVICE
    with pytest.raises(ix.ImpgxError) as e:
        ix.project_batch([(0, 10)], recs[:1], runs[: int(offs[1])], offs[:2])
    assert e.value.code == ix.E_NO_DEVICE
    with pytest.raises(ix.ImpgxError) as e:
        ix.Impg.from_paf("/nonexistent.paf")
    assert e.value.code == ix.E_NO_DEVICE


def test_write_cigar_text_roundtrip(tmp_path):
    cfg = small_cfg()
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    path = str(tmp_path / "cigars.txt")
    o, l = ix.write_cigar_text(runs, offs, path)
    blob = open(path, "rb").read()
    assert len(blob) == int(o[-1] + l[-1])
    for i in (0, 5, len(recs) - 1):
        txt = blob[int(o[i]):int(o[i] + l[i])].decode()
        assert list(O.parse_cigar(txt)) == list(runs[int(offs[i]):int(offs[i + 1])])


def test_bed_and_range_parsers(tmp_path):
    """parse_bed_file / parse_target_range semantics (reference src/commands/partition.rs:1719-1789)."""
    p = tmp_path / "r.bed"
    p.write_text("chr1\t10\t200\tgeneA\nchr2:0-50\t5\t40\t.\nchr3\t0\t7\n" "chr4\t1\t9\t  spaced  \textra\n")
    rows = ix.parse_bed_file(str(p))
    assert rows == [("chr1", (10, 200), "geneA"), ("chr2:0-50", (5, 40), "chr2:0-50:5-40"),
                    ("chr3", (0, 7), "chr3:0-7"), ("chr4", (1, 9), "spaced")]
    for bad in ("chr1\t10\n", "chr1\t20\t10\n", "chr1\t5\t5\n", "chr1\tx\t10\n", "chr1\t 1\t10\n"):
        p.write_text(bad)
        with pytest.raises(ix.ImpgxError) as e:
            ix.parse_bed_file(str(p))
        assert e.value.code == ix.E_PARSE
    # names that contain ':' need the range appended, the split is on the last ':'
    assert ix.parse_target_range("C4FIXTURE#0#short_floor:0-250:0-250") == (
        "C4FIXTURE#0#short_floor:0-250", (0, 250), "C4FIXTURE#0#short_floor:0-250:0-250")
    assert ix.parse_target_range("A:25-75") == ("A", (25, 75), "A:25-75")
    for bad in ("A", "A:10", "A:10-5", "A:1-2-3", "A:x-9"):
        with pytest.raises(ix.ImpgxError):
            ix.parse_target_range(bad)


def test_cli_fails_loudly_without_gpu():
    if ix.device_count() > 0:
        pytest.skip("a GPU is present")
    import subprocess
    cli = os.path.join(ROOT, "impg_b200", "impgx-query")
    r = subprocess.run([cli, "-a", os.path.join(ROOT, "tests", "golden", "short_floor.paf"), "-r",
                        "C4FIXTURE#0#short_floor:0-250:0-250", "-d", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
    r = subprocess.run([cli, "-a", "x.paf", "-r", "A:0-10"], capture_output=True, text=True)
    assert r.returncode != 0 and "merge-distance is required" in r.stderr


# ---------------------------------------------------------------- target-sharded index (host logic)
def _small_world(seed=3):
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, seed)
    return cfg, ix.synth_generate(cfg)


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 8])
def test_owner_map_is_balanced_and_deterministic(n_ranks):
    cfg, (recs, runs, offs, lens, names) = _small_world()
    owner = ix.assign_owners(recs, offs, len(lens), n_ranks)
    assert owner.max() < n_ranks and (owner == ix.assign_owners(recs, offs, len(lens), n_ranks)).all()
    # weight per rank (entry bytes + stream bytes) within 1.5x of the mean on an all-vs-all set
    nr = np.diff(offs.astype(np.int64))
    w = np.zeros(len(lens))
    np.add.at(w, recs["target_id"], nr)
    np.add.at(w, recs["query_id"], nr)
    load = np.bincount(owner, weights=w, minlength=n_ranks)
    assert load.max() <= 1.5 * load.mean() + 1


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_shard_columns_partition_the_full_index(n_ranks):
    """The shards' entry columns are exactly the full index's, target by target
    (same order, same visit ranks): a target's entries are never split."""
    cfg, (recs, runs, offs, lens, names) = _small_world()
    recs = recs.copy()
    recs["query_id"][:5] = recs["target_id"][:5]  # self alignments: no reversed entry
    n_seqs = len(lens)
    full = ix.host_columns(recs, offs, n_seqs)
    owner = ix.assign_owners(recs, offs, n_seqs, n_ranks)
    seen = 0
    for r in range(n_ranks):
        part = ix.host_columns(recs, offs, n_seqs, owner=owner, rank=r)
        for s in range(n_seqs):
            a, b = int(part["tgt_off"][s]), int(part["tgt_off"][s + 1])
            fa, fb = int(full["tgt_off"][s]), int(full["tgt_off"][s + 1])
            if owner[s] != r:
                assert a == b
                continue
            assert b - a == fb - fa
            for k in ("e_start", "e_end", "e_pmax", "e_vrank", "e_query_id", "e_flags", "e_aln"):
                assert (part[k][a:b] == full[k][fa:fb]).all(), (r, s, k)
            seen += b - a
        # a shard built from only the alignments it needs (relative order kept) is identical
        keep = ix.shard_records(recs, offs, owner, r)
        sub_offs = np.zeros(len(keep) + 1, np.uint64)
        np.cumsum(np.diff(offs.astype(np.int64))[keep], out=sub_offs[1:])
        sub = ix.host_columns(recs[keep], sub_offs, n_seqs, owner=owner, rank=r)
        for k in ("e_start", "e_end", "e_pmax", "e_vrank", "e_query_id", "e_flags"):
            assert (sub[k] == part[k]).all(), (r, k)
    assert seen == len(full["e_start"])


def test_merge_shard_columns_orders_by_row_then_sequence():
    rng = np.random.default_rng(0)
    n_rows, n_seqs, n_ranks = 7, 10, 3
    owner = rng.integers(0, n_ranks, n_seqs)
    rows = np.sort(rng.integers(0, n_rows, 200))
    q = rng.integers(0, n_seqs, 200)
    order = np.lexsort((q, rows))
    rows, q = rows[order], q[order]
    first = np.arange(200, dtype=np.int32)  # position in the full output: must come back as 0..199
    parts = []
    for r in range(n_ranks):
        m = owner[q] == r
        ro = np.zeros(n_rows + 1, np.uint64)
        np.cumsum(np.bincount(rows[m], minlength=n_rows), out=ro[1:])
        parts.append({"row_offsets": ro, "q_id": q[m].astype(np.uint32), "q_first": first[m], "q_last": first[m],
                      "t_id": q[m].astype(np.uint32), "t_first": first[m], "t_last": first[m]})
    out = ix.merge_shard_columns(parts)
    assert out["q_first"].tolist() == list(range(200))
    assert out["row_offsets"].tolist() == np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_rows))]).tolist()


def test_local_comm_group_without_gpu():
    comms = ix.Comm.local_group(3)
    assert [c.rank for c in comms] == [0, 1, 2] and all(c.size == 3 for c in comms)
    assert comms[1].traffic() == {"bytes_sent": 0, "bytes_received": 0, "exchanges": 0}
    with pytest.raises(ix.ImpgxError):
        ix.Comm.local_group(0)


def test_synth_generate_shard_matches_the_full_generator():
    cfg, (recs, runs, offs, lens, names) = _small_world()
    for r in range(3):
        sr, sruns, soffs, sl, sn, owner = ix.synth_generate_shard(cfg, 3, r)
        keep = ix.shard_records(recs, offs, owner, r)
        assert (sr == recs[keep]).all() and sn == names
        assert (np.concatenate([runs[int(offs[k]):int(offs[k + 1])] for k in keep]) == sruns).all()
        a = ix.host_columns(recs, offs, len(lens), owner=owner, rank=r)
        b = ix.host_columns(sr, soffs, len(lens), owner=owner, rank=r)
        for k in ("e_start", "e_end", "e_vrank", "e_flags", "e_query_id"):
            assert (a[k] == b[k]).all()


def test_bench_contig_subworld_is_exact():
    """bench.py times the CPU reference of c4 on the alignments of one contig: rows on that contig
    must give identical results in the sub-world and in the full index."""
    cfg = ix.synth_cfg(6, 3, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    bed = ix.synth_bed(cfg, 120, seed=9, min_len=200, max_len=12000)
    sr, sruns, soffs, _, _ = ix.synth_generate_contig(cfg, 0)  # what bench.py's CpuWorld builds for c4
    rows = bed[bed["target_id"] % cfg.contigs == 0]
    assert len(sr) * 3 == len(recs) and 0 < len(rows) < len(bed)
    full, sub = O.Index.build(recs, runs, offs, lens), O.Index.build(sr, sruns, soffs, lens)
    p = O.make_params(mode=O.MODE_BFS, max_depth=3, store_cigar=True)
    a, ao = full.query_batch(rows[:25], p)
    b, bo = sub.query_batch(rows[:25], p)
    ac, bc = a.columns(), b.columns()
    assert ao.tolist() == bo.tolist() and all((ac[k] == bc[k]).all() for k in ac)


def test_subset_filter_kat_and_fuzz_against_oracle():
    """--subset-sequence-list matching (reference src/subset_filter.rs): the reference's own test on the
    product, then random lists / names against the oracle's restatement."""
    contents = "# comment\nchr1\nchr2\n\nchr1\t\n  chr3  \nHG00097_hap1_hprc_r2_v1.0.1\nHG00098#2#chr5\n"
    for name in ["chr1", "chr1:10-20", "chr3", "HG00097#1#chr7", "HG00097#1", "HG00098#2#chr5"]:
        assert ix.subset_matches(contents, name), name
    assert not ix.subset_matches(contents, "HG00098#1#chr5")
    import random
    rnd = random.Random(3)
    atoms = ["HG1", "HG2", "NA3", "chr1", "chr2", "c", ""]

    def name():
        k = rnd.randrange(8)
        a, b = rnd.choice(atoms), rnd.choice(atoms)
        h = rnd.choice(["1", "2", "", "x", "12"])
        base = [a, f"{a}#{h}#{b}", f"{a}#{h}", f"{a}_hap{h}_{b}", f"{a}_hap{h}", f"{a}#{b}", f" {a} ", f"#{a}"][k]
        return base + (f":{rnd.randrange(100)}-{rnd.randrange(100, 200)}" if rnd.random() < 0.3 else "")

    for _ in range(300):
        text = "".join(name() + rnd.choice(["\n", "\r\n", "\t\n"]) for _ in range(rnd.randrange(1, 6)))
        if rnd.random() < 0.5:
            text = text.rstrip("\n")
        for _ in range(10):
            q = name().strip() or "q"
            assert ix.subset_matches(text, q) == O.subset_matches(text, q)[0], (text, q)


def test_parse_merge_distance_reference_kats():
    # reference src/main.rs:13702-13715
    import ctypes as C

    def parse(t):
        v = C.c_int32(0)
        code = ix.lib().impgx_parse_merge_distance(t.encode(), C.byref(v))
        return v.value if code == 0 else None

    assert parse("50000") == 50_000 and parse("50k") == 50_000 and parse("1m") == 1_000_000
    assert parse("1M") == 1_000_000 and parse("1.5k") == 1_500 and parse(" 7 ") == 7 and parse("0") == 0
    assert parse("2147483647") == 2147483647 and parse("2g") == 2_000_000_000
    for bad in ("10kb", "3g", "", "k", "-5", "1..5k", "abc", "2147483648"):
        assert parse(bad) is None, bad


def test_parse_subsequence_coordinates_kat_and_fuzz():
    from test_oracle_kat import SUBSEQ_KATS
    for name, want in SUBSEQ_KATS:
        assert ix.parse_subsequence_coordinates(name) == want, name
    import random
    rnd = random.Random(9)
    pieces = ["chr1", "a#1#b", ":", "-", "12", "0", "+5", "-7", "99999999999", "2147483647", "2147483648", "x", "", "3-4", ":8-9"]
    for _ in range(2000):
        name = "".join(rnd.choice(pieces) for _ in range(rnd.randrange(1, 6)))
        assert ix.parse_subsequence_coordinates(name) == O.parse_subsequence_coordinates(name), name


def test_new_entry_points_reject_null_arguments_without_crashing():
    """No exception, abort or crash crosses the ABI: NULL / out-of-range arguments come back as codes."""
    import ctypes as C
    L = ix.lib()
    L.impgx_partitions_format_bed.restype = C.c_void_p
    L.impgx_format_bed_batch.restype = C.c_void_p
    L.impgx_impg_seq_name.restype = C.c_char_p
    assert L.impgx_partition(None, None, None) == ix.E_INVALID
    assert L.impgx_partitions_view(None, None) == ix.E_INVALID
    assert L.impgx_partitions_format_bed(None, None, C.c_int64(-1)) is None
    L.impgx_partitions_free(None)
    assert L.impgx_partitioner_new(None, None, C.c_uint32(3), None, None) == ix.E_INVALID
    assert L.impgx_partitioner_next(None, None, None, None) == ix.E_INVALID
    assert L.impgx_partitioner_feed(None, C.c_size_t(0), None, None, None) == ix.E_INVALID
    assert L.impgx_partitioner_finish(None, None) == ix.E_INVALID
    L.impgx_partitioner_free(None)
    assert L.impgx_impg_open(None, None) == ix.E_INVALID
    assert L.impgx_impg_records(None, None, None, None, None) == ix.E_INVALID
    assert L.impgx_impg_version(None) == 0 and L.impgx_impg_num_seqs(None) == 0
    assert L.impgx_impg_seq_name(None, C.c_uint32(0)) is None
    L.impgx_impg_close(None)
    assert L.impgx_impg_write(None, C.c_size_t(0), C.c_int(1), None) == ix.E_INVALID
    assert L.impgx_index_from_impg(None, None, C.c_size_t(0), C.c_int(0), None) == ix.E_INVALID
    assert L.impgx_format_bed_batch(None, None, None, None) is None
    assert L.impgx_subset_matches(None, None) == ix.E_INVALID
    L.impgx_subset_mask.restype = C.c_long
    assert L.impgx_subset_mask(None, None, None) == ix.E_INVALID
    assert L.impgx_parse_merge_distance(None, None) == ix.E_INVALID
    assert L.impgx_parse_subsequence_coordinates(None, None, C.c_size_t(0), None) == ix.E_INVALID
    assert L.impgx_index_set_original_coordinates(None, C.c_int(1)) == ix.E_INVALID
    # a stepper over zero sequences is done at once
    pp = ix.make_partition_params(window_size=10, merge_distance=0)
    st = ix.Partitioner(np.zeros(0, np.uint64), pp)
    assert st.next() is None and st.finish().rows() == []
