"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs (bit-exact: all integer work), against the committed
golden fixtures, and — at larger sizes — through size-independent properties."""
import hashlib
import json
import os

import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix
from test_oracle_kat import LIFTOVER_KATS

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = json.load(open(os.path.join(GOLD, "golden.json")))


def rec(t_start, t_end, q_start, q_end, strand, reversed_entry=False, qid=0, tid=1):
    r = np.zeros(1, dtype=ix.RECORD_DTYPE)
    r["query_id"], r["target_id"] = qid, tid
    r["query_start"], r["query_end"], r["target_start"], r["target_end"] = q_start, q_end, t_start, t_end
    r["strand"] = 1 if strand else 0
    r["reserved"] = 1 if reversed_entry else 0
    return r


def test_liftover_kernel_kats():
    """K2 against the reference's own 14 known-answer vectors (src/impg.rs:2981-3156)."""
    recs, reqs, runs, offs = [], [], [], [0]
    for req, r, cg, exp in LIFTOVER_KATS:
        recs.append(rec(*r))
        reqs.append(req)
        c = O.cigar(cg)
        runs.append(c)
        offs.append(offs[-1] + len(c))
    reqs.append((65, 65))  # empty request → None (src/impg.rs:3029-3033)
    recs.append(rec(0, 100, 50, 200, False))
    c = O.cigar("10=5I5D50=50I35=")
    runs.append(c)
    offs.append(offs[-1] + len(c))
    got = ix.project_batch(reqs, np.concatenate(recs), np.concatenate(runs), np.array(offs, np.uint64))
    for (req, r, cg, exp), g in zip(LIFTOVER_KATS, got):
        assert g is not None
        assert (g[0], g[1]) == (exp[0], exp[1])
        assert O.cigar_str(g[2]) == exp[2]
        if exp[3] is not None:
            assert (g[3], g[4]) == (exp[3], exp[4])
    assert got[-1] is None
    got = ix.project_batch(reqs, np.concatenate(recs), np.concatenate(runs), np.array(offs, np.uint64), want_cigar=False)
    for (req, r, cg, exp), g in zip(LIFTOVER_KATS, got):
        assert g is not None and (g[0], g[1]) == (exp[0], exp[1])
        if exp[3] is not None:
            assert (g[3], g[4]) == (exp[3], exp[4])
    assert got[-1] is None


def random_cigar(rng, n_ops, zero_len=False):
    ops = []
    for _ in range(n_ops):
        op = "=XIDM"[int(rng.integers(0, 5))]
        ln = int(rng.integers(0 if zero_len else 1, 40))
        ops.append(O.run(op, ln))
    return np.array(ops, np.uint32)


@pytest.mark.parametrize("seed,n_ops,zero_len", [(1, 5, False), (2, 40, False), (3, 33, True), (4, 200, False),
                                                  (5, 1000, True), (6, 64, False), (7, 1, True)])
def test_liftover_kernel_random_vs_oracle(seed, n_ops, zero_len):
    """Random CIGARs (incl. zero-length ops, all five op kinds), all four entry
    orientations, requests straddling every boundary: bit-exact vs the oracle."""
    rng = np.random.default_rng(seed)
    recs, reqs, runs, offs, expect = [], [], [], [0], []
    for k in range(400):
        cg = random_cigar(rng, int(rng.integers(1, n_ops + 1)), zero_len)
        op, ln = cg >> 29, cg & 0x1FFFFFFF
        tlen, qlen = int(ln[op != 2].sum()), int(ln[op != 3].sum())
        strand = bool(rng.integers(0, 2))
        reversed_entry = bool(rng.integers(0, 2))
        ts, qs = int(rng.integers(0, 1000)), int(rng.integers(0, 1000))
        if reversed_entry:  # entry keyed by the original query: roles swapped
            e_t, e_q = (qs, qs + qlen), (ts, ts + tlen)
        else:
            e_t, e_q = (ts, ts + tlen), (qs, qs + qlen)
        span = max(e_t[1] - e_t[0], 1)
        a = e_t[0] + int(rng.integers(-5, span + 5))
        b = a + int(rng.integers(1, span + 10))
        if k % 7 == 0:
            a, b = e_t[0], e_t[1]
        recs.append(rec(e_t[0], e_t[1], e_q[0], e_q[1], strand, reversed_entry))
        reqs.append((a, b))
        runs.append(cg)
        offs.append(offs[-1] + len(cg))
        walk = O.invert(cg, strand) if reversed_entry else cg
        expect.append(O.project((a, b), (e_t[0], e_t[1], e_q[0], e_q[1], strand), walk))
    got = ix.project_batch(reqs, np.concatenate(recs), np.concatenate(runs), np.array(offs, np.uint64))
    for i, (g, e) in enumerate(zip(got, expect)):
        if e is None:
            assert g is None, (i, reqs[i], g)
        else:
            assert g is not None, (i, reqs[i], e)
            assert (g[0], g[1], g[3], g[4]) == (e[0], e[1], e[3], e[4]), (i, reqs[i])
            assert list(g[2]) == list(e[2]), (i, reqs[i], O.cigar_str(g[2]), O.cigar_str(e[2]))
    # the endpoint kernel (8 lanes per hit, boundary blocks only) must agree on the coordinates
    got = ix.project_batch(reqs, np.concatenate(recs), np.concatenate(runs), np.array(offs, np.uint64),
                           want_cigar=False)
    for i, (g, e) in enumerate(zip(got, expect)):
        if e is None:
            assert g is None, ("ends", i, reqs[i], g)
        else:
            assert g is not None, ("ends", i, reqs[i], e)
            assert (g[0], g[1], g[3], g[4]) == (e[0], e[1], e[3], e[4]), ("ends", i, reqs[i])


# ---------------------------------------------------------------- whole path
def build_both(recs, runs, offs, lens, names=None, bidirectional=True):
    orc = O.Index.build(recs, runs, offs, lens, bidirectional=bidirectional, names=names)
    gpu = ix.Impg.from_records(recs, runs, offs, lens, names=names, bidirectional=bidirectional)
    return orc, gpu


def compare_raw(orc, gpu, ranges, o_params, g_params, check_cigar=False):
    ores, ooffs = orc.query_batch(ranges, o_params)
    oc = ores.columns()
    gres = gpu.query_batch(ranges, g_params)
    gc = gres.columns()
    assert gc["row_offsets"].tolist() == ooffs.tolist()
    for k in ("q_id", "q_first", "q_last", "t_id", "t_first", "t_last"):
        assert (gc[k] == oc[k]).all(), k
    if check_cigar:
        assert gc["cigar_offsets"].tolist() == oc["cigar_offsets"].tolist()
        assert (gc["cigar_runs"] == oc["cigar_runs"]).all()
    return len(oc["q_id"])


def compare_bed(orc, gpu, ranges, o_params, g_params):
    ores, ooffs = orc.query_batch(ranges, o_params, bed_merge=True)
    oc = ores.columns()
    gc = gpu.query_batch_bed(ranges, g_params).columns()
    assert gc["row_offsets"].tolist() == ooffs.tolist()
    for k in ("q_id", "q_first", "q_last"):
        assert (gc[k] == oc[k]).all(), k
    return len(oc["q_id"])


def params_pair(**kw):
    o = O.make_params(mode=kw.get("mode", 0), max_depth=kw.get("max_depth", 2),
                      min_transitive_len=kw.get("min_transitive_len", 101), min_dist=kw.get("min_dist", 10),
                      min_output_length=kw.get("min_output_length", -1), store_cigar=kw.get("store_cigar", False),
                      min_identity=kw.get("min_identity", float("nan")), subset_mask=kw.get("subset_mask"),
                      merge_distance=kw.get("merge_distance", 0), merge_strands=kw.get("merge_strands", True),
                      masked_regions=kw.get("masked_regions"))
    mol = kw.get("min_output_length", -1)
    mi = kw.get("min_identity", float("nan"))
    g = ix.make_params(mode=kw.get("mode", 0), max_depth=kw.get("max_depth", 2),
                       min_transitive_len=kw.get("min_transitive_len", 101),
                       min_distance_between_ranges=kw.get("min_dist", 10),
                       min_output_length=None if mol < 0 else mol, store_cigar=kw.get("store_cigar", False),
                       min_identity=None if mi != mi else mi, subset_mask=kw.get("subset_mask"),
                       merge_distance=kw.get("merge_distance", 0), merge_strands=kw.get("merge_strands", True),
                       masked_regions=kw.get("masked_regions"))
    return o, g


@pytest.fixture(scope="module")
def small():
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    bed = ix.synth_bed(cfg, 300, seed=9, min_len=200, max_len=12000)
    return cfg, orc, gpu, bed


def test_query_depth1_raw(small):
    cfg, orc, gpu, bed = small
    n = compare_raw(orc, gpu, bed, *params_pair(mode=0))
    assert n > 300 * 5


def test_query_depth1_with_cigar_and_filters(small):
    cfg, orc, gpu, bed = small
    compare_raw(orc, gpu, bed, *params_pair(mode=0, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, bed, *params_pair(mode=0, min_output_length=3000))
    compare_raw(orc, gpu, bed, *params_pair(mode=0, min_identity=0.955, store_cigar=True), check_cigar=True)
    mask = np.zeros(12, np.uint8)
    mask[[1, 4, 5, 9]] = 1
    compare_raw(orc, gpu, bed, *params_pair(mode=0, subset_mask=mask))


@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_bfs_raw(small, depth):
    cfg, orc, gpu, bed = small
    compare_raw(orc, gpu, bed[:120], *params_pair(mode=1, max_depth=depth))


@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_dfs_raw(small, depth):
    """Transitive DFS (reference src/impg.rs:2057-2309): lock-step over rows on the device."""
    cfg, orc, gpu, bed = small
    compare_raw(orc, gpu, bed[:60], *params_pair(mode=2, max_depth=depth))


def test_dfs_options_and_bed(small):
    cfg, orc, gpu, bed = small
    b = bed[:40]
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=3, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=0, min_transitive_len=0, min_dist=0))
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=3, min_output_length=2500, min_identity=0.95))
    compare_bed(orc, gpu, b, *params_pair(mode=2, max_depth=2, merge_distance=1000))
    compare_bed(orc, gpu, b, *params_pair(mode=2, max_depth=0, merge_distance=0, merge_strands=False))


def test_bfs_options(small):
    cfg, orc, gpu, bed = small
    b = bed[:100]
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=2, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, min_transitive_len=0, min_dist=0))
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, min_transitive_len=2000, min_dist=500))
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=2, min_output_length=2500))
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, min_identity=0.95))
    mask = np.zeros(12, np.uint8)
    mask[[0, 2, 3, 7, 8]] = 1
    compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, subset_mask=mask))


@pytest.mark.parametrize("mode,depth", [(0, 1), (1, 2), (1, 3)])
@pytest.mark.parametrize("d,merge_strands", [(0, True), (1000, True), (1000, False), (-1, True), (-1, False)])
def test_bed_merge(small, mode, depth, d, merge_strands):
    cfg, orc, gpu, bed = small
    compare_bed(orc, gpu, bed[:150], *params_pair(mode=mode, max_depth=depth, merge_distance=d,
                                                   merge_strands=merge_strands))


def random_mask(rng, n_seqs, seq_len, density):
    """A SortedRanges per sequence: sorted, disjoint, non-touching [start, end) pairs."""
    regions = {}
    for s in range(n_seqs):
        pos, out = int(rng.integers(0, 3000)), []
        while pos < seq_len and rng.random() < density:
            ln = int(rng.integers(50, 6000))
            out.append((pos, min(pos + ln, seq_len)))
            pos += ln + int(rng.integers(1, 9000))
        regions[s] = out
    return ix.mask_csr(regions, n_seqs)


@pytest.mark.parametrize("mode", [1, 2, 4, 5])
def test_masked_regions(small, mode):
    """masked_regions of the transitive queries (src/impg.rs:2331-2373; what partition passes,
    src/commands/partition.rs:359-391): the visited sets start from the mask, the self interval
    becomes the unmasked pieces of the range. BFS, DFS and both MultiImpg walks."""
    cfg, orc, gpu, bed = small
    if mode >= 3:  # the MultiImpg walks are checked against the oracle's MultiImpg (one sub-index)
        recs, runs, offs, lens, names = ix.synth_generate(cfg)
        orc = O.MultiIndex.build(recs, runs, offs, lens, np.zeros(len(recs), np.uint32), 1)
    rng = np.random.default_rng(100 + mode)
    b = bed[:80]
    for density, depth in ((0.5, 2), (0.9, 3), (0.0, 2), (0.97, 0)):
        mask = random_mask(rng, 12, 60000, density)
        compare_raw(orc, gpu, b, *params_pair(mode=mode, max_depth=depth, masked_regions=mask))
    mask = random_mask(rng, 12, 60000, 0.8)
    compare_raw(orc, gpu, b, *params_pair(mode=mode, max_depth=2, masked_regions=mask, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, b, *params_pair(mode=mode, max_depth=3, masked_regions=mask, min_output_length=2500,
                                           min_transitive_len=0, min_dist=0))
    compare_bed(orc, gpu, b, *params_pair(mode=mode, max_depth=2, masked_regions=mask, merge_distance=1000))
    compare_bed(orc, gpu, b, *params_pair(mode=mode, max_depth=2, masked_regions=mask, merge_distance=0,
                                           merge_strands=False))
    # a mask that covers every row entirely: nothing is output
    full = ix.mask_csr({s: [(0, 60000)] for s in range(12)}, 12)
    res = gpu.query_batch(b, params_pair(mode=mode, masked_regions=full)[1])
    assert res.n_results == 0
    compare_raw(orc, gpu, b, *params_pair(mode=mode, masked_regions=full))
    # malformed masks are rejected
    bad = (np.array([0, 2] + [2] * 11, np.uint64), np.array([10, 50, 40, 90], np.int32))
    with pytest.raises(ix.ImpgxError) as e:
        gpu.query_batch(b, params_pair(mode=mode, masked_regions=bad)[1])
    assert e.value.code == ix.E_INVALID


def test_unidirectional_and_self_alignments():
    cfg = ix.synth_cfg(4, 1, 40000, 5, 25, 500, 21)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    # self alignments get no reversed entry (src/impg.rs:1584)
    recs = recs.copy()
    recs["query_id"][:6] = recs["target_id"][:6]
    for bidir in (True, False):
        orc, gpu = build_both(recs, runs, offs, lens, names, bidirectional=bidir)
        bed = ix.synth_bed(cfg, 80, seed=4, min_len=300, max_len=9000)
        compare_raw(orc, gpu, bed, *params_pair(mode=0))
        compare_raw(orc, gpu, bed, *params_pair(mode=1, max_depth=3))


def test_edge_cases():
    cfg = ix.synth_cfg(3, 1, 30000, 3, 20, 500, 5)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    o, g = params_pair(mode=1, max_depth=2)
    # empty batch
    r = gpu.query_batch(np.zeros(0, ix.RANGE_DTYPE), g)
    assert r.n_rows == 0 and r.n_results == 0
    # a range with no alignments at all → only the self interval
    gap = np.array([(0, 9995, 10001)], ix.RANGE_DTYPE)
    compare_raw(orc, gpu, gap, o, g)
    # touching ranges (closed-interval visit, rejected by the liftover): boundaries of an alignment
    t0 = recs[0]
    edges = np.array([(t0["target_id"], max(t0["target_start"] - 50, 0), t0["target_start"]),
                      (t0["target_id"], t0["target_end"], t0["target_end"] + 50),
                      (t0["target_id"], t0["target_start"], t0["target_start"] + 1),
                      (t0["target_id"], t0["target_end"] - 1, t0["target_end"]),
                      (t0["target_id"], 0, 30000)], ix.RANGE_DTYPE)
    edges = edges[edges["start"] < edges["end"]]
    compare_raw(orc, gpu, edges, *params_pair(mode=0, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, edges, *params_pair(mode=1, max_depth=0, min_transitive_len=0))
    # invalid rows are rejected like the reference (src/main.rs:11620-11639)
    for bad in ([(99, 0, 10)], [(0, 10, 10)], [(0, 20, 10)], [(0, 0, 30001)], [(0, -5, 10)]):
        with pytest.raises(ix.ImpgxError) as e:
            gpu.query_batch(np.array(bad, ix.RANGE_DTYPE), g)
        assert e.value.code == ix.E_INVALID


@pytest.mark.parametrize("fixture", sorted(GOLDEN))
def test_fixture_pafs_against_golden(fixture):
    """The reference's own PAF fixtures through impgx_index_from_paf + the BED
    entry point, against golden text produced by the oracle in the build container."""
    path = os.path.join(GOLD, fixture)
    gpu = ix.Impg.from_paf(path)
    orc = O.Index.from_paf(path)
    assert [gpu.seq_name(i) for i in range(gpu.n_seqs)] == [orc.seq_name(i) for i in range(orc.n_seqs)]
    for c in GOLDEN[fixture]:
        sid = gpu.seq_id(c["seq"])
        region = f"{c['seq']}:{c['start']}-{c['end']}"
        rng = np.array([(sid, c["start"], c["end"])], ix.RANGE_DTYPE)
        if c["format"] == "bed":
            g = ix.make_params(mode=c["mode"], max_depth=c["max_depth"], min_transitive_len=0, merge_distance=c["d"])
            text = gpu.format_bed(gpu.query_batch_bed(rng, g), 0, region)
        else:  # bedpe / paf: raw results with CIGARs from the device, merge + text on the host side of the library
            g = ix.make_params(mode=c["mode"], max_depth=c["max_depth"], min_transitive_len=0, store_cigar=True)
            res = gpu.query_batch(rng, g)
            text = (ix.format_bedpe if c["format"] == "bedpe" else ix.format_paf)(gpu, res, 0, region, c["d"])
        assert hashlib.sha256(text.encode()).hexdigest() == c["sha256"], (fixture, c)
    # raw results incl. CIGARs vs the oracle for every sequence
    ranges = np.array([(s, 0, orc.seq_len(s)) for s in range(orc.n_seqs)], ix.RANGE_DTYPE)
    compare_raw(orc, gpu, ranges, *params_pair(mode=0, store_cigar=True), check_cigar=True)
    compare_raw(orc, gpu, ranges, *params_pair(mode=1, max_depth=0, min_transitive_len=0, store_cigar=True),
                check_cigar=True)


def test_bedpe_paf_text_vs_oracle(small):
    """BEDPE / PAF lines (CIGAR merge with f32 surgery, gi/bi formatting) against the oracle's writers."""
    cfg, orc, gpu, bed = small
    rows = bed[:40]
    for mode, depth in ((0, 1), (1, 2)):
        for d in (0, 500, -1):
            o, g = params_pair(mode=mode, max_depth=depth, store_cigar=True)
            gres = gpu.query_batch(rows, g)
            for r in range(len(rows)):
                name = f"r{r}"
                for fmt, fn in (("bedpe", ix.format_bedpe), ("paf", ix.format_paf)):
                    ores = orc.perform_query(int(rows[r]["target_id"]), int(rows[r]["start"]), int(rows[r]["end"]), o)
                    ores.drop_first()
                    assert fn(gpu, gres, r, name, d) == orc.format(ores, fmt, name, d), (mode, d, r, fmt)


def test_bed_text_of_a_whole_batch_vs_oracle(small):
    """impgx_format_bed_batch (every row of the BED file, formatted on all host cores) against the
    oracle's output_results_bed, row by row, and against the per-row formatter."""
    cfg, orc, gpu, bed = small
    rows = bed[:60]
    names = [f"r{k}" if k % 3 else f"{orc.seq_name(int(rows[k]['target_id']))}:{int(rows[k]['start'])}-{int(rows[k]['end'])}"
             for k in range(len(rows))]
    for d, ms in ((1000, True), (0, False), (-1, True)):
        o, g = params_pair(mode=1, max_depth=2, merge_distance=d, merge_strands=ms)
        gres = gpu.query_batch_bed(rows, g)
        text = gpu.format_bed_batch(gres, names)
        assert text == "".join(gpu.format_bed(gres, r, names[r]) for r in range(len(rows)))
        want = ""
        for r in range(len(rows)):
            ores = orc.perform_query(int(rows[r]["target_id"]), int(rows[r]["start"]), int(rows[r]["end"]), o)
            want += orc.format(ores, "bed", names[r], d, merge_strands=ms)
        assert text == want
    assert gpu.format_bed_batch(gpu.query_batch_bed(rows[:0], g), []) == ""


def test_medium_scale_properties():
    """~20k alignments, 2k rows, depth 2: oracle parity on a row sample plus
    size-independent properties on everything (BED rows sorted, disjoint beyond
    the merge distance, idempotent under a second merge)."""
    cfg = ix.synth_cfg(12, 2, 400000, 20, 60, 150, 17)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    bed = ix.synth_bed(cfg, 2000, seed=23)
    o, g = params_pair(mode=1, max_depth=2, merge_distance=1000)
    gc = gpu.query_batch_bed(bed, g).columns()
    sample = np.arange(0, 2000, 40)
    ores, ooffs = orc.query_batch(bed[sample], o, bed_merge=True)
    oc = ores.columns()
    for j, r in enumerate(sample):
        a, b = int(gc["row_offsets"][r]), int(gc["row_offsets"][r + 1])
        oa, ob = int(ooffs[j]), int(ooffs[j + 1])
        assert b - a == ob - oa
        for k in ("q_id", "q_first", "q_last"):
            assert (gc[k][a:b] == oc[k][oa:ob]).all()
    ro = gc["row_offsets"]
    lo = np.minimum(gc["q_first"], gc["q_last"]).astype(np.int64)
    hi = np.maximum(gc["q_first"], gc["q_last"]).astype(np.int64)
    for r in range(0, 2000, 7):
        a, b = int(ro[r]), int(ro[r + 1])
        assert b > a
        q, s, e = gc["q_id"][a:b].astype(np.int64), lo[a:b], hi[a:b]
        key = q * (1 << 32) + s
        assert (np.diff(key) > 0).all()
        same = q[1:] == q[:-1]
        assert (s[1:][same] > e[:-1][same] + 1000).all()
    st = gpu.stats()
    assert st["liftovers"] > 0 and st["kernel_launches"] > 0 and st["lift_ms"] > 0


def test_cli_driver_matches_golden_and_oracle(tmp_path):
    """impgx-query (the `impg query` look-alike) end to end: BASELINE config 1 plus a BED batch
    with -x, all three output formats, and a gzip-compressed PAF."""
    import gzip
    import subprocess
    cli = os.path.join(os.path.dirname(GOLD), "..", "impg_b200", "impgx-query")
    paf = os.path.join(GOLD, "short_floor.paf")
    out = subprocess.run([cli, "-a", paf, "-r", "C4FIXTURE#0#short_floor:0-250:0-250", "-d", "0",
                          "--min-transitive-len", "0", "-o", "bed"], capture_output=True, text=True, check=True).stdout
    c = [x for x in GOLDEN["short_floor.paf"] if x["seq"] == "C4FIXTURE#0#short_floor:0-250" and x["start"] == 0 and
         x["end"] == 250 and x["mode"] == 0 and x["d"] == 0 and x["format"] == "bed"][0]
    assert out == c["text"]
    # BED batch, transitive, every format, against the oracle's writers; PAF read through gzip
    gz = tmp_path / "short_floor.paf.gz"
    gz.write_bytes(gzip.compress(open(paf, "rb").read()))
    orc = O.Index.from_paf(paf)
    bed = tmp_path / "q.bed"
    rows = [(orc.seq_name(s), 0, orc.seq_len(s), "." if s % 2 else f"row{s}") for s in range(orc.n_seqs)]
    bed.write_text("".join(f"{a}\t{b}\t{c}\t{d}\n" for a, b, c, d in rows))
    for fmt in ("bed", "bedpe", "paf"):
        got = subprocess.run([cli, "-a", str(gz), "-b", str(bed), "-x", "-m", "3", "-d", "50", "--min-transitive-len", "0",
                              "-o", fmt], capture_output=True, text=True, check=True).stdout
        want = ""
        for a, b, c, d in rows:
            name = d if d != "." else f"{a}:{b}-{c}"
            p = O.make_params(mode=O.MODE_BFS, max_depth=3, min_transitive_len=0, merge_distance=50,
                              store_cigar=fmt != "bed")
            res = orc.perform_query(orc.seq_id(a), b, c, p)
            if fmt != "bed":
                res.drop_first()
            want += orc.format(res, fmt, name, 50)
        assert got == want, fmt
    # default --min-transitive-len 101 rejects short rows like the reference (src/main.rs:10387-10403)
    r = subprocess.run([cli, "-a", paf, "-r", "C4FIXTURE#0#short_floor:0-250:0-50", "-d", "0"], capture_output=True,
                       text=True)
    assert r.returncode != 0 and "below minimum of 101 bp" in r.stderr


def test_full_size_c3_properties_and_sample_parity():
    """BASELINE configs[2] at full size (50 genomes, 999,600 alignments, 10,000-row BED, -x -m 2,
    -d 1000 -o bed): oracle parity on a row sample, size-independent properties on every row
    (rows sorted by (sequence, start), merged rows further apart than -d, every row contains its
    own query range, device-resident and host paths agree, checksum stable across two runs)."""
    cfg = ix.synth_cfg(50, 8, 2500000, 51, 100, 100, 1)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    assert len(recs) == 999600
    gpu = ix.Impg.from_records(recs, runs, offs, lens, names=names)
    bed = ix.synth_bed(cfg, 10000, seed=2)
    o, g = params_pair(mode=1, max_depth=2, merge_distance=1000)
    gc = gpu.query_batch_bed(bed, g).columns()
    st = gpu.stats()
    assert st["liftovers"] > 5e7 and st["lift_window_runs"] > st["liftovers"]
    ro = gc["row_offsets"].astype(np.int64)
    assert len(ro) == 10001 and (np.diff(ro) > 0).all()
    lo = np.minimum(gc["q_first"], gc["q_last"]).astype(np.int64)
    hi = np.maximum(gc["q_first"], gc["q_last"]).astype(np.int64)
    rowid = np.repeat(np.arange(10000), np.diff(ro))
    key = (rowid.astype(np.int64) << 44) + (gc["q_id"].astype(np.int64) << 32) + lo
    assert (np.diff(key) > 0).all()                                   # sorted, no duplicates
    same = (rowid[1:] == rowid[:-1]) & (gc["q_id"][1:] == gc["q_id"][:-1])
    assert (lo[1:][same] > hi[:-1][same] + 1000).all()                # nothing left to merge
    # the self interval survives inside a merged row on the query's own sequence
    for r in range(0, 10000, 97):
        a, b = ro[r], ro[r + 1]
        m = gc["q_id"][a:b] == bed["target_id"][r]
        assert ((lo[a:b][m] <= bed["start"][r]) & (hi[a:b][m] >= bed["end"][r])).any()
    # idempotence / determinism: a second run gives the same bytes
    gc2 = gpu.query_batch_bed(bed, g).columns()
    for k in ("row_offsets", "q_id", "q_first", "q_last"):
        assert (gc[k] == gc2[k]).all()
    # oracle parity on a sample of rows (the oracle needs seconds per row at this size)
    orc = O.Index.build(recs, runs, offs, lens, names=names)
    sample = np.unique(np.linspace(0, 9999, 64).astype(np.int64))
    ores, ooffs = orc.query_batch(bed[sample], o, bed_merge=True)
    oc = ores.columns()
    for j, r in enumerate(sample):
        a, b = ro[r], ro[r + 1]
        oa, ob = int(ooffs[j]), int(ooffs[j + 1])
        assert b - a == ob - oa
        for k in ("q_id", "q_first", "q_last"):
            assert (gc[k][a:b] == oc[k][oa:ob]).all()


def test_concurrent_callers_share_one_index(small):
    """The handle is Send + Sync like the reference's ImpgIndex (src/impg_index.rs:21): concurrent
    callers on one index are serialised inside the library and every one gets the right answer."""
    import threading
    cfg, orc, gpu, bed = small
    jobs = [(bed[:60], params_pair(mode=1, max_depth=2, merge_distance=1000)),
            (bed[60:140], params_pair(mode=0, merge_distance=0)),
            (bed[140:200], params_pair(mode=2, max_depth=2, merge_distance=1000)),
            (bed[200:260], params_pair(mode=1, max_depth=3, merge_distance=-1))]
    want = []
    for b, (o, g) in jobs:
        res, offs = orc.query_batch(b, o, bed_merge=True)
        want.append((offs.tolist(), res.columns()))
    got, errs = [None] * len(jobs), []

    def work(k):
        try:
            for _ in range(3):
                got[k] = gpu.query_batch_bed(jobs[k][0], jobs[k][1][1]).columns()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for k, (offs, oc) in enumerate(want):
        assert got[k]["row_offsets"].tolist() == offs
        for c in ("q_id", "q_first", "q_last"):
            assert (got[k][c] == oc[c]).all(), (k, c)


def test_device_resident_entry_on_a_side_stream(small):
    """impgx_query_batch_bed_device with rows in HBM and a caller-owned stream returns the same
    rows as the host entry (columns read back from the device view)."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    cfg, orc, gpu, bed = small
    o, g = params_pair(mode=1, max_depth=2, merge_distance=1000)
    want = gpu.query_batch_bed(bed, g).columns()
    stream = torch.cuda.Stream()
    d_bed = torch.from_numpy(bed.view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        res = gpu.query_batch_bed_device(d_bed.data_ptr(), len(bed), g, stream.cuda_stream)
    stream.synchronize()
    v = res.view
    assert v.n_rows == len(bed) and v.n_results == len(want["q_id"])

    def dev(ptr, n, dtype):
        out = torch.empty(n, dtype=dtype, device="cuda")
        # cudaMemcpy resolved through libimpgx's own dependency on libcudart (device -> device)
        rc = ix.lib().cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(n * out.element_size()), C.c_int(3))
        assert rc == 0
        return out.cpu().numpy()

    assert (dev(v.row_offsets, v.n_rows + 1, torch.int64).astype(np.uint64) == want["row_offsets"]).all()
    assert (dev(v.q_id, v.n_results, torch.int32).astype(np.uint32) == want["q_id"]).all()
    assert (dev(v.q_first, v.n_results, torch.int32) == want["q_first"]).all()
    assert (dev(v.q_last, v.n_results, torch.int32) == want["q_last"]).all()


@pytest.mark.parametrize("min_class", [1, 2, 3, 4])
def test_every_segment_kernel_variant(small, monkeypatch, min_class):
    """The fused BED merge picks a kernel variant by segment size (warp: <= 128 / 256 boxes; CTA of
    256 / 128 / 512 threads: <= 512 / 1024 / 4096). IMPGX_SEG_MIN_CLASS pushes the small segments of the
    test world into the larger classes so that every variant is checked against the oracle."""
    cfg, orc, gpu, bed = small
    monkeypatch.setenv("IMPGX_SEG_MIN_CLASS", str(min_class))
    b = bed[:120]
    compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=2, merge_distance=1000))
    compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=0, merge_distance=0, merge_strands=False))
    compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=3, merge_distance=-1, merge_strands=True))
    compare_bed(orc, gpu, b, *params_pair(mode=0, merge_distance=50000))


def test_sort_based_segment_merge_still_matches(small, monkeypatch):
    """The bucket merge is the default of the direct BED path; IMPGX_MERGE_SORTED selects the older fused merge
    behind one global sort (still used on a sharded index and with the identity filter)."""
    cfg, orc, gpu, bed = small
    monkeypatch.setenv("IMPGX_MERGE_SORTED", "1")
    b = bed[:150]
    compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=2, merge_distance=1000))
    compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=0, merge_distance=0, merge_strands=False))
    compare_bed(orc, gpu, b, *params_pair(mode=0, merge_distance=-1, merge_strands=True))


def test_large_segments_and_global_fallback(monkeypatch):
    """A world with few sequences and many alignments per pair: hundreds to thousands of boxes per
    (row, sequence) segment, so the CTA variants run at their real sizes; with a tiny segment limit the
    batch falls back to the global two-sort merge. All against the oracle."""
    cfg = ix.synth_cfg(3, 1, 400000, 400, 12, 300, 17)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    rows = np.array([(s, 1000 + 37 * k, 390000 - 91 * k) for s in range(3) for k in range(4)], ix.RANGE_DTYPE)
    for d, ms in ((1000, True), (0, False), (-1, True)):
        n = compare_bed(orc, gpu, rows, *params_pair(mode=1, max_depth=3, merge_distance=d, merge_strands=ms))
    st = gpu.stats()
    assert st["liftovers"] / len(rows) > 2000  # thousands of boxes per row over 3 sequences
    # the same 12 rows through the batched path (calls of <= 64 rows try the single-launch walk first)
    monkeypatch.setenv("IMPGX_NO_SMALL_BFS", "1")
    for d, ms in ((1000, True), (0, False), (-1, True)):
        compare_bed(orc, gpu, rows, *params_pair(mode=1, max_depth=3, merge_distance=d, merge_strands=ms))
    monkeypatch.delenv("IMPGX_NO_SMALL_BFS")
    monkeypatch.setenv("IMPGX_MERGE_GLOBAL", "1")
    compare_bed(orc, gpu, rows, *params_pair(mode=1, max_depth=3, merge_distance=1000))


def test_segment_beyond_the_largest_class_falls_back():
    """One (row, sequence) segment with ~12,000 boxes (> 4096): the fused merge declines the batch and
    the global two-sort path produces the same rows as the oracle."""
    cfg = ix.synth_cfg(2, 1, 1600000, 6000, 8, 200, 23)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    rows = np.array([(0, 0, 1600000), (1, 5, 1599000), (0, 1000, 2000)], ix.RANGE_DTYPE)
    compare_bed(orc, gpu, rows, *params_pair(mode=0, merge_distance=10))
    compare_bed(orc, gpu, rows, *params_pair(mode=0, merge_distance=0, merge_strands=False))
    assert gpu.stats()["liftovers"] > 20000
