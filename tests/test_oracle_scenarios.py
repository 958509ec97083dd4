"""Replays the behavioural scenarios of the reference's CLI integration tests
(reference tests/test_transitive_integrity.rs:74-767) on the oracle: tiny inline
PAFs, `query -d 0 [-x -m N] --min-transitive-len 0 -r REGION`, default BED
output. The assertions are the reference's own (set membership / bands)."""
import numpy as np
import pytest

import _oracle as O


def P(a, al, as_, ae, strand, b, bl, bs, be, cg):
    return f"{a}\t{al}\t{as_}\t{ae}\t{strand}\t{b}\t{bl}\t{bs}\t{be}\t100\t100\t60\tcg:Z:{cg}"


def make_index(tmp_path, lines, name="test.paf"):
    p = tmp_path / name
    p.write_text("".join(l + "\n" for l in lines))
    return O.Index.from_paf(str(p))


def query_bed(idx, region, transitive=False, max_depth=2, d=0, min_transitive_len=0, dfs=False):
    """impg query -d D -r REGION [-x -m N] --min-transitive-len L  (BED rows)."""
    name, rng = region.rsplit(":", 1)  # parse_target_range splits on the LAST ':'
    s, e = (int(x) for x in rng.split("-"))
    mode = O.MODE_QUERY if not transitive else (O.MODE_DFS if dfs else O.MODE_BFS)
    params = O.make_params(mode=mode, max_depth=max_depth, min_transitive_len=min_transitive_len, merge_distance=d)
    res = idx.perform_query(idx.seq_id(name), s, e, params)
    text = idx.format(res, "bed", f"{name}:{s}-{e}", d)
    rows = []
    for line in text.splitlines():
        f = line.split("\t")
        rows.append((f[0], int(f[1]), int(f[2]), f[3], f[5]))
    return rows


def test_non_overlapping_regions_stay_separate(tmp_path):  # :74
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 0, 100, "100="),
                                P("A", 1000, 500, 600, "+", "C", 1000, 0, 100, "100=")])
    names = {r[0] for r in query_bed(idx, "A:0-100", transitive=True)}
    assert "A" in names and "B" in names and "C" not in names
    names = {r[0] for r in query_bed(idx, "A:500-600", transitive=True)}
    assert "A" in names and "C" in names and "B" not in names


def test_transitive_coordinate_accuracy(tmp_path):  # :155
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 0, 100, "100="),
                                P("B", 1000, 0, 100, "+", "C", 1000, 0, 100, "100=")])
    rows = query_bed(idx, "A:25-75", transitive=True)
    assert {r[0] for r in rows} == {"A", "B", "C"}
    for name, s, e, _, _ in rows:
        assert 45 <= e - s <= 55
        if name != "A":
            assert 20 <= s <= 30 and 70 <= e <= 80


def test_bidirectional_symmetry(tmp_path):  # :226 (exact)
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 200, 300, "100=")])
    rows = query_bed(idx, "A:0-100")
    assert [(r[0], r[1], r[2]) for r in rows if r[0] == "B"] == [("B", 200, 300)]
    rows = query_bed(idx, "B:200-300")
    assert [(r[0], r[1], r[2]) for r in rows if r[0] == "A"] == [("A", 0, 100)]


def test_reverse_strand(tmp_path):  # :297
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "-", "B", 1000, 0, 100, "100=")])
    rows = [r for r in query_bed(idx, "A:0-50") if r[0] == "B"]
    assert rows
    for _, s, e, _, strand in rows:
        assert (s + e) / 2 >= 50
        assert strand == "-"
    assert [(r[1], r[2]) for r in rows] == [(50, 100)]


def test_no_collapse_at_depth_3(tmp_path):  # :348
    idx = make_index(tmp_path, [P("A", 2000, 0, 100, "+", "B", 1000, 0, 100, "100="),
                                P("A", 2000, 1000, 1100, "+", "C", 1000, 0, 100, "100="),
                                P("B", 1000, 0, 100, "+", "D", 1000, 0, 100, "100="),
                                P("C", 1000, 0, 100, "+", "D", 1000, 500, 600, "100=")])
    d_rows = [r for r in query_bed(idx, "A:0-100", transitive=True, max_depth=3) if r[0] == "D"]
    assert d_rows and all(r[1] < 200 for r in d_rows)
    d_rows = [r for r in query_bed(idx, "A:1000-1100", transitive=True, max_depth=3) if r[0] == "D"]
    assert d_rows and all(r[1] >= 400 for r in d_rows)


def test_indel_accuracy(tmp_path):  # :452
    idx = make_index(tmp_path, [P("A", 1000, 0, 110, "+", "B", 1000, 0, 100, "50=10I50=")])
    b = [r for r in query_bed(idx, "A:0-50") if r[0] == "B"]
    assert len(b) == 1 and b[0][1] <= 5 and 45 <= b[0][2] <= 55
    b = [r for r in query_bed(idx, "A:60-110") if r[0] == "B"]
    assert len(b) == 1 and 45 <= b[0][1] <= 55 and b[0][2] >= 95


def test_two_alignments_stay_separate(tmp_path):  # :535
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 0, 100, "100="),
                                P("A", 1000, 0, 100, "+", "B", 1000, 500, 600, "100=")])
    b = sorted((r[1], r[2]) for r in query_bed(idx, "A:0-100") if r[0] == "B")
    assert b == [(0, 100), (500, 600)]


def test_empty_region_only_self(tmp_path):  # :648
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 0, 100, "100=")])
    rows = query_bed(idx, "A:500-600")
    assert [(r[0], r[1], r[2]) for r in rows] == [("A", 500, 600)]


@pytest.mark.parametrize("dfs", [False, True])
def test_depth_limit(tmp_path, dfs):  # :688
    idx = make_index(tmp_path, [P("A", 1000, 0, 100, "+", "B", 1000, 0, 100, "100="),
                                P("B", 1000, 0, 100, "+", "C", 1000, 0, 100, "100="),
                                P("C", 1000, 0, 100, "+", "D", 1000, 0, 100, "100=")])
    assert {r[0] for r in query_bed(idx, "A:0-100", transitive=True, max_depth=1, dfs=dfs)} == {"A", "B"}
    names = {r[0] for r in query_bed(idx, "A:0-100", transitive=True, max_depth=2, dfs=dfs)}
    assert "C" in names and "D" not in names
    assert {r[0] for r in query_bed(idx, "A:0-100", transitive=True, max_depth=0, dfs=dfs)} == {"A", "B", "C", "D"}


def test_bfs_proximity_and_piece_rules(tmp_path):
    """H1 of SURVEY.md: the fold is order-sensitive through
    min_distance_between_ranges and the per-piece min_transitive_len gate
    (reference src/impg.rs:2513-2556)."""
    # two A→B alignments whose projections on B overlap partially
    idx = make_index(tmp_path, [P("A", 5000, 0, 1000, "+", "B", 5000, 0, 1000, "1000="),
                                P("A", 5000, 0, 1000, "+", "B", 5000, 500, 1500, "1000="),
                                P("B", 5000, 0, 1500, "+", "C", 5000, 0, 1500, "1500=")])
    name = "A"
    params = O.make_params(mode=O.MODE_BFS, max_depth=2, min_transitive_len=101, min_dist=10)
    res = idx.perform_query(idx.seq_id(name), 0, 1000, params).tuples()
    ids = {n: idx.seq_id(n) for n in "ABC"}
    on_c = sorted((r[1], r[2]) for r in res if r[0] == ids["C"])
    # frontier on B after level 1 is the merged [0,1500) → one hit on C
    assert on_c == [(0, 1500)]
    assert res[0][:3] == (ids["A"], 0, 1000)


def test_merge_query_orientation_rule():
    """merge_query_adjusted_intervals (reference src/main.rs:12474-12560):
    cross-strand merge keeps the orientation of the longer span, ties keep
    the current one; tie ORDER matters."""
    r = O.Results.from_tuples([(0, 0, 100, 9, 0, 100), (0, 150, 50, 9, 0, 100), (0, 160, 50, 9, 0, 100)])
    r.merge_query(0, True)
    assert [(t[1], t[2]) for t in r.tuples()] == [(0, 160)]
    r = O.Results.from_tuples([(0, 0, 100, 9, 0, 100), (0, 50, 170, 9, 0, 100), (0, 210, 50, 9, 0, 100)])
    r.merge_query(0, True)
    # [0,100)+ then [50,170)+ → [0,170)+ ; then '-' [50,210) len 160 <= 170 keeps '+'
    assert [(t[1], t[2]) for t in r.tuples()] == [(0, 210)]
    r = O.Results.from_tuples([(0, 0, 100, 9, 0, 100), (0, 300, 50, 9, 0, 100)])
    r.merge_query(0, True)
    assert [(t[1], t[2]) for t in r.tuples()] == [(300, 0)]
    # strands kept apart
    r = O.Results.from_tuples([(0, 0, 100, 9, 0, 100), (0, 150, 50, 9, 0, 100), (0, 60, 200, 9, 0, 100)])
    r.merge_query(0, False)
    assert [(t[1], t[2]) for t in r.tuples()] == [(0, 100), (150, 50), (60, 200)]
    # --no-merge (d = -1) with merge_strands only sorts
    r = O.Results.from_tuples([(1, 5, 9, 9, 0, 4), (0, 7, 3, 9, 0, 4), (0, 3, 8, 9, 0, 5)])
    r.merge_query(-1, True)
    assert [(t[0], t[1], t[2]) for t in r.tuples()] == [(0, 3, 8), (0, 7, 3), (1, 5, 9)]


def test_merge_2d():
    """merge_adjusted_intervals_gap_2d (reference src/main.rs:12858-13011)."""
    rows = [(1, 0, 100, 2, 0, 100), (1, 120, 200, 2, 130, 210), (1, 500, 600, 2, 140, 240),
            (1, 130, 220, 3, 0, 90), (1, 100, 0, 2, 0, 100)]
    r = O.Results.from_tuples(rows)
    r.merge_2d(50)
    got = r.tuples()
    # first two chain (q_gap 20, t_gap 30, forward progress); third is too far on q;
    # different target / strand never merge; output order = first member's index
    assert [(t[0], t[1], t[2], t[3], t[4], t[5]) for t in got] == [
        (1, 0, 200, 2, 0, 210), (1, 500, 600, 2, 140, 240), (1, 130, 220, 3, 0, 90), (1, 100, 0, 2, 0, 100)]
    r = O.Results.from_tuples(rows)
    r.merge_2d(-1)
    assert len(r) == 5


def test_bed_bedpe_paf_writers(tmp_path):
    idx = make_index(tmp_path, [P("A", 1000, 0, 110, "+", "B", 1000, 0, 100, "50=10I50="),
                                P("A", 1000, 200, 300, "-", "C", 1000, 0, 100, "60=2X38=")])
    a = idx.seq_id("A")
    res = idx.perform_query(a, 0, 300, O.make_params(mode=O.MODE_QUERY, store_cigar=True))
    res.drop_first()
    text = idx.format(res, "bedpe", "reg", 0)
    lines = sorted(text.splitlines())
    assert lines == [
        "B\t0\t100\tA\t0\t110\treg\t0\t+\t+\tgi:f:0.990099\tbi:f:0.909091",
        "C\t0\t100\tA\t200\t300\treg\t0\t-\t+\tgi:f:0.98\tbi:f:0.98",
    ]
    res = idx.perform_query(a, 0, 300, O.make_params(mode=O.MODE_QUERY, store_cigar=True))
    res.drop_first()
    lines = sorted(idx.format(res, "paf", "reg", 0).splitlines())
    assert lines[0] == "B\t1000\t0\t100\t+\tA\t1000\t0\t110\t100\t110\t255\tgi:f:0.990099\tbi:f:0.909091\tcg:Z:50=10D50=\tan:Z:reg"
    assert lines[1] == "C\t1000\t0\t100\t-\tA\t1000\t200\t300\t98\t100\t255\tgi:f:0.98\tbi:f:0.98\tcg:Z:38=2X60=\tan:Z:reg"
