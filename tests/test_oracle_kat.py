"""Pins the oracle against every known-answer test the reference holds for the
hot path (reference src/impg.rs:2981-3264, src/paf.rs:368-415,
src/alignment_record.rs). Vectors are restated as numbers."""
import os

import numpy as np
import pytest

import _oracle as O
from _oracle import cigar, cigar_str

F, R = False, True
MIXED = "10=5I5D50=50I35="

# (req, (t_start,t_end,q_start,q_end,strand_rev), cigar, expected (q_start,q_end,cigar,t_start,t_end) or partial)
LIFTOVER_KATS = [
    ((100, 200), (100, 200, 0, 100, F), "100=", (0, 100, "100=", 100, 200)),      # impg.rs:2982-2990
    ((100, 200), (100, 200, 0, 100, R), "100=", (100, 0, "100=", 100, 200)),      # :2993-3001
    ((0, 100), (0, 100, 50, 200, F), MIXED, (50, 200, MIXED, 0, 100)),            # :3014-3018
    ((50, 55), (0, 100, 50, 200, F), MIXED, (100, 105, "5=", 50, 55)),            # :3019-3023
    ((50, 64), (0, 100, 50, 200, F), MIXED, (100, 114, "14=", 50, 64)),           # :3024-3028
    ((50, 65), (0, 100, 50, 200, F), MIXED, (100, 165, "15=50I", 50, 65)),        # :3034-3039
    ((50, 66), (0, 100, 50, 200, F), MIXED, (100, 166, "15=50I1=", 50, 66)),      # :3040-3049
    ((70, 95), (0, 100, 50, 200, F), MIXED, (170, 195, "25=", 70, 95)),           # :3050-3054
    ((100, 200), (100, 200, 100, 200, F), "100=", (100, 200, "100=", 100, 200)),  # :3059-3069
    ((100, 200), (100, 200, 100, 200, R), "100=", (200, 100, "100=", 100, 200)),  # :3073-3083
    ((50, 150), (50, 150, 50, 160, F), "50=10I50=", (50, 160, "50=10I50=", None, None)),   # :3087-3098
    ((50, 150), (50, 150, 50, 140, F), "50=10D40=", (50, 140, "50=10D40=", None, None)),   # :3102-3113
    ((150, 250), (100, 200, 200, 300, R), "50=10D10I40=", (250, 200, "10D10I40=", None, None)),  # :3117-3134
    ((0, 10), (0, 50, 0, 40, F), "10=20D8=1X1=10I10=", (0, 10, "10=", 0, 10)),    # :3138-3156
]


@pytest.mark.parametrize("req,rec,cg,exp", LIFTOVER_KATS)
def test_liftover_kat(oracle, req, rec, cg, exp):
    got = oracle.project(req, rec, cigar(cg))
    assert got is not None
    assert got[0] == exp[0] and got[1] == exp[1]
    assert cigar_str(got[2]) == exp[2]
    if exp[3] is not None:
        assert got[3] == exp[3] and got[4] == exp[4]


def test_liftover_empty_target_range_not_emitted(oracle):
    # impg.rs:3029-3033: (65,65) used to give the pure insertion; now None.
    assert oracle.project((65, 65), (0, 100, 50, 200, F), cigar(MIXED)) is None


def test_parse_cigar_basic(oracle):
    # impg.rs:3159-3168
    assert cigar_str(oracle.parse_cigar("10=5I5D")) == "10=5I5D"
    assert list(oracle.parse_cigar("10=5I5D")) == [10, (2 << 29) | 5, (3 << 29) | 5]
    assert len(oracle.parse_cigar("")) == 0
    with pytest.raises(ValueError):
        oracle.parse_cigar("10=5Q")  # CigarOp::new panics (impg.rs:88)


def test_invert_forward(oracle):
    # impg.rs:3201-3220
    assert cigar_str(oracle.invert(cigar("10=5I3D7X"), F)) == "10=5D3I7X"


def test_invert_reverse(oracle):
    # impg.rs:3223-3239
    assert cigar_str(oracle.invert(cigar("10=5I3D"), R)) == "3I5D10="


def test_invert_empty_and_matches_only(oracle):
    # impg.rs:3242-3264
    assert len(oracle.invert(np.zeros(0, np.uint32), F)) == 0
    assert len(oracle.invert(np.zeros(0, np.uint32), R)) == 0
    assert cigar_str(oracle.invert(cigar("100=50X"), F)) == "100=50X"
    assert cigar_str(oracle.invert(cigar("100=50X"), R)) == "50X100="


def test_identity(oracle):
    # impg.rs:2952-2973: matches / (matches + mismatches + #ins + #del)
    assert oracle.identity(cigar("90=10X")) == pytest.approx(0.9)
    assert oracle.identity(cigar("90=5I5D")) == pytest.approx(90 / 92)
    assert oracle.identity(np.zeros(0, np.uint32)) == 0.0


def test_parse_paf_valid(oracle, tmp_path):
    # impg.rs:3177-3198 and paf.rs:369-392: cg:Z: offset 45, 3 bytes, forward
    p = tmp_path / "a.paf"
    p.write_bytes(b"seq1\t100\t10\t20\t+\tt1\t200\t30\t40\t10\t20\t255\tcg:Z:10M\n")
    idx = oracle.Index.from_paf(str(p))
    recs, offs, runs, lens, names = idx.export()
    assert names == ["seq1", "t1"] and list(lens) == [100, 200]
    r = recs[0]
    assert (r["query_id"], r["query_start"], r["query_end"]) == (0, 10, 20)
    assert (r["target_id"], r["target_start"], r["target_end"]) == (1, 30, 40)
    assert r["strand"] == 0
    assert cigar_str(runs) == "10M"
    assert p.read_bytes()[45:48] == b"10M"


def test_parse_paf_errors(oracle, tmp_path):
    # paf.rs:394-414: too few fields / bad integer / bad strand
    for bad in (b"seq1\t100\t10\n", b"seq1\tx\t10\t20\t+\tt1\t200\t30\t40\t10\t20\t255\n",
                b"seq1\t100\t10\t20\t*\tt1\t200\t30\t40\t10\t20\t255\n"):
        p = tmp_path / "bad.paf"
        p.write_bytes(bad)
        with pytest.raises(ValueError):
            oracle.Index.from_paf(str(p))


def test_parse_paf_reverse_strand_and_first_appearance_ids(oracle, tmp_path):
    p = tmp_path / "b.paf"
    p.write_bytes(b"q\t50\t0\t10\t-\tt\t60\t5\t15\t10\t10\t60\tzz:i:1\tcg:Z:10=\n"
                  b"t\t60\t0\t10\t+\tu\t70\t0\t10\t10\t10\t60\tcg:Z:4=2X4=\n")
    recs, offs, runs, lens, names = oracle.Index.from_paf(str(p)).export()
    assert names == ["q", "t", "u"] and list(lens) == [50, 60, 70]
    assert recs[0]["strand"] == 1 and recs[1]["strand"] == 0
    assert list(offs) == [0, 1, 4]


def test_sorted_ranges_insert(oracle):
    # impg.rs:270-353 with min_distance = 0 (the only value the path uses, :2053)
    r, p = oracle.sorted_ranges_insert([], 1000, 0, (100, 200))
    assert r == [(100, 200)] and p == [(100, 200)]
    r, p = oracle.sorted_ranges_insert([(100, 200)], 1000, 0, (150, 300))
    assert r == [(100, 300)] and p == [(200, 300)]
    r, p = oracle.sorted_ranges_insert([(100, 200), (300, 400)], 1000, 0, (50, 450))
    assert r == [(50, 450)] and p == [(50, 100), (200, 300), (400, 450)]
    r, p = oracle.sorted_ranges_insert([(100, 200)], 1000, 0, (120, 180))
    assert r == [(100, 200)] and p == []
    # reversed input is normalised; touching ranges merge (>=)
    r, p = oracle.sorted_ranges_insert([(100, 200)], 1000, 0, (300, 200))
    assert r == [(100, 300)] and p == [(200, 300)]
    # end beyond the sequence is clamped
    r, p = oracle.sorted_ranges_insert([], 250, 0, (200, 400))
    assert r == [(200, 250)] and p == [(200, 250)]


def test_subset_filter_matches_variants():
    # reference src/subset_filter.rs:185-206, verbatim
    contents = "# comment\nchr1\nchr2\n\nchr1\t\n  chr3  \nHG00097_hap1_hprc_r2_v1.0.1\nHG00098#2#chr5\n"
    yes = ["chr1", "chr1:10-20", "chr3", "HG00097#1#chr7", "HG00097#1", "HG00098#2#chr5"]
    no = ["HG00098#1#chr5"]
    for name in yes:
        assert O.subset_matches(contents, name)[0], name
    for name in no:
        assert not O.subset_matches(contents, name)[0], name
    assert O.subset_matches(contents, "chr1")[1] == 5  # entry_count = distinct trimmed lines


SUBSEQ_KATS = [("HG002#1#chr1:5116130-6116563", ("HG002#1#chr1", 5116130)),   # reference src/main.rs:13330-13346
               ("GRCh38#0#chr1:5477602-6474357", ("GRCh38#0#chr1", 5477602)),
               ("chr1", None), ("chr1:invalid", None)]


def test_parse_subsequence_coordinates():
    for name, want in SUBSEQ_KATS:
        assert O.parse_subsequence_coordinates(name) == want, name
