"""An index opened from a `.impg` file (reference src/impg.rs:1777-1850) answers exactly like
the index built from the PAF itself — same sequence ids, same result order, same CIGARs — and like the oracle."""
import glob
import os

import numpy as np
import pytest

import _impg_format as F
import _oracle as O
import impg_b200 as ix

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAFS = sorted(glob.glob(os.path.join(GOLD, "*.paf")))


def all_rows(idx):
    return np.array([(s, 0, idx.seq_len(s)) for s in range(idx.n_seqs)], ix.RANGE_DTYPE)


@pytest.mark.parametrize("paf", PAFS, ids=[os.path.basename(p) for p in PAFS])
def test_index_from_impg_equals_index_from_paf(paf, tmp_path):
    out = str(tmp_path / "x.impg")
    ix.impg_write([paf], out)
    a, b = ix.Impg.from_paf(paf), ix.Impg.from_impg(out, [paf])
    assert b.n_seqs == a.n_seqs and b.n_entries == a.n_entries
    assert [b.seq_name(i) for i in range(b.n_seqs)] == [a.seq_name(i) for i in range(a.n_seqs)]
    orc = O.Index.from_paf(paf)
    for mode, depth in ((ix.MODE_QUERY, 1), (ix.MODE_BFS, 0)):
        p = ix.make_params(mode=mode, max_depth=depth, min_transitive_len=0, store_cigar=True)
        ra, rb = a.query_batch(all_rows(a), p), b.query_batch(all_rows(b), p)
        ca, cb = ra.columns(), rb.columns()
        assert all((ca[k] == cb[k]).all() for k in ca)
        op = O.make_params(mode=mode, max_depth=depth, min_transitive_len=0, store_cigar=True)
        for r in range(min(b.n_seqs, 6)):
            want = orc.perform_query(r, 0, orc.seq_len(r), op).tuples()
            assert rb.row_tuples(r, cb) == want


def test_index_from_a_scrambled_multi_file_impg(tmp_path):
    # the file as stock impg could have written it: hash-ordered maps and trees, two alignment files
    pafs = PAFS[:2]
    names, lens, recs = F.parse_paf_like_reference(pafs)
    trees = F.entries_by_target(recs)
    p = tmp_path / "m.impg"
    p.write_bytes(F.encode(names, lens, trees, map_order=list(reversed(range(len(names)))), tree_order=sorted(trees, reverse=True)))
    b = ix.Impg.from_impg(str(p), pafs)
    a = ix.MultiImpg.from_pafs(pafs).idx
    assert [b.seq_name(i) for i in range(b.n_seqs)] == names
    prm = ix.make_params(mode=ix.MODE_BFS, max_depth=2, min_transitive_len=0, merge_distance=100)
    ca, cb = a.query_batch_bed(all_rows(a), prm).columns(), b.query_batch_bed(all_rows(b), prm).columns()
    assert all((ca[k] == cb[k]).all() for k in ca)
    with pytest.raises(ix.ImpgxError) as e:  # the second alignment file is missing
        ix.Impg.from_impg(str(p), pafs[:1])
    assert e.value.code == ix.E_INVALID
    short = tmp_path / "short.paf"  # not the file the index was built from: the offsets fall off its end
    short.write_bytes(open(pafs[1], "rb").read()[:40])
    with pytest.raises(ix.ImpgxError) as e:
        ix.Impg.from_impg(str(p), [pafs[0], str(short)])
    assert e.value.code == ix.E_PARSE


def test_cli_index_then_query_with_index(tmp_path):
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "impg_b200", "impgx-query")
    paf = os.path.join(GOLD, "short_floor.paf")
    impg = str(tmp_path / "sf.impg")
    r = subprocess.run([cli, "index", "-a", paf, "-i", impg], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    args = ["-r", "C4FIXTURE#0#short_floor:0-250:0-250", "-d", "0", "--min-transitive-len", "0", "-x", "-o", "bed"]
    a = subprocess.run([cli, "-a", paf] + args, capture_output=True, text=True)
    b = subprocess.run([cli, "-a", paf, "-i", impg] + args, capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    assert a.stdout == b.stdout and a.stdout.count("\n") >= 2
    p1 = subprocess.run([cli, "partition", "-a", paf, "-i", impg, "-w", "200", "-d", "50", "--output-folder", str(tmp_path / "p1")],
                        capture_output=True, text=True)
    p2 = subprocess.run([cli, "partition", "-a", paf, "-w", "200", "-d", "50", "--output-folder", str(tmp_path / "p2")],
                        capture_output=True, text=True)
    assert p1.returncode == 0 and p2.returncode == 0, (p1.stderr, p2.stderr)
    assert (tmp_path / "p1" / "partitions.bed").read_text() == (tmp_path / "p2" / "partitions.bed").read_text()


def test_cli_subset_sequence_list_uses_the_reference_matching_rules(tmp_path):
    """--subset-sequence-list through impgx_subset_mask: exact names with coordinates select single sequences,
    the sample+haplotype key `C4FIXTURE#0` (reference src/subset_filter.rs:147-178) selects all of them."""
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "impg_b200", "impgx-query")
    paf = os.path.join(GOLD, "short_floor.paf")
    args = [cli, "-a", paf, "-r", "C4FIXTURE#0#short_floor:0-250:0-250", "-d", "0", "-o", "bed"]
    full = subprocess.run(args, capture_output=True, text=True)
    assert full.returncode == 0, full.stderr
    lst = tmp_path / "subset.txt"
    # an exact name also registers its sample + haplotype key (parse_subset_filter, :117-145), and every sequence
    # of this fixture is C4FIXTURE haplotype 0 — so one listed sequence keeps them all, as in the reference
    lst.write_text("# one sequence by exact name\n  C4FIXTURE#0#short_floor:60-360\t\n")
    one = subprocess.run(args + ["--subset-sequence-list", str(lst)], capture_output=True, text=True)
    assert one.returncode == 0 and one.stdout == full.stdout, one.stderr
    lst.write_text("C4FIXTURE#1\n")  # the other haplotype: nothing but the query sequence itself is kept
    hap1 = subprocess.run(args + ["--subset-sequence-list", str(lst)], capture_output=True, text=True)
    assert hap1.returncode == 0 and {l.split("\t")[0] for l in hap1.stdout.splitlines()} == {"C4FIXTURE#0#short_floor:0-250"}
    lst.write_text("C4FIXTURE#0\n")  # sample + haplotype: every sequence of the fixture
    allseq = subprocess.run(args + ["--subset-sequence-list", str(lst)], capture_output=True, text=True)
    assert allseq.returncode == 0 and allseq.stdout == full.stdout
    lst.write_text("OTHER#1\n")
    none = subprocess.run(args + ["--subset-sequence-list", str(lst)], capture_output=True, text=True)
    assert none.returncode == 0 and {l.split("\t")[0] for l in none.stdout.splitlines()} == {"C4FIXTURE#0#short_floor:0-250"}
    lst.write_text("# nothing\n\n")
    bad = subprocess.run(args + ["--subset-sequence-list", str(lst)], capture_output=True, text=True)
    assert bad.returncode != 0 and "did not contain any sequence names" in bad.stderr


def test_original_sequence_coordinates_in_the_writers(tmp_path):
    """--original-sequence-coordinates (reference src/main.rs:4642-4678, :11869-11881, :11913-11934): names of
    the form base:start-end are reported in the coordinates of `base`, by BED (per row and per batch) and BEDPE."""
    paf = os.path.join(GOLD, "short_floor.paf")
    gpu, orc = ix.Impg.from_paf(paf), O.Index.from_paf(paf)
    rows = np.array([(s, 0, orc.seq_len(s)) for s in range(orc.n_seqs)], ix.RANGE_DTYPE)
    names = [f"r{k}" for k in range(len(rows))]
    for on in (True, False):
        gpu.set_original_coordinates(on)
        orc.set_original_coordinates(on)
        gp = ix.make_params(mode=ix.MODE_BFS, max_depth=2, min_transitive_len=0, merge_distance=10)
        op = O.make_params(mode=O.MODE_BFS, max_depth=2, min_transitive_len=0, merge_distance=10)
        res = gpu.query_batch_bed(rows, gp)
        want = "".join(orc.format(orc.perform_query(int(r["target_id"]), int(r["start"]), int(r["end"]), op), "bed", names[k], 10)
                       for k, r in enumerate(rows))
        assert gpu.format_bed_batch(res, names) == want
        assert "".join(gpu.format_bed(res, k, names[k]) for k in range(len(rows))) == want
        assert (":0-250" not in want) == on and ("C4FIXTURE#0#short_floor\t" in want) == on
        gc = ix.make_params(mode=ix.MODE_QUERY, store_cigar=True)
        oc = O.make_params(mode=O.MODE_QUERY, store_cigar=True)
        raw = gpu.query_batch(rows, gc)
        for k, r in enumerate(rows):
            ores = orc.perform_query(int(r["target_id"]), int(r["start"]), int(r["end"]), oc)
            ores.drop_first()
            assert ix.format_bedpe(gpu, raw, k, names[k], 0) == orc.format(ores, "bedpe", names[k], 0)
        if on:
            with pytest.raises(ix.ImpgxError):
                ix.format_paf(gpu, raw, 0, names[0], 0)
        else:
            assert ix.format_paf(gpu, raw, 0, names[0], 0)


@pytest.mark.parametrize("block", [64, 900])
def test_index_from_impg_over_a_bgzf_paf(block, tmp_path):
    """The alignment file behind the index is BGZF: every CIGAR is fetched through its virtual position
    (reference src/paf.rs:68-114). Same answers as the index over the plain PAF."""
    paf = PAFS[1]
    bz, _ = F.bgzf_compress(open(paf, "rb").read(), block)
    gz = str(tmp_path / "x.paf.bgz")
    open(gz, "wb").write(bz)
    out = str(tmp_path / "x.impg")
    ix.impg_write([gz], out)
    a, b = ix.Impg.from_paf(paf), ix.Impg.from_impg(out, [gz])
    assert b.n_seqs == a.n_seqs and b.n_entries == a.n_entries
    p = ix.make_params(mode=ix.MODE_BFS, max_depth=0, min_transitive_len=0, store_cigar=True)
    ca, cb = a.query_batch(all_rows(a), p).columns(), b.query_batch(all_rows(b), p).columns()
    assert all((ca[k] == cb[k]).all() for k in ca)
    # the index of the PLAIN file does not open the compressed one (its offsets are not virtual positions)
    plain = str(tmp_path / "plain.impg")
    ix.impg_write([paf], plain)
    with pytest.raises(ix.ImpgxError):
        ix.Impg.from_impg(plain, [gz])
