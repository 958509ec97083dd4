"""`.impg` index files on the CPU: libimpgx's writer and reader against the independent
Python restatement of the format in tests/_impg_format.py (byte for byte), on the reference's fixture
PAFs and on synthetic files that exercise every varint width. No device involved."""
import glob
import os
import random

import numpy as np
import pytest

import _impg_format as F
import _oracle as O
import impg_b200 as ix

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAFS = sorted(glob.glob(os.path.join(GOLD, "*.paf")))


def rec_tuples(recs):
    return [(int(r["query_id"]), int(r["target_id"]), int(r["query_start"]), int(r["query_end"]), int(r["target_start"]),
             int(r["target_end"]), int(r["strand"])) for r in recs]


@pytest.mark.parametrize("paf", PAFS, ids=[os.path.basename(p) for p in PAFS])
@pytest.mark.parametrize("bidirectional", [True, False])
def test_writer_matches_the_python_restatement_byte_for_byte(paf, bidirectional, tmp_path):
    out = str(tmp_path / "x.impg")
    ix.impg_write([paf], out, bidirectional=bidirectional)
    names, lens, recs = F.parse_paf_like_reference([paf])
    want = F.encode(names, lens, F.entries_by_target(recs, bidirectional))
    assert open(out, "rb").read() == want
    # and the reader returns the alignments of the PAF, in PAF order, with offsets that point at the CIGAR text
    f = ix.ImpgFile(out)
    assert (f.version, f.bidirectional) == (2, bidirectional and any(r[0] != r[1] for r in recs))
    assert f.names == names and f.lens.tolist() == lens
    got, fi, off, ln = f.records()
    assert rec_tuples(got) == [r[:7] for r in recs]
    assert fi.tolist() == [0] * len(recs) and off.tolist() == [r[8] for r in recs] and ln.tolist() == [r[9] for r in recs]
    assert f.n_entries == sum(len(v) for v in F.entries_by_target(recs, bidirectional).values())
    text = open(paf, "rb").read()
    orc = O.Index.from_paf(paf)
    o_recs, o_offs, o_runs, o_lens, o_names = orc.export()
    assert o_names == names and rec_tuples(o_recs) == rec_tuples(got)
    for k in range(len(recs)):
        cg = text[int(off[k]):int(off[k]) + int(ln[k])]
        assert text[int(off[k]) - 5:int(off[k])] == b"cg:Z:"
        assert O.parse_cigar(cg).tolist() == o_runs[int(o_offs[k]):int(o_offs[k + 1])].tolist()


def test_multi_file_index_and_hash_ordered_maps(tmp_path):
    # two alignment files behind one index; the Python encoder writes maps and trees in a scrambled order, as the
    # reference's FxHashMap iteration would — any order must read back the same
    pafs = PAFS[:3]
    names, lens, recs = F.parse_paf_like_reference(pafs)
    trees = F.entries_by_target(recs)
    rnd = random.Random(5)
    mo = list(range(len(names)))
    to = list(trees)
    rnd.shuffle(mo)
    rnd.shuffle(to)
    p = tmp_path / "scrambled.impg"
    p.write_bytes(F.encode(names, lens, trees, map_order=mo, tree_order=to))
    f = ix.ImpgFile(str(p))
    got, fi, off, ln = f.records()
    assert f.names == names and f.lens.tolist() == lens and f.bidirectional
    assert rec_tuples(got) == [r[:7] for r in recs] and fi.tolist() == [r[7] for r in recs]
    out = str(tmp_path / "w.impg")
    ix.impg_write(pafs, out)
    assert F.decode(open(out, "rb").read()) == (names, lens, trees)
    assert open(out, "rb").read() == F.encode(names, lens, trees)


def test_every_varint_width_and_legacy_magic(tmp_path):
    names = ["s%d" % i for i in range(300)] + ["x" * 300]  # > 250 sequences, a name longer than 250 bytes
    lens = [10, 250, 251, 65535, 65536, 2**31 - 1] + [1000] * 295
    big = (1 << 40) + 12345  # an offset beyond 2^32
    recs = [(0, 5, 0, 10, 100, 2**31 - 1, 1, 0, big, 70000),
            (299, 300, 250, 251, 65535, 65536, 0, 0, 7, 250),
            (3, 3, 5, 50, 5, 50, 0, 1, 251, 251)]
    trees = F.entries_by_target(recs)
    for magic, ver in ((b"IMPGIDX2", 2), (b"IMPGIDX1", 1)):
        p = tmp_path / ("v%d.impg" % ver)
        data = F.encode(names, lens, trees, magic=magic)
        p.write_bytes(data)
        assert F.decode(data) == (names, lens, trees)
        f = ix.ImpgFile(str(p))
        assert f.version == ver and f.names == names and f.lens.tolist() == lens
        got, fi, off, ln = f.records()
        order = sorted(range(3), key=lambda k: (recs[k][7], recs[k][8]))
        assert rec_tuples(got) == [recs[k][:7] for k in order]
        assert off.tolist() == [recs[k][8] for k in order] and ln.tolist() == [recs[k][9] for k in order]
        assert f.n_entries == 5 and f.n_records == 3  # the self alignment has no reversed copy


def test_crlf_paf_offsets_follow_the_reference_rule(tmp_path):
    # BufRead::lines strips "\r\n" and the reference adds len + 1 per line (src/paf.rs:182-191), so on a CRLF
    # file its recorded offsets drift by one byte per line; an index written here must carry the same numbers
    crlf = tmp_path / "crlf.paf"
    crlf.write_bytes(open(PAFS[0], "rb").read().replace(b"\n", b"\r\n"))
    out = str(tmp_path / "c.impg")
    ix.impg_write([str(crlf)], out)
    names, lens, recs = F.parse_paf_like_reference([str(crlf)])
    assert open(out, "rb").read() == F.encode(names, lens, F.entries_by_target(recs))
    plain = F.parse_paf_like_reference([PAFS[0]])[2]
    assert [r[8] for r in recs] == [r[8] for r in plain] and [r[9] for r in recs] == [r[9] for r in plain]


def test_reader_rejects_damaged_files(tmp_path):
    out = str(tmp_path / "x.impg")
    ix.impg_write([PAFS[0]], out)
    data = open(out, "rb").read()
    for bad in (b"NOTANIDX" + data[8:], data[:40], data[:8] + (2**40).to_bytes(8, "little") + data[16:], b""):
        p = tmp_path / "bad.impg"
        p.write_bytes(bad)
        with pytest.raises(ix.ImpgxError) as e:
            ix.ImpgFile(str(p))
        assert e.value.code == ix.E_PARSE
    with pytest.raises(ix.ImpgxError) as e:
        ix.ImpgFile(str(tmp_path / "missing.impg"))
    assert e.value.code == ix.E_IO
    import gzip
    gz = tmp_path / "a.paf.gz"
    gz.write_bytes(gzip.compress(open(PAFS[0], "rb").read()))
    with pytest.raises(ix.ImpgxError, match="regular gzip, not BGZF") as e:  # the reference's error (src/paf.rs:310-318)
        ix.impg_write([str(gz)], out)
    assert e.value.code == ix.E_PARSE


@pytest.mark.parametrize("paf", PAFS[:3], ids=[os.path.basename(p) for p in PAFS[:3]])
def test_bgzf_paf_behind_an_index_uses_virtual_positions(paf, tmp_path):
    """A .gz / .bgz alignment file is BGZF and its CIGARs are addressed by virtual positions
    (reference src/paf.rs:68-114, :199-302): the writer must produce what the Python restatement computes from its
    own block table, for blocks small enough that CIGARs start right at and straddle block borders."""
    text = open(paf, "rb").read()
    for block in (700, 64, 65280):
        bz, table = F.bgzf_compress(text, block)
        gz = str(tmp_path / f"x{block}.paf.gz")
        open(gz, "wb").write(bz)
        out = str(tmp_path / f"x{block}.impg")
        ix.impg_write([gz], out)
        names, lens, recs = F.parse_paf_like_reference([paf])
        vrecs = [r[:8] + (F.virtual_position(table, r[8]), r[9]) for r in recs]
        assert open(out, "rb").read() == F.encode(names, lens, F.entries_by_target(vrecs))
        got, fi, off, ln = ix.ImpgFile(out).records()
        assert off.tolist() == [r[8] for r in vrecs] and ln.tolist() == [r[9] for r in recs]
    # a plain gzip file under a .gz name is refused like the reference does (src/paf.rs:306-318)
    import gzip
    plain = str(tmp_path / "plain.paf.gz")
    with gzip.open(plain, "wb") as f:
        f.write(text)
    with pytest.raises(ix.ImpgxError, match="not BGZF"):
        ix.impg_write([plain], str(tmp_path / "p.impg"))
