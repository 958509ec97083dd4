"""`impg partition -o bed` on the device: impgx_partition — one masked transitive
query + BED merge per window on the GPU, window / mask bookkeeping on the host — against the
oracle's restatement of partition_alignments (reference src/commands/partition.rs:158-712).
Bit-exact rows of partitions.bed, the same windows queried, the same totals."""
import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix
from test_partition_cpu import CASES, norm, patchy_world

pytestmark = pytest.mark.gpu


def both(recs, runs, offs, lens, names):
    return ix.Impg.from_records(recs, runs, offs, lens, names=names), O.Index.build(recs, runs, offs, lens, names=names)


def check(g, o, kw):
    want = o.partition(O.make_partition_params(**kw))
    got = g.partition(ix.make_partition_params(**kw))
    assert got.rows() == norm(want["rows"])
    assert (got.n_partitions, got.partitioned_bp, got.total_bp, got.n_windows) == (
        want["n_partitions"], want["partitioned_bp"], want["total_bp"], len(want["windows"]))
    assert got.format_bed(g) == want["bed"]
    return got, want


@pytest.mark.parametrize("k", range(len(CASES)))
def test_partition_matches_oracle_uniform(k):
    cfg = ix.synth_cfg(5, 2, 60000, 6, 40, 300, 11 + k)
    g, o = both(*ix.synth_generate(cfg))
    check(g, o, CASES[k])


@pytest.mark.parametrize("k", range(len(CASES)))
def test_partition_matches_oracle_patchy(k):
    kw = dict(CASES[k])
    kw["window_size"] = max(3000, kw["window_size"] // 3)
    world, o = patchy_world(seed=101 + k, keep=0.2 + 0.05 * (k % 5))
    g = ix.Impg.from_records(*world[:4], names=world[4])
    got, want = check(g, o, kw)
    # --separate-files: partition<N>.bed holds the rows of partition N without the number column
    p = int(got.partition_num[len(got.partition_num) // 2])
    lines = [l.rsplit("\t", 1)[0] for l in want["bed"].splitlines() if l.rsplit("\t", 1)[1] == str(p)]
    assert got.format_bed(g, p) == "".join(l + "\n" for l in lines)


def test_partition_identity_filter_and_fixture_paf(tmp_path):
    world, o = patchy_world(seed=77, keep=0.5, genomes=5)
    g = ix.Impg.from_records(*world[:4], names=world[4])
    check(g, o, dict(window_size=8000, merge_distance=300, min_identity=0.975))
    # the reference's scenario PAF (tests/test_transitive_integrity.rs:592-646)
    paf = tmp_path / "t.paf"
    paf.write_text("A\t10000\t0\t1000\t+\tB\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000=\n"
                   "A\t10000\t5000\t6000\t+\tC\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000=\n")
    g2, o2 = ix.Impg.from_paf(str(paf)), O.Index.from_paf(str(paf))
    got, _ = check(g2, o2, dict(window_size=2000, merge_distance=100000))
    assert len(got.format_bed(g2).splitlines()) >= 2


def test_partition_medium_world_tiles_the_pangenome():
    # size-independent property at a size the oracle does not need to replay: every base of every
    # sequence lands in exactly one partition interval
    cfg = ix.synth_cfg(12, 3, 400000, 16, 60, 200, 5)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    g = ix.Impg.from_records(recs, runs, offs, lens, names=names)
    got = g.partition(ix.make_partition_params(window_size=50000, merge_distance=10000))
    assert got.partitioned_bp == got.total_bp == int(lens.sum())
    order = np.lexsort((got.start, got.seq_id))
    s, a, b = got.seq_id[order], got.start[order], got.end[order]
    first = np.r_[True, s[1:] != s[:-1]]
    last = np.r_[s[1:] != s[:-1], True]
    assert (a[first] == 0).all() and (b[last] == lens[s[last]].astype(np.int64)).all()
    assert (a[~first] == b[:-1][~first[1:]]).all()
    assert got.n_windows >= got.n_partitions > 0


def test_cli_partition_writes_the_reference_files(tmp_path):
    """`impgx-query partition` end to end on the reference's scenario PAF and a fixture PAF:
    partitions.bed / partition<N>.bed equal to the oracle's text."""
    import os
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "impg_b200", "impgx-query")
    paf = tmp_path / "t.paf"
    paf.write_text("A\t10000\t0\t1000\t+\tB\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000=\n"
                   "A\t10000\t5000\t6000\t+\tC\t5000\t0\t1000\t1000\t1000\t60\tcg:Z:1000=\n")
    o = O.Index.from_paf(str(paf))
    want = o.partition(O.make_partition_params(window_size=2000, merge_distance=100000))
    out = tmp_path / "single"
    r = subprocess.run([cli, "partition", "-a", str(paf), "-w", "2000", "-d", "100k", "--output-folder", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (out / "partitions.bed").read_text() == want["bed"]
    assert "Partitioned into" in r.stderr
    # --separate-files, no rehoming, a starting-sequences file
    start = tmp_path / "start.txt"
    start.write_text("# comment\n\nC\tignored\nnot_there\n")
    want = o.partition(O.make_partition_params(window_size=3000, merge_distance=0, rehome_singletons=False,
                                               starting_seqs=[o.seq_id("C")], selection_mode="total"))
    out = tmp_path / "sep"
    r = subprocess.run([cli, "partition", "-a", str(paf), "-w", "3000", "-d", "0", "--separate-files",
                        "--no-rehome-singletons", "--starting-sequences-file", str(start), "--selection-mode", "total",
                        "--output-folder", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    by_part = {}
    for l in want["bed"].splitlines():
        body, p = l.rsplit("\t", 1)
        by_part.setdefault(p, []).append(body + "\n")
    assert sorted(f.name for f in out.iterdir()) == sorted(f"partition{p}.bed" for p in by_part)
    for p, lines in by_part.items():
        assert (out / f"partition{p}.bed").read_text() == "".join(lines)
    r = subprocess.run([cli, "partition", "-a", str(paf), "-w", "3000"], capture_output=True, text=True)
    assert r.returncode != 0 and "merge-distance is required" in r.stderr
