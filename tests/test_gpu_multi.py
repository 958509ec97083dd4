"""MultiImpg semantics (reference src/multi_impg.rs) on the device: the oracle
keeps one sub-index per alignment file and fans every query out, exactly like
the reference; the product holds all files in one HBM index and reproduces
MultiImpg's hit order (5-key sort), its self-interval handling and its
transitive walk (sorted queue popped at the front / back) with the
IMPGX_MODE_MULTI_* modes. Bit-exact incl. order and CIGARs."""
import os

import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix
from test_gpu_parity import params_pair

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def split_parts(recs, runs, offs, file_of):
    parts = []
    for f in range(int(file_of.max()) + 1):
        keep = np.nonzero(file_of == f)[0]
        po = np.zeros(len(keep) + 1, np.uint64)
        np.cumsum(np.diff(offs.astype(np.int64))[keep], out=po[1:])
        pr = np.concatenate([runs[int(offs[k]):int(offs[k + 1])] for k in keep]) if len(keep) else np.zeros(0, np.uint32)
        parts.append((recs[keep], pr, po))
    return parts


def compare(orc, gpu, ranges, mode, check_cigar=False, bed=False, **kw):
    o, g = params_pair(mode=mode, **kw)
    if bed:
        ores, ooffs = orc.query_batch(ranges, o, bed_merge=True)
        gc = gpu.query_batch_bed(ranges, g).columns()
        keys = ("q_id", "q_first", "q_last")
    else:
        ores, ooffs = orc.query_batch(ranges, o)
        gc = gpu.query_batch(ranges, g).columns()
        keys = ("q_id", "q_first", "q_last", "t_id", "t_first", "t_last")
    oc = ores.columns()
    assert gc["row_offsets"].tolist() == ooffs.tolist()
    for k in keys:
        assert (gc[k] == oc[k]).all(), k
    if check_cigar:
        assert gc["cigar_offsets"].tolist() == oc["cigar_offsets"].tolist()
        assert (gc["cigar_runs"] == oc["cigar_runs"]).all()
    return len(oc["q_id"])


@pytest.fixture(scope="module")
def world():
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    recs = recs.copy()
    recs["query_id"][:8] = recs["target_id"][:8]  # self alignments (incl. hits equal to the self interval's sequence)
    rng = np.random.default_rng(7)
    file_of = rng.integers(0, 4, len(recs)).astype(np.uint32)
    file_of.sort()  # files hold contiguous record blocks, like concatenated PAFs
    orc = O.MultiIndex.build(recs, runs, offs, lens, file_of, 4)
    gpu = ix.MultiImpg.from_record_sets(split_parts(recs, runs, offs, file_of), lens, names=names)
    bed = ix.synth_bed(cfg, 150, seed=9, min_len=200, max_len=12000)
    return cfg, recs, orc, gpu, bed


def test_multi_query(world):
    cfg, recs, orc, gpu, bed = world
    n = compare(orc, gpu.idx, bed, O.MODE_MULTI_QUERY)
    assert n > 150 * 5
    compare(orc, gpu.idx, bed, O.MODE_MULTI_QUERY, store_cigar=True, check_cigar=True)
    compare(orc, gpu.idx, bed, O.MODE_MULTI_QUERY, min_output_length=3000)
    compare(orc, gpu.idx, bed, O.MODE_MULTI_QUERY, min_identity=0.955)
    mask = np.zeros(12, np.uint8)
    mask[[1, 4, 5, 9]] = 1
    compare(orc, gpu.idx, bed, O.MODE_MULTI_QUERY, subset_mask=mask)
    # whole sequences: ranges that coincide with self alignments exercise the duplicate-self rule
    whole = np.array([(s, 0, 60000) for s in range(12)], ix.RANGE_DTYPE)
    compare(orc, gpu.idx, whole, O.MODE_MULTI_QUERY)
    t = recs[0]
    exact = np.array([(t["target_id"], t["target_start"], t["target_end"])], ix.RANGE_DTYPE)
    compare(orc, gpu.idx, exact, O.MODE_MULTI_QUERY)


@pytest.mark.parametrize("mode", [O.MODE_MULTI_BFS, O.MODE_MULTI_DFS])
@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_multi_transitive(world, mode, depth):
    cfg, recs, orc, gpu, bed = world
    compare(orc, gpu.idx, bed[:60], mode, max_depth=depth)


@pytest.mark.parametrize("mode", [O.MODE_MULTI_BFS, O.MODE_MULTI_DFS])
def test_multi_transitive_options(world, mode):
    cfg, recs, orc, gpu, bed = world
    b = bed[:40]
    compare(orc, gpu.idx, b, mode, max_depth=3, store_cigar=True, check_cigar=True)
    compare(orc, gpu.idx, b, mode, max_depth=0, min_transitive_len=0, min_dist=0)
    compare(orc, gpu.idx, b, mode, max_depth=3, min_transitive_len=2000, min_dist=500)
    compare(orc, gpu.idx, b, mode, max_depth=3, min_output_length=2500, min_identity=0.95)
    mask = np.zeros(12, np.uint8)
    mask[[0, 2, 3, 7, 8]] = 1
    compare(orc, gpu.idx, b, mode, max_depth=3, subset_mask=mask)
    compare(orc, gpu.idx, b, mode, max_depth=2, merge_distance=1000, bed=True)
    compare(orc, gpu.idx, b, mode, max_depth=0, merge_distance=0, merge_strands=False, bed=True)
    compare(orc, gpu.idx, b, mode, max_depth=2, merge_distance=-1, merge_strands=False, bed=True)


def test_multi_mirror_and_file_partition_independence(world):
    """The Python mirror maps Impg modes onto MultiImpg's, and the result does not depend on how
    the alignments are spread over files (every partition equals the oracle's 4-file MultiImpg)."""
    cfg, recs, orc, gpu, bed = world
    o, g = params_pair(mode=O.MODE_MULTI_BFS, max_depth=2)
    want = gpu.idx.query_batch(bed[:50], g).columns()
    _, g_plain = params_pair(mode=1, max_depth=2)
    got = gpu.query_batch(bed[:50], g_plain).columns()
    for k in ("row_offsets", "q_id", "q_first", "q_last", "t_id", "t_first", "t_last"):
        assert (got[k] == want[k]).all()
    r = bed[3]
    rows = gpu.query_transitive_bfs(int(r["target_id"]), int(r["start"]), int(r["end"]), max_depth=2)
    a, b = int(want["row_offsets"][3]), int(want["row_offsets"][4])
    assert [x[:3] for x in rows] == list(zip(want["q_id"][a:b].tolist(), want["q_first"][a:b].tolist(),
                                              want["q_last"][a:b].tolist()))


def test_multi_from_paf_files(tmp_path):
    """impgx_index_from_pafs: unified ids by first appearance over the files
    (src/multi_impg.rs:159-176), i.e. the ids of the concatenated PAF."""
    files = [os.path.join(GOLD, f) for f in ("easy_shared_flank.paf", "duplicated_repeat.paf", "short_floor.paf")]
    cat = tmp_path / "cat.paf"
    sizes = []
    with open(cat, "w") as out:
        for f in files:
            lines = [l for l in open(f).read().splitlines() if l.strip()]
            sizes.append(len(lines))
            out.write("\n".join(lines) + "\n")
    whole = O.Index.from_paf(str(cat))
    recs, offs, runs, lens, names = whole.export()
    file_of = np.repeat(np.arange(len(files), dtype=np.uint32), sizes)
    orc = O.MultiIndex.build(recs, runs, offs, lens, file_of, len(files))
    gpu = ix.MultiImpg.from_pafs(files)
    assert [gpu.idx.seq_name(i) for i in range(gpu.idx.n_seqs)] == names
    ranges = np.array([(s, 0, int(lens[s])) for s in range(len(lens))], ix.RANGE_DTYPE)
    for mode in (O.MODE_MULTI_QUERY, O.MODE_MULTI_BFS, O.MODE_MULTI_DFS):
        compare(orc, gpu.idx, ranges, mode, max_depth=0, min_transitive_len=0, store_cigar=True, check_cigar=True)
        compare(orc, gpu.idx, ranges, mode, max_depth=2, merge_distance=0, bed=True)


def test_cli_per_file_index_mode(tmp_path):
    """impgx-query with several alignment files and --index-mode per-file (the reference's MultiImpg
    switch): BED text equals the oracle's MultiImpg + writers."""
    import subprocess
    files = [os.path.join(GOLD, f) for f in ("easy_shared_flank.paf", "duplicated_repeat.paf")]
    cat = tmp_path / "cat.paf"
    sizes = []
    with open(cat, "w") as out:
        for f in files:
            lines = [l for l in open(f).read().splitlines() if l.strip()]
            sizes.append(len(lines))
            out.write("\n".join(lines) + "\n")
    whole = O.Index.from_paf(str(cat))
    recs, offs, runs, lens, names = whole.export()
    orc = O.MultiIndex.build(recs, runs, offs, lens, np.repeat(np.arange(2, dtype=np.uint32), sizes), 2)
    bed = tmp_path / "q.bed"
    rows = [(names[s], 0, int(lens[s])) for s in range(len(names))]
    bed.write_text("".join(f"{a}\t{b}\t{c}\n" for a, b, c in rows))
    cli = os.path.join(os.path.dirname(GOLD), "..", "impg_b200", "impgx-query")
    got = subprocess.run([cli, "-a", *files, "--index-mode", "per-file", "-b", str(bed), "-x", "-m", "2", "-d", "100",
                          "--min-transitive-len", "0", "-o", "bed"], capture_output=True, text=True, check=True).stdout
    p = O.make_params(mode=O.MODE_MULTI_BFS, max_depth=2, min_transitive_len=0, merge_distance=100)
    ranges = np.array([(s, 0, int(lens[s])) for s in range(len(names))], ix.RANGE_DTYPE)
    res, offs2 = orc.query_batch(ranges, p, bed_merge=True)
    c = res.columns()
    want = ""
    for r, (a, b, e) in enumerate(rows):
        for i in range(int(offs2[r]), int(offs2[r + 1])):
            f, l = int(c["q_first"][i]), int(c["q_last"][i])
            want += f"{names[int(c['q_id'][i])]}\t{min(f, l)}\t{max(f, l)}\t{a}:{b}-{e}\t.\t{'+' if f <= l else '-'}\n"
    assert got == want


@pytest.mark.parametrize("kw", [dict(window_size=15000, merge_distance=500),
                                dict(window_size=9000, merge_distance=0, transitive_dfs=True, max_depth=3),
                                dict(window_size=20000, merge_distance=-1, selection_mode="total")])
def test_partition_over_a_multi_impg(world, kw):
    """`partition` with --index-mode per-file: the windows are answered by MultiImpg's transitive walk
    (reference src/multi_impg.rs:687-755 behind src/commands/partition.rs:359-391)."""
    cfg, recs, orc, gpu, bed = world
    names = [gpu.idx.seq_name(i) for i in range(gpu.idx.n_seqs)]
    want = orc.partition(O.make_partition_params(multi_impg=True, **kw), names)
    got = gpu.idx.partition(ix.make_partition_params(multi_impg=True, **kw))
    assert got.rows() == [(p, s, min(a, b), max(a, b)) for p, s, a, b in want["rows"]]
    assert (got.n_partitions, got.partitioned_bp, got.n_windows) == (want["n_partitions"], want["partitioned_bp"],
                                                                     len(want["windows"]))
    assert got.partitioned_bp == got.total_bp
