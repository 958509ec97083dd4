"""refine's flank search and its cache entry points on the device: impgx_refine,
impgx_query_with_cache_batch and impgx_populate_cigar_cache against the oracle's literal restatement of
refine_single_range / query_with_cache / populate_cigar_cache (reference src/commands/refine.rs:144-877,
src/impg.rs:1930-2035). The device answers every sweep of every locus as ONE batch of Impg::query rows;
records (chosen flanks, support counts, supporting entities) must be identical."""
import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    # tiles of 20 kb with ~0.6 kb between them: a locus near a tile border gains support when a flank lets the
    # pieces on both sides of the border merge
    cfg = ix.synth_cfg(9, 2, 200000, 10, 60, 200, 41)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    keep = np.ones(len(recs), bool)
    keep[::7] = False  # some pairs lose tiles: support differs between loci
    sub_offs = np.zeros(int(keep.sum()) + 1, np.uint64)
    nr = np.diff(offs.astype(np.int64))
    np.cumsum(nr[keep], out=sub_offs[1:])
    sub_runs = np.concatenate([runs[int(offs[i]):int(offs[i + 1])] for i in np.nonzero(keep)[0]])
    recs = recs[keep]
    gpu = ix.Impg.from_records(recs, sub_runs, sub_offs, lens, names=names)
    orc = O.Index.build(recs, sub_runs, sub_offs, lens, names=names)
    return cfg, gpu, orc, lens


def loci_of(cfg, n, seed):
    rng = np.random.default_rng(seed)
    out = np.zeros(n, ix.RANGE_DTYPE)
    T = cfg.contig_len // cfg.tiles
    for i in range(n):
        seq = int(rng.integers(0, cfg.genomes * cfg.contigs))
        kind = i % 4
        if kind == 0:    # across a tile border
            s = int(rng.integers(1, cfg.tiles)) * T - int(rng.integers(100, 3000))
            e = s + int(rng.integers(400, 6000))
        elif kind == 1:  # at the start of the sequence (flanks clamp at 0)
            s, e = int(rng.integers(0, 300)), int(rng.integers(800, 9000))
        elif kind == 2:  # at the end of the sequence
            e = cfg.contig_len - int(rng.integers(0, 300))
            s = e - int(rng.integers(800, 9000))
        else:
            s = int(rng.integers(0, cfg.contig_len - 12000))
            e = s + int(rng.integers(300, 12000))
        out[i] = (seq, max(s, 0), min(e, cfg.contig_len))
    return out


def test_populate_cigar_cache_counts(world):
    cfg, gpu, orc, lens = world
    for seq, s, e in loci_of(cfg, 24, 1).tolist():
        assert gpu.populate_cigar_cache(seq, s, e) == orc.populate_cigar_cache(seq, s, e)
    assert gpu.populate_cigar_cache(0, 0, int(lens[0])) == orc.populate_cigar_cache(0, 0, int(lens[0])) > 0


@pytest.mark.parametrize("store_cigar", [False, True])
def test_query_with_cache_batch_matches_query_with_cache(world, store_cigar):
    cfg, gpu, orc, lens = world
    p = ix.make_params(mode=ix.MODE_QUERY, store_cigar=store_cigar)
    for seq, s, e in loci_of(cfg, 8, 2).tolist():
        left = np.array([0, 500, 1000, 2500, 0, 700, 10 ** 9], np.int32)
        right = np.array([0, 0, 0, 0, 3000, 700, 10 ** 9], np.int32)
        res = gpu.query_with_cache_batch(seq, s, e, left, right, p)
        cols = res.columns()
        assert len(cols["row_offsets"]) == len(left) + 1
        for k in range(len(left)):
            cs, ce = max(s - int(left[k]), 0), min(e + int(right[k]), int(lens[seq]))
            want = orc.query_with_cache(seq, cs, ce, (max(s - 2500, 0), min(e + 3000, int(lens[seq]))),
                                        store_cigar=store_cigar).tuples()
            got = res.row_tuples(k, cols)
            if not store_cigar:
                want = [w[:6] + ("",) for w in want]
            assert got == want, (seq, s, e, k)
    # a candidate that is empty after clamping: no results for it, the others unaffected
    res = gpu.query_with_cache_batch(1, 5, 6, np.array([0, 0], np.int32), np.array([0, 4000], np.int32), p)
    r2 = gpu.query_with_cache_batch(1, 10, 5, np.array([0, 3], np.int32), np.array([0, 9], np.int32), p).columns()
    assert r2["row_offsets"].tolist()[:2] == [0, 0] and r2["row_offsets"][2] > 0
    with pytest.raises(ix.ImpgxError):
        gpu.query_with_cache_batch(1, 5, 600, np.array([0], np.int32), np.array([0], np.int32), ix.make_params(mode=ix.MODE_BFS))


CASES = [
    dict(),
    dict(merge_distance=2000, span_bp=300, extension_step=500, max_extension=0.9),
    dict(merge_distance=5000, span_bp=1000, extension_step=1500, max_extension=4000.0),
    dict(merge_distance=-1, span_bp=0, extension_step=800, max_extension=1.0),
    dict(support_level=1, merge_distance=3000, extension_step=600, max_extension=0.8),
    dict(support_level=2, merge_distance=1000, span_bp=200, extension_step=900, max_extension=2500.0),
    dict(transitive=1, max_depth=2, merge_distance=2000, extension_step=2000, max_extension=0.7),
    dict(transitive=2, max_depth=2, merge_distance=2000, extension_step=2500, max_extension=0.7, min_transitive_len=50),
    dict(min_identity=0.985, merge_distance=2500, extension_step=700),
    dict(max_extension=0.0),
]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_refine_records_match_oracle(world, k):
    cfg, gpu, orc, lens = world
    loci = loci_of(cfg, 16, 100 + k)
    p = ix.make_refine_params(**CASES[k])
    got, (n_cand, n_batches) = gpu.refine(loci, p)
    want = orc.refine(loci, p)
    assert got == want
    assert n_batches <= 4 and n_cand >= len(loci)  # baseline + at most three sweeps, each ONE device batch


def test_refine_subset_and_blacklist(world):
    cfg, gpu, orc, lens = world
    n_seqs = len(lens)
    loci = loci_of(cfg, 12, 7)
    rng = np.random.default_rng(3)
    mask = (rng.random(n_seqs) < 0.7).astype(np.uint8)
    bl = {int(q): [(int(a), int(a) + 40000)] for q, a in zip(rng.integers(0, n_seqs, 6), rng.integers(0, 150000, 6))}
    for kw in (dict(subset_mask=mask, merge_distance=2000, extension_step=700),
               dict(blacklist=bl, n_seqs=n_seqs, merge_distance=2000, extension_step=700, support_level=1),
               dict(subset_mask=mask, blacklist=bl, n_seqs=n_seqs, transitive=1, merge_distance=1500, extension_step=1500)):
        p = ix.make_refine_params(**kw)
        assert gpu.refine(loci, p)[0] == orc.refine(loci, p)
    some_support = [r["support_count"] for r in orc.refine(loci, ix.make_refine_params(merge_distance=2000))]
    assert max(some_support) > 0


def test_refine_errors(world):
    cfg, gpu, orc, lens = world
    with pytest.raises(ix.ImpgxError):
        gpu.refine(np.array([(0, 500, 500)], ix.RANGE_DTYPE), ix.make_refine_params())
    with pytest.raises(ix.ImpgxError):
        gpu.refine(np.array([(10 ** 6, 5, 500)], ix.RANGE_DTYPE), ix.make_refine_params())
    with pytest.raises(ix.ImpgxError):
        gpu.refine(np.array([(0, 5, 500)], ix.RANGE_DTYPE), ix.make_refine_params(extension_step=0))
