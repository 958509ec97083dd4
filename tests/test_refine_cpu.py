"""CPU checks of the oracle's restatement of refine's building blocks (reference src/commands/refine.rs,
src/impg.rs:1930-2035): query_with_cache equals Impg::query whatever the cache holds, populate_cigar_cache counts the
alignments under the range, and the flank search behaves as the reference documents it (test infrastructure only)."""
import numpy as np

import _oracle as O
import impg_b200 as ix


def world():
    cfg = ix.synth_cfg(6, 2, 120000, 6, 50, 200, 9)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    return cfg, O.Index.build(recs, runs, offs, lens, names=names), recs, lens


def test_query_with_cache_is_query_for_any_cache():
    cfg, orc, recs, lens = world()
    p = O.make_params(mode=O.MODE_QUERY, store_cigar=True)
    for seq, s, e in [(0, 1000, 9000), (3, 19000, 23000), (7, 100000, 120000)]:
        want = orc.perform_query(seq, s, e, p).tuples()
        for cache in [(s, e), (0, int(lens[seq])), (0, 1), (e, e + 1)]:  # full, wider, disjoint: cache misses fall back
            assert orc.query_with_cache(seq, s, e, cache, store_cigar=True).tuples() == want


def test_populate_cigar_cache_counts_alignments_under_the_range():
    cfg, orc, recs, lens = world()
    for seq, s, e in [(0, 1000, 9000), (3, 19000, 23000), (7, 0, 120000)]:
        # closed stab: entries of the tree of `seq` with first <= e and s <= last; forward entries under the target,
        # reversed entries under the query (self alignments have one)
        fwd = (recs["target_id"] == seq) & (recs["target_start"] <= e) & (recs["target_end"] >= s)
        rev = (recs["query_id"] == seq) & (recs["query_id"] != recs["target_id"]) & (recs["query_start"] <= e) & (recs["query_end"] >= s)
        assert orc.populate_cigar_cache(seq, s, e) == int(fwd.sum() + rev.sum())


def test_refine_prefers_the_smallest_flanks_that_maximise_support():
    cfg, orc, recs, lens = world()
    T = cfg.contig_len // cfg.tiles
    # a locus across a tile border: no alignment spans it, but with -d large enough the pieces left and right merge
    loci = np.array([(2, 2 * T - 1500, 2 * T + 1500)], ix.RANGE_DTYPE)
    none = orc.refine(loci, ix.make_refine_params(merge_distance=0, span_bp=100, extension_step=500))[0]
    some = orc.refine(loci, ix.make_refine_params(merge_distance=3000, span_bp=100, extension_step=500))[0]
    assert none["support_count"] == 0 and (none["applied_left_extension"], none["applied_right_extension"]) == (0, 0)
    assert some["support_count"] > 0 and some["original_support_count"] <= some["support_count"]
    # every reported entity is a sequence other than the target, sorted by (name, start)
    ents = some["support_entities"]
    assert ents and all(q != 2 for q, _, _ in ents) and len({q for q, _, _ in ents}) == len(ents)
    # max_extension 0: only the baseline is evaluated
    base = orc.refine(loci, ix.make_refine_params(merge_distance=3000, max_extension=0.0))[0]
    assert (base["refined_start"], base["refined_end"]) == (int(loci[0]["start"]), int(loci[0]["end"]))
