"""GPU parity of the single-launch path for calls of a few rows (csrc/small_bfs.cuh): what `impg partition`
and `impg refine` issue per window / flank (src/commands/partition.rs:359-391, src/commands/refine.rs:493-531).
Every case is compared with the oracle; `kernel_launches` proves which path answered (the batched pipeline
needs dozens of launches for the same call)."""
import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix
from test_gpu_parity import build_both, compare_bed, compare_raw, params_pair, random_mask

pytestmark = pytest.mark.gpu

SMALL_MAX_LAUNCHES = 8  # walk + six bucket-merge launches + finish


@pytest.fixture(scope="module")
def world():
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    bed = ix.synth_bed(cfg, 96, seed=11, min_len=200, max_len=12000)
    return cfg, orc, gpu, bed


def took_small_path(gpu, bed_out):
    st = gpu.stats()
    return st["kernel_launches"] <= (SMALL_MAX_LAUNCHES if bed_out else 2)


@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_one_row_calls_raw(world, depth):
    cfg, orc, gpu, bed = world
    for k in range(24):
        compare_raw(orc, gpu, bed[k:k + 1], *params_pair(mode=1, max_depth=depth))
        assert took_small_path(gpu, False)


@pytest.mark.parametrize("rows", [2, 5, 16, 48])
def test_few_row_calls_raw_and_bed(world, rows):
    cfg, orc, gpu, bed = world
    for k in range(0, 96 - rows + 1, rows):
        b = bed[k:k + rows]
        compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=2))
        assert took_small_path(gpu, False)
        compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=2, merge_distance=1000))
        assert took_small_path(gpu, True)


def test_more_rows_than_ctas_take_the_batched_path(world):
    cfg, orc, gpu, bed = world
    compare_raw(orc, gpu, bed[:64], *params_pair(mode=1, max_depth=2))
    assert took_small_path(gpu, False)
    compare_raw(orc, gpu, bed[:65], *params_pair(mode=1, max_depth=2))
    assert not took_small_path(gpu, False)


def test_query_mode_and_filters(world):
    cfg, orc, gpu, bed = world
    mask = np.zeros(12, np.uint8)
    mask[[1, 4, 5, 9]] = 1
    for k in range(0, 48, 3):
        b = bed[k:k + 3]
        compare_raw(orc, gpu, b, *params_pair(mode=0))
        assert took_small_path(gpu, False)
        compare_raw(orc, gpu, b, *params_pair(mode=0, min_output_length=3000))
        compare_raw(orc, gpu, b, *params_pair(mode=0, subset_mask=mask))
        compare_bed(orc, gpu, b, *params_pair(mode=0, merge_distance=0))
        assert took_small_path(gpu, True)
        compare_bed(orc, gpu, b, *params_pair(mode=0, merge_distance=500, min_output_length=2000))
        compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, min_transitive_len=0, min_dist=0))
        compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, min_transitive_len=2000, min_dist=500))
        compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=2, min_output_length=2500))
        compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, subset_mask=mask))
        assert took_small_path(gpu, False)


@pytest.mark.parametrize("d,merge_strands", [(0, True), (1000, True), (1000, False), (-1, True), (100000, True)])
def test_bed_merge_variants(world, d, merge_strands):
    cfg, orc, gpu, bed = world
    for k in range(0, 40, 4):
        compare_bed(orc, gpu, bed[k:k + 4], *params_pair(mode=1, max_depth=3, merge_distance=d, merge_strands=merge_strands))
        assert took_small_path(gpu, True)


def test_unsorted_bed_output_stays_on_the_batched_path(world):
    cfg, orc, gpu, bed = world
    compare_bed(orc, gpu, bed[:4], *params_pair(mode=1, max_depth=2, merge_distance=-1, merge_strands=False))
    assert not took_small_path(gpu, True)


def test_masked_windows(world):
    """What partition passes: one window, the regions assigned so far as masked_regions."""
    cfg, orc, gpu, bed = world
    rng = np.random.default_rng(77)
    for density, depth in ((0.5, 2), (0.9, 3), (0.0, 2), (0.97, 0)):
        mask = random_mask(rng, 12, 60000, density)
        for k in range(0, 32, 2):
            b = bed[k:k + 2]
            compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=depth, masked_regions=mask))
            assert took_small_path(gpu, False)
            compare_bed(orc, gpu, b, *params_pair(mode=1, max_depth=depth, masked_regions=mask, merge_distance=1000))
            assert took_small_path(gpu, True)
            compare_raw(orc, gpu, b, *params_pair(mode=1, max_depth=3, masked_regions=mask, min_output_length=2500,
                                                   min_transitive_len=0, min_dist=0))
    full = ix.mask_csr({s: [(0, 60000)] for s in range(12)}, 12)
    res = gpu.query_batch(bed[:3], params_pair(mode=1, masked_regions=full)[1])
    assert res.n_results == 0 and took_small_path(gpu, False)
    bad = (np.array([0, 2] + [2] * 11, np.uint64), np.array([10, 50, 40, 90], np.int32))
    with pytest.raises(ix.ImpgxError) as e:
        gpu.query_batch(bed[:3], params_pair(mode=1, masked_regions=bad)[1])
    assert e.value.code == ix.E_INVALID


def test_edges_and_invalid_rows():
    cfg = ix.synth_cfg(3, 1, 30000, 3, 20, 500, 5)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    o, g = params_pair(mode=1, max_depth=2)
    gap = np.array([(0, 9995, 10001)], ix.RANGE_DTYPE)
    compare_raw(orc, gpu, gap, o, g)
    compare_bed(orc, gpu, gap, o, g)
    t0 = recs[0]
    edges = np.array([(t0["target_id"], max(t0["target_start"] - 50, 0), t0["target_start"]),
                      (t0["target_id"], t0["target_end"], t0["target_end"] + 50),
                      (t0["target_id"], t0["target_start"], t0["target_start"] + 1),
                      (t0["target_id"], t0["target_end"] - 1, t0["target_end"]),
                      (t0["target_id"], 0, 30000)], ix.RANGE_DTYPE)
    edges = edges[edges["start"] < edges["end"]]
    compare_raw(orc, gpu, edges, *params_pair(mode=0))
    compare_raw(orc, gpu, edges, *params_pair(mode=1, max_depth=0, min_transitive_len=0))
    compare_bed(orc, gpu, edges, *params_pair(mode=1, max_depth=0, min_transitive_len=0, merge_distance=0))
    assert took_small_path(gpu, True)
    for bad in ([(99, 0, 10)], [(0, 10, 10)], [(0, 20, 10)], [(0, 0, 30001)], [(0, -5, 10)],
                [(0, 5, 100), (0, 20, 10)]):
        with pytest.raises(ix.ImpgxError) as e:
            gpu.query_batch(np.array(bad, ix.RANGE_DTYPE), g)
        assert e.value.code == ix.E_INVALID
        assert "range" in str(e.value)


def test_self_alignments_and_unidirectional_index():
    cfg = ix.synth_cfg(4, 1, 40000, 5, 25, 500, 21)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    recs = recs.copy()
    recs["query_id"][:6] = recs["target_id"][:6]
    for bidir in (True, False):
        orc, gpu = build_both(recs, runs, offs, lens, names, bidirectional=bidir)
        bed = ix.synth_bed(cfg, 40, seed=4, min_len=300, max_len=9000)
        for k in range(0, 40, 5):
            compare_raw(orc, gpu, bed[k:k + 5], *params_pair(mode=0))
            compare_raw(orc, gpu, bed[k:k + 5], *params_pair(mode=1, max_depth=3))
            compare_bed(orc, gpu, bed[k:k + 5], *params_pair(mode=1, max_depth=3, merge_distance=100))
            assert took_small_path(gpu, True)


def test_rows_beyond_the_capacities_fall_back_exactly():
    """Two sequences, 6000 alignments per pair: a whole-contig row lifts ~12,000 hits in one hop (more than the
    8192 the single-launch path holds), a 700 kbp row ~5,000 boxes on one query sequence (a bucket beyond the
    largest on-chip merge class). Such calls are answered by the batched path, bit-exact, and the index backs
    off from retrying on every call."""
    cfg = ix.synth_cfg(2, 1, 1600000, 6000, 8, 200, 23)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    big = np.array([(0, 0, 1600000), (1, 5, 1599000)], ix.RANGE_DTYPE)
    mid = np.array([(0, 1000, 700000)], ix.RANGE_DTYPE)
    tiny = np.array([(0, 5000, 5400), (1, 100000, 100900)], ix.RANGE_DTYPE)
    compare_bed(orc, gpu, big, *params_pair(mode=0, merge_distance=10))
    assert not took_small_path(gpu, True) and gpu.stats()["liftovers"] > 20000
    compare_raw(orc, gpu, big, *params_pair(mode=1, max_depth=2))
    assert not took_small_path(gpu, False)
    for _ in range(3):  # past the back-off: the attempt is made again and declined again
        compare_bed(orc, gpu, mid, *params_pair(mode=0, merge_distance=10))
        assert not took_small_path(gpu, True)
    assert 4096 < gpu.stats()["liftovers"] <= 8192
    # the same rows in reference order fit (no bucket involved)
    for _ in range(12):
        compare_raw(orc, gpu, mid, *params_pair(mode=0))
    assert took_small_path(gpu, False)
    # the small rows of the same index fit again once the back-off has run out
    fits = 0
    for _ in range(40):
        compare_bed(orc, gpu, tiny, *params_pair(mode=1, max_depth=2, merge_distance=1000))
        fits += took_small_path(gpu, True)
    assert fits >= 20
    # mixed: one row that fits and one that does not
    mixed = np.concatenate([tiny[:1], big[:1]])
    compare_raw(orc, gpu, mixed, *params_pair(mode=1, max_depth=3))
    compare_bed(orc, gpu, mixed, *params_pair(mode=1, max_depth=3, merge_distance=10))


def test_medium_buckets_through_the_cta_merge_classes():
    """Rows with hundreds to thousands of boxes per query sequence that still fit: every CTA class of the
    bucket merge is reached from the lists the walk leaves on the device."""
    cfg = ix.synth_cfg(3, 1, 400000, 400, 12, 300, 17)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc, gpu = build_both(recs, runs, offs, lens, names)
    small_hits = 0
    for span in (3000, 12000, 40000, 90000, 150000):
        rows = np.array([(s, 20000 + 1000 * s, 20000 + 1000 * s + span) for s in range(3)], ix.RANGE_DTYPE)
        for d, ms in ((1000, True), (0, False), (-1, True)):
            compare_bed(orc, gpu, rows, *params_pair(mode=1, max_depth=2, merge_distance=d, merge_strands=ms))
            small_hits += took_small_path(gpu, True)
        compare_raw(orc, gpu, rows, *params_pair(mode=1, max_depth=3))
    assert small_hits >= 6


def test_switch_selects_the_batched_path(world, monkeypatch):
    cfg, orc, gpu, bed = world
    monkeypatch.setenv("IMPGX_NO_SMALL_BFS", "1")
    compare_raw(orc, gpu, bed[:2], *params_pair(mode=1, max_depth=2))
    assert not took_small_path(gpu, False)
    monkeypatch.delenv("IMPGX_NO_SMALL_BFS")
    compare_raw(orc, gpu, bed[:2], *params_pair(mode=1, max_depth=2))
    assert took_small_path(gpu, False)


@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_dfs_calls(world, depth):
    """query_transitive_dfs (reference src/impg.rs:2057-2309) in the single launch: the stack of the row lives in
    its scratch, every round pops one range, walks it and re-sorts / joins the stack."""
    cfg, orc, gpu, bed = world
    for k in range(0, 30, 3):
        compare_raw(orc, gpu, bed[k:k + 1], *params_pair(mode=2, max_depth=depth))
        assert took_small_path(gpu, False)
        compare_raw(orc, gpu, bed[k:k + 3], *params_pair(mode=2, max_depth=depth))
        assert took_small_path(gpu, False)
    compare_bed(orc, gpu, bed[:20], *params_pair(mode=2, max_depth=depth, merge_distance=1000))
    assert took_small_path(gpu, True)


def test_dfs_options_and_masks(world):
    cfg, orc, gpu, bed = world
    b = bed[40:52]
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=0, min_transitive_len=0, min_dist=0))
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=3, min_output_length=2500))
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=3, min_transitive_len=2000, min_dist=500))
    mask = np.zeros(12, np.uint8)
    mask[[0, 2, 3, 7, 8]] = 1
    compare_raw(orc, gpu, b, *params_pair(mode=2, max_depth=3, subset_mask=mask))
    compare_bed(orc, gpu, b, *params_pair(mode=2, max_depth=0, merge_distance=0, merge_strands=False))
    assert took_small_path(gpu, True)
    rng = np.random.default_rng(5)
    for density, depth in ((0.5, 2), (0.9, 3), (0.97, 0)):
        m = random_mask(rng, 12, 60000, density)
        for k in range(0, 24, 4):
            compare_raw(orc, gpu, bed[k:k + 4], *params_pair(mode=2, max_depth=depth, masked_regions=m))
            assert took_small_path(gpu, False)
            compare_bed(orc, gpu, bed[k:k + 4], *params_pair(mode=2, max_depth=depth, masked_regions=m, merge_distance=500))

