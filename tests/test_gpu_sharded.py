"""Target-sharded index against the CPU oracle, through the C
ABI. The shards of these tests are virtual ranks that share cuda:0 and exchange
hits / frontier ranges through the in-process transport, so the whole exchange
logic (routing by owner, global frontier order, stage-A-local / stage-B-remote
BED merge) is exercised on a 1-GPU box; the NCCL transport is covered by
test_nccl_two_ranks when two devices are present."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import _oracle as O
import impg_b200 as ix
from test_gpu_parity import params_pair

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def compare_bed_sharded(orc, sh, ranges, o_params, g_params):
    ores, ooffs = orc.query_batch(ranges, o_params, bed_merge=True)
    oc = ores.columns()
    parts = sh.query_batch_bed_parts(ranges, g_params)
    gc = ix.merge_shards(parts).columns()
    assert gc["row_offsets"].tolist() == ooffs.tolist()
    for k in ("q_id", "q_first", "q_last"):
        assert (gc[k] == oc[k]).all(), k
    # the numpy reassembly used across processes agrees with the library's
    pc = ix.merge_shard_columns([p.columns() for p in parts])
    assert pc["row_offsets"].tolist() == ooffs.tolist()
    for k in ("q_id", "q_first", "q_last"):
        assert (pc[k] == oc[k]).all(), k
    # every rank returned only rows of sequences it owns
    for r, p in enumerate(parts):
        assert (sh.owner[p.columns()["q_id"]] == r).all()
    return len(oc["q_id"])


@pytest.fixture(scope="module")
def world():
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc = O.Index.build(recs, runs, offs, lens, names=names)
    bed = ix.synth_bed(cfg, 200, seed=9, min_len=200, max_len=12000)
    shards = {n: ix.ShardedImpg.from_records(recs, runs, offs, lens, [0] * n, names=names) for n in (1, 2, 3, 5)}
    return cfg, orc, shards, bed


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 5])
@pytest.mark.parametrize("mode,depth", [(0, 1), (1, 1), (1, 2), (1, 3), (1, 0)])
def test_sharded_bed_matches_oracle(world, n_ranks, mode, depth):
    cfg, orc, shards, bed = world
    n = compare_bed_sharded(orc, shards[n_ranks], bed[:120], *params_pair(mode=mode, max_depth=depth, merge_distance=1000))
    assert n > 120


@pytest.mark.parametrize("d,merge_strands", [(0, True), (1000, False), (-1, True), (50000, True)])
def test_sharded_merge_options(world, d, merge_strands):
    cfg, orc, shards, bed = world
    for n in (2, 3):
        compare_bed_sharded(orc, shards[n], bed[:100], *params_pair(mode=1, max_depth=2, merge_distance=d,
                                                                    merge_strands=merge_strands))


def test_sharded_filters(world):
    cfg, orc, shards, bed = world
    sh = shards[3]
    b = bed[:100]
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=3, min_transitive_len=0, min_dist=0, merge_distance=0))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=3, min_transitive_len=2000, min_dist=500,
                                                 merge_distance=100))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=2, min_output_length=2500, merge_distance=1000))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=0, min_output_length=3000, merge_distance=1000))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=3, min_identity=0.95, merge_distance=1000))
    mask = np.zeros(12, np.uint8)
    mask[[0, 2, 3, 7, 8]] = 1
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=3, subset_mask=mask, merge_distance=1000))


def test_sharded_matches_unsharded_at_medium_scale():
    """Bigger than the oracle handles quickly: the sharded path must reproduce
    the single-index CUDA path bit for bit (which the oracle tests pin)."""
    cfg = ix.synth_cfg(12, 2, 400000, 20, 60, 100, 11)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    one = ix.Impg.from_records(recs, runs, offs, lens, names=names)
    bed = ix.synth_bed(cfg, 3000, seed=5)
    g = ix.make_params(mode=ix.MODE_BFS, max_depth=2, merge_distance=1000)
    want = one.query_batch_bed(bed, g).columns()
    for n in (2, 4):
        sh = ix.ShardedImpg.from_records(recs, runs, offs, lens, [0] * n, names=names)
        got = sh.query_batch_bed(bed, g).columns()
        assert got["row_offsets"].tolist() == want["row_offsets"].tolist()
        for k in ("q_id", "q_first", "q_last"):
            assert (got[k] == want[k]).all(), (n, k)
        st = sh.stats()
        assert sum(s["liftovers"] for s in st) == one.stats()["liftovers"]
        tr = [c.traffic() for c in sh.comms]
        assert sum(t["bytes_sent"] for t in tr) == sum(t["bytes_received"] for t in tr) > 0
        # shard memory: every alignment is walked by at most two owners
        assert sum(s.device_bytes for s in sh.shards) < 2.2 * one.device_bytes


def test_sharded_edge_cases(world):
    cfg, orc, shards, bed = world
    sh = shards[2]
    o, g = params_pair(mode=1, max_depth=2, merge_distance=0)
    # empty batch: still a collective call
    parts = sh.query_batch_bed_parts(np.zeros(0, ix.RANGE_DTYPE), g)
    assert all(p.n_rows == 0 and p.n_results == 0 for p in parts)
    # all rows on one sequence (one rank owns every seed), and a single row
    one_seq = np.array([(3, 100 + 50 * k, 5000 + 50 * k) for k in range(40)], ix.RANGE_DTYPE)
    compare_bed_sharded(orc, sh, one_seq, o, g)
    compare_bed_sharded(orc, sh, one_seq[:1], o, g)
    # rows too short to expand: nothing but the seeds, frontier empty on every rank
    tiny = np.array([(s, 10, 60) for s in range(12)], ix.RANGE_DTYPE)
    compare_bed_sharded(orc, sh, tiny, o, g)
    # unsupported on a shard: raw results, the MultiImpg walks, unsorted --no-merge output
    with pytest.raises(ix.ImpgxError):
        sh.shards[0].query_batch(one_seq, g)
    for bad in (ix.make_params(mode=ix.MODE_MULTI_DFS), ix.make_params(mode=ix.MODE_BFS, merge_distance=-1, merge_strands=False)):
        fresh = ix.ShardedImpg(sh.shards, ix.Comm.local_group(2), sh.owner)  # a failed collective poisons its group
        with pytest.raises(ix.ImpgxError) as e:
            fresh.query_batch_bed_parts(one_seq, bad)
        assert e.value.code == ix.E_UNSUPPORTED
    # invalid rows are rejected on every rank
    fresh = ix.ShardedImpg(sh.shards, ix.Comm.local_group(2), sh.owner)
    with pytest.raises(ix.ImpgxError) as e:
        fresh.query_batch_bed_parts(np.array([(99, 0, 10)], ix.RANGE_DTYPE), g)
    assert e.value.code == ix.E_INVALID


def test_sharded_row_batches_agree(world, monkeypatch):
    """Row batching must be cut identically on every rank (sizes come from an all-gather)."""
    cfg, orc, shards, bed = world
    monkeypatch.setenv("IMPGX_HITS_PER_BATCH", "3000")
    compare_bed_sharded(orc, shards[3], bed, *params_pair(mode=1, max_depth=2, merge_distance=1000))
    monkeypatch.setenv("IMPGX_ROWS_PER_BATCH", "7")
    compare_bed_sharded(orc, shards[2], bed[:50], *params_pair(mode=1, max_depth=3, merge_distance=1000))


@pytest.mark.skipif(ix.device_count() < 2, reason="needs two CUDA devices")
def test_nccl_two_ranks(tmp_path):
    """One process per GPU over NCCL (torchrun): merged output equals the oracle's."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, pickle
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np, torch, torch.distributed as dist
        import impg_b200 as ix
        from impg_b200 import dist as D
        rank, local, world = D.env_rank()
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = D.nccl_comm(rank, world, local)
        cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
        recs, runs, offs, lens, names = ix.synth_generate(cfg)
        owner = ix.assign_owners(recs, offs, len(lens), world)
        shard = ix.Impg.from_records_shard(recs, runs, offs, lens, owner, rank, world, names=names, device=local)
        bed = ix.synth_bed(cfg, 150, seed=9, min_len=200, max_len=12000)
        g = ix.make_params(mode=ix.MODE_BFS, max_depth=3, merge_distance=1000)
        cols = shard.query_batch_bed_sharded(comm, bed, g).columns()
        parts = D.gather_columns(cols)
        if rank == 0:
            pickle.dump(ix.merge_shard_columns(parts), open(os.path.join({str(tmp_path)!r}, "out.pkl"), "wb"))
        dist.barrier()
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    import pickle
    got = pickle.load(open(tmp_path / "out.pkl", "rb"))
    cfg = ix.synth_cfg(6, 2, 60000, 8, 30, 300, 3)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    orc = O.Index.build(recs, runs, offs, lens, names=names)
    bed = ix.synth_bed(cfg, 150, seed=9, min_len=200, max_len=12000)
    o, _ = params_pair(mode=1, max_depth=3, merge_distance=1000)
    ores, ooffs = orc.query_batch(bed, o, bed_merge=True)
    oc = ores.columns()
    assert got["row_offsets"].tolist() == ooffs.tolist()
    for k in ("q_id", "q_first", "q_last"):
        assert (got[k] == oc[k]).all(), k


@pytest.mark.parametrize("n_ranks", [2, 3, 5])
def test_sharded_masked_regions(world, n_ranks):
    """masked_regions (what partition passes, reference src/commands/partition.rs:359-391; src/impg.rs:2331-2373) on a
    sharded index: every rank derives the unmasked seed pieces of all rows, the owner of a row's target walks them."""
    from test_gpu_parity import random_mask
    cfg, orc, shards, bed = world
    sh = shards[n_ranks]
    rng = np.random.default_rng(40 + n_ranks)
    b = bed[:90]
    for density, depth, d in ((0.5, 2, 1000), (0.9, 3, 0), (0.0, 2, 1000), (0.97, 0, 100)):
        mask = random_mask(rng, 12, 60000, density)
        compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=depth, masked_regions=mask, merge_distance=d))
    mask = random_mask(rng, 12, 60000, 0.8)
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=2, masked_regions=mask, merge_distance=-1, merge_strands=True))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=1, max_depth=3, masked_regions=mask, merge_distance=500,
                                                 merge_strands=False, min_transitive_len=0, min_dist=0))
    # a mask that covers every row entirely: nothing comes back
    full = ix.mask_csr({s: [(0, 60000)] for s in range(12)}, 12)
    parts = sh.query_batch_bed_parts(b, params_pair(mode=1, masked_regions=full, merge_distance=1000)[1])
    assert sum(p.n_results for p in parts) == 0


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 5])
@pytest.mark.parametrize("depth", [1, 2, 3, 0])
def test_sharded_dfs(world, n_ranks, depth):
    """Transitive DFS (reference src/impg.rs:2057-2309) on a sharded index: replicated stacks, hits routed to the owner
    of the sequence they land on, uncovered pieces all-gathered; the boxes carry (round, visit rank) as their ordinal."""
    cfg, orc, shards, bed = world
    sh = shards[n_ranks]
    compare_bed_sharded(orc, sh, bed[:70], *params_pair(mode=2, max_depth=depth, merge_distance=1000))


def test_sharded_dfs_options_and_masks(world):
    from test_gpu_parity import random_mask
    cfg, orc, shards, bed = world
    sh = shards[3]
    b = bed[:50]
    compare_bed_sharded(orc, sh, b, *params_pair(mode=2, max_depth=0, min_transitive_len=0, min_dist=0, merge_distance=0))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=2, max_depth=3, min_output_length=2500, merge_distance=500,
                                                 merge_strands=False))
    compare_bed_sharded(orc, sh, b, *params_pair(mode=2, max_depth=3, min_transitive_len=2000, min_dist=500,
                                                 merge_distance=-1, merge_strands=True))
    rng = np.random.default_rng(91)
    for density, depth in ((0.5, 2), (0.9, 3), (0.97, 0)):
        mask = random_mask(rng, 12, 60000, density)
        compare_bed_sharded(orc, sh, b, *params_pair(mode=2, max_depth=depth, masked_regions=mask, merge_distance=1000))
    # one sequence only, a single row, rows too short to walk on
    one_seq = np.array([(3, 100 + 50 * k, 5000 + 50 * k) for k in range(20)], ix.RANGE_DTYPE)
    compare_bed_sharded(orc, sh, one_seq, *params_pair(mode=2, max_depth=2, merge_distance=0))
    compare_bed_sharded(orc, sh, one_seq[:1], *params_pair(mode=2, max_depth=0, merge_distance=0))
    tiny = np.array([(s, 10, 60) for s in range(12)], ix.RANGE_DTYPE)
    compare_bed_sharded(orc, sh, tiny, *params_pair(mode=2, max_depth=2, merge_distance=0))


@pytest.mark.parametrize("n_ranks", [2, 3])
def test_partition_over_the_sharded_index(world, n_ranks):
    """`impg partition -o bed` (reference src/commands/partition.rs:158-712) with every window answered by the
    collective masked walk of the sharded index: the same partitions as the oracle's."""
    cfg, orc, shards, bed = world
    sh = shards[n_ranks]
    for kw in (dict(window_size=20000, merge_distance=1000), dict(window_size=7000, merge_distance=0, max_depth=3),
               dict(window_size=30000, merge_distance=5000, transitive_dfs=True, max_depth=2),
               dict(window_size=15000, merge_distance=100, selection_mode="total", min_missing_size=500,
                    min_boundary_distance=200)):
        want = orc.partition(O.make_partition_params(**kw))
        got = sh.partition(ix.make_partition_params(**kw))
        norm = [(p, s, min(x, y), max(x, y)) for p, s, x, y in want["rows"]]
        assert got.rows() == norm, kw
        assert got.n_windows == len(want["windows"])
    with pytest.raises(ix.ImpgxError) as e:
        sh.partition(ix.make_partition_params(window_size=20000, merge_distance=-1))
    assert e.value.code == ix.E_UNSUPPORTED

