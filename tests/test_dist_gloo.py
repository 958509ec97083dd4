"""N > 1 path on CPU: world_size-2 gloo processes exercise the row sharding and
the max-over-ranks / sum reductions bench.py uses (no data-path collective)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from impg_b200 import dist as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            cuts = [D.shard_rows(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_reductions(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np
        from impg_b200 import dist as D
        import impg_b200 as ix
        rank, local, world = D.env_rank()
        D.init("gloo")
        # every rank draws its own rows (weak scaling) and reports a fake time
        cfg = ix.synth_cfg(4, 1, 40000, 4, 30, 100, 1)
        bed = ix.synth_bed(cfg, 50, seed=D.rank_seed(2, rank))
        lo, hi = D.shard_rows(101, rank, world)
        t = D.max_over_ranks([10.0 + rank, 5.0 - rank])
        total = D.gather_row_counts(len(bed))
        out = dict(rank=rank, t=t, total=total, lo=lo, hi=hi, first=int(bed["start"][0]))
        open(os.path.join({str(tmp_path)!r}, f"out{{rank}}.json"), "w").write(json.dumps(out))
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    o = [json.load(open(tmp_path / f"out{k}.json")) for k in range(2)]
    assert o[0]["t"] == [11.0, 5.0] and o[1]["t"] == [11.0, 5.0]
    assert o[0]["total"] == 100 and o[1]["total"] == 100
    assert (o[0]["lo"], o[0]["hi"], o[1]["lo"], o[1]["hi"]) == (0, 51, 51, 101)
    assert o[0]["first"] != o[1]["first"]  # different BEDs per rank


def test_two_rank_gloo_sharded_host_logic(tmp_path):
    """Host side of the target-sharded path over gloo: both ranks derive the same
    owner map, their record subsets cover every alignment (each at most twice),
    and rank-local result columns gathered with all_gather_object reassemble into
    the per-row, per-sequence order."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np
        from impg_b200 import dist as D
        import impg_b200 as ix
        rank, local, world = D.env_rank()
        D.init("gloo")
        cfg = ix.synth_cfg(5, 2, 40000, 4, 30, 100, 1)
        recs, runs, offs, lens, names = ix.synth_generate(cfg)
        owner = ix.assign_owners(recs, offs, len(lens), world)
        keep = ix.shard_records(recs, offs, owner, rank)
        # fake rank-local BED rows: one row per (input row, owned sequence)
        n_rows = 4
        mine = np.nonzero(owner == rank)[0].astype(np.uint32)
        q = np.tile(mine, n_rows)
        ro = (np.arange(n_rows + 1) * len(mine)).astype(np.uint64)
        val = (np.repeat(np.arange(n_rows), len(mine)) * 1000 + q).astype(np.int32)
        cols = dict(row_offsets=ro, q_id=q, q_first=val, q_last=val, t_id=q, t_first=val, t_last=val)
        parts = D.gather_columns(cols)
        merged = ix.merge_shard_columns(parts)
        out = dict(rank=rank, owner=owner.tolist(), keep=keep.tolist(), merged=merged["q_first"].tolist(),
                   ro=merged["row_offsets"].tolist(), hist=D.owner_histogram(owner, world).tolist())
        open(os.path.join({str(tmp_path)!r}, f"out{{rank}}.json"), "w").write(json.dumps(out))
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    o = [json.load(open(tmp_path / f"out{k}.json")) for k in range(2)]
    assert o[0]["owner"] == o[1]["owner"] and sorted(set(o[0]["owner"])) == [0, 1]
    assert o[0]["hist"] == [5, 5]
    n_aln = 5 * 4 * 2 * 4
    cover = np.bincount(np.array(o[0]["keep"] + o[1]["keep"]), minlength=n_aln)
    assert cover.min() >= 1 and cover.max() <= 2
    want = [r * 1000 + s for r in range(4) for s in range(10)]
    assert o[0]["merged"] == want and o[1]["merged"] == want
    assert o[0]["ro"] == [0, 10, 20, 30, 40]


def test_two_rank_gloo_partition_over_shards(tmp_path):
    """dist.partition_sharded over gloo, the device replaced by the oracle: every rank steps through the same windows,
    answers each with the BED rows of the sequences IT owns (what impgx_query_batch_bed_sharded returns on a shard),
    the parts are all-gathered and merged, and both ranks arrive at the oracle's partitions."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, os.path.join({ROOT!r}, "tests"))
        import numpy as np
        from impg_b200 import dist as D
        import impg_b200 as ix
        import _oracle as O
        rank, local, world = D.env_rank()
        D.init("gloo")
        cfg = ix.synth_cfg(5, 2, 40000, 5, 30, 200, 3)
        recs, runs, offs, lens, names = ix.synth_generate(cfg)
        owner = ix.assign_owners(recs, offs, len(lens), world)
        orc = O.Index.build(recs, runs, offs, lens, names=names)

        class Cols:
            def __init__(self, c):
                self.c = c
            def columns(self):
                return self.c

        class OracleShard:  # the part of a sharded answer this rank would hold: rows of the sequences it owns
            n_seqs = len(lens)
            def seq_len(self, i):
                return int(lens[i])
            def seq_name(self, i):
                return names[i]
            def query_batch_bed_sharded(self, comm, window, gp):
                mo = np.ctypeslib.as_array((ix.C.c_uint64 * (len(lens) + 1)).from_address(gp.mask_offsets)).copy()
                mr = np.ctypeslib.as_array((ix.C.c_int32 * max(2 * int(mo[-1]), 2)).from_address(gp.mask_ranges)).copy()
                qp = O.make_params(mode=O.MODE_DFS if gp.mode == ix.MODE_DFS else O.MODE_BFS, max_depth=gp.max_depth,
                                   min_transitive_len=gp.min_transitive_len, min_dist=gp.min_distance_between_ranges,
                                   masked_regions=(mo, mr), merge_distance=gp.merge_distance, merge_strands=True)
                res, roff = orc.query_batch(window, qp, bed_merge=True)
                c = res.columns()
                sel = owner[c["q_id"]] == rank
                out = {{k: np.asarray(c[k])[sel] for k in ("q_id", "q_first", "q_last", "t_id", "t_first", "t_last")}}
                out["row_offsets"] = np.array([0, int(sel.sum())], np.uint64)
                return Cols(out)

        kw = dict(window_size=15000, merge_distance=1000)
        got = D.partition_sharded(OracleShard(), None, ix.make_partition_params(**kw))
        want = orc.partition(O.make_partition_params(**kw))
        norm = [[p, s, min(a, b), max(a, b)] for p, s, a, b in want["rows"]]
        out = dict(rank=rank, equal=[list(r) for r in got.rows()] == norm, windows=int(got.n_windows),
                   want_windows=len(want["windows"]), owned=int((owner == rank).sum()))
        open(os.path.join({str(tmp_path)!r}, f"out{{rank}}.json"), "w").write(json.dumps(out))
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       env=env, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    o = [json.load(open(tmp_path / f"out{k}.json")) for k in range(2)]
    for k in range(2):
        assert o[k]["equal"], o[k]
        assert o[k]["windows"] == o[k]["want_windows"] > 1
        assert o[k]["owned"] > 0  # both ranks contribute rows
