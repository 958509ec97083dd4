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
