"""impgx_partition on the reference's fixture PAFs against the committed golden text
(tests/golden/golden_partition.json, produced by the oracle in the build container): the CUDA path checked without
the oracle's help."""
import pytest

import impg_b200 as ix

pytestmark = pytest.mark.gpu


def test_partition_of_the_fixture_pafs_against_committed_golden():
    """impgx_partition on the reference's fixture PAFs against tests/golden/golden_partition.json (text produced by the
    oracle in the build container): the CUDA path is checked without the oracle's help."""
    import hashlib
    import json
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    pgold = json.load(open(os.path.join(gold, "golden_partition.json")))
    for fixture in sorted(pgold):
        g = ix.Impg.from_paf(os.path.join(gold, fixture))
        for c in pgold[fixture]["partition"]:
            got = g.partition(ix.make_partition_params(**c["params"]))
            text = got.format_bed(g)
            assert hashlib.sha256(text.encode()).hexdigest() == c["sha256"], (fixture, c["params"])
            assert (int(got.n_windows), int(got.n_partitions)) == (c["windows"], c["n_partitions"])
