"""PAF parsing of libimpgx (reference src/paf.rs:118-194 + the one-time CIGAR decode) on the CPU, through a test hook:
the plain loop and the multi-threaded path for large files must extract exactly what the oracle's parser extracts —
records, sequence ids by first appearance, decoded runs, the reference's CIGAR byte offsets — and report the same
first error."""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import _impg_format as F
import _oracle as O
import impg_b200 as ix

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAFS = sorted(glob.glob(os.path.join(GOLD, "*.paf")))

SCRIPT = r"""
import ctypes as C, json, sys
import numpy as np
import impg_b200 as ix
L = ix.lib()
L.impgx_debug_parse_paf.restype = C.c_long
L.impgx_debug_parse_paf_seqs.restype = C.c_void_p
path = sys.argv[1].encode()
nr, ns = C.c_uint64(0), C.c_uint32(0)
n = L.impgx_debug_parse_paf(path, None, None, None, None, None, C.byref(nr), C.byref(ns))
if n < 0:
    print(json.dumps({"error": L.impgx_last_error().decode(), "code": int(n)}))
    sys.exit(0)
recs = np.zeros(n, ix.RECORD_DTYPE); off = np.zeros(n + 1, np.uint64); runs = np.zeros(max(nr.value, 1), np.uint32)
co = np.zeros(max(n, 1), np.uint64); cl = np.zeros(max(n, 1), np.uint64)
p = lambda a: a.ctypes.data_as(C.c_void_p)
assert L.impgx_debug_parse_paf(path, p(recs), p(off), p(runs), p(co), p(cl), None, None) == n
ptr = L.impgx_debug_parse_paf_seqs(path); seqs = C.string_at(ptr).decode(); L.impgx_free(C.c_void_p(ptr))
print(json.dumps({"recs": [[int(x) for x in r] for r in recs.tolist()], "off": off.tolist(), "runs": runs[:nr.value].tolist(),
                  "cg_off": co[:n].tolist(), "cg_len": cl[:n].tolist(), "seqs": seqs}))
"""


def product_parse(path, parallel):
    """Runs the hook in a fresh process: the serial / parallel switch is read once per process."""
    import json
    env = dict(os.environ, IMPGX_PAF_PARALLEL_MIN_BYTES="0" if parallel else str(1 << 40), OMP_NUM_THREADS="5",
               PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", SCRIPT, path], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


def oracle_parse(path):
    orc = O.Index.from_paf(path)
    recs, offs, runs, lens, names = orc.export()
    seqs = "".join(f"{n}\t{int(l)}\n" for n, l in zip(names, lens))
    return {"recs": [[int(x) for x in r] for r in recs.tolist()], "off": offs.tolist(), "runs": runs.tolist(), "seqs": seqs}


@pytest.mark.parametrize("paf", PAFS, ids=[os.path.basename(p) for p in PAFS])
def test_both_parse_paths_match_the_oracle_on_the_fixtures(paf):
    want = oracle_parse(paf)
    ref = F.parse_paf_like_reference([paf])[2]
    for parallel in (False, True):
        got = product_parse(paf, parallel)
        for k in ("recs", "off", "runs", "seqs"):
            assert got[k] == want[k], (k, parallel)
        assert got["cg_off"] == [r[8] for r in ref] and got["cg_len"] == [r[9] for r in ref]


def test_parallel_path_on_a_larger_synthetic_paf_with_quirks(tmp_path):
    cfg = ix.synth_cfg(4, 2, 300000, 9, 60, 300, 5)
    recs, runs, offs, lens, names = ix.synth_generate(cfg)
    ops = "=XIDM"
    lines = []
    for i in range(len(recs)):
        r = recs[i]
        cg = "".join(f"{int(v) & 0x1FFFFFFF}{ops[int(v) >> 29]}" for v in runs[int(offs[i]):int(offs[i + 1])])
        tags = ["tp:A:P", f"cg:Z:{cg}", "NM:i:3"] if i % 3 else [f"cg:Z:{cg}"]
        lines.append("\t".join([names[r["query_id"]], str(lens[r["query_id"]]), str(r["query_start"]), "+" + str(r["query_end"]),
                                "-" if r["strand"] else "+", names[r["target_id"]], str(lens[r["target_id"]]),
                                str(r["target_start"]), str(r["target_end"]), "10", "20", "60"] + tags))
    for ending, last_nl in (("\n", True), ("\r\n", True), ("\n", False)):
        p = tmp_path / "s.paf"
        p.write_bytes((ending.join(lines) + (ending if last_nl else "")).encode())
        a, b = product_parse(str(p), False), product_parse(str(p), True)
        assert a == b and len(a["recs"]) == len(recs)
        if ending == "\n":
            # (on CRLF files the reference's recorded offsets drift by a byte per line, so it — and the oracle —
            # cannot find the CIGARs again; libimpgx decodes them while parsing and is not affected)
            want = oracle_parse(str(p))
            for k in ("recs", "off", "runs", "seqs"):
                assert b[k] == want[k], k


def test_both_paths_report_the_same_first_error(tmp_path):
    good = open(PAFS[0]).read().splitlines()
    cases = {"fields": "a\tb\tc", "number": good[1].replace("\t", "\tx", 2), "strand": "\t".join(good[2].split("\t")[:4] + ["?"] + good[2].split("\t")[5:]),
             "cigar": good[3].replace("cg:Z:", "cg:Z:5Q"), "nocg": "\t".join(f for f in good[4].split("\t") if not f.startswith("cg:Z:")),
             "blank": ""}
    for name, bad in cases.items():
        p = tmp_path / (name + ".paf")
        body = good[:5] + [bad] + good[5:8] + ["also\tbad"]
        p.write_text("\n".join(body) + "\n")
        a, b = product_parse(str(p), False), product_parse(str(p), True)
        assert "error" in a and a == b, (name, a, b)
        assert "line 6 of" in a["error"] and a["code"] == ix.E_PARSE


def test_gzip_input_parses_like_plain_text(tmp_path):
    import gzip
    gz = tmp_path / "x.paf.gz"
    gz.write_bytes(gzip.compress(open(PAFS[1], "rb").read()))
    plain = product_parse(PAFS[1], False)
    assert product_parse(str(gz), False) == plain and product_parse(str(gz), True) == plain
