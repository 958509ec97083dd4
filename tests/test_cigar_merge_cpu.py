"""BEDPE / PAF output (merge_adjusted_intervals with its CIGAR surgery + the writers, reference
src/main.rs:12563-12845, :13014-13180, :11894-12103) on the CPU: libimpgx's host code, reached through a test hook
that takes rows as arrays, against the oracle on random chains of alignments — contiguous, overlapping, gapped, both
strands, shuffled. The device delivers such rows in tests/test_gpu_parity.py; here the rows are adversarial."""
import ctypes as C
import random

import numpy as np

import _oracle as O
import impg_b200 as ix


def random_cigar(rnd, allow_m):
    ops, last = [], None
    for _ in range(rnd.randrange(1, 7)):
        op = rnd.choice("==XID" + ("M" if allow_m else ""))
        if op == last:
            continue
        ops.append((op, rnd.randrange(1, 40)))
        last = op
    return ops


def make_rows(rnd, n_seqs):
    rows = []
    for _ in range(rnd.randrange(1, 5)):  # chains on a (query, target, strand)
        q, t = rnd.randrange(n_seqs), rnd.randrange(n_seqs)
        rev = rnd.random() < 0.4
        qs, ts = rnd.randrange(0, 2000), rnd.randrange(0, 2000)
        allow_m = rnd.random() < 0.2
        for _ in range(rnd.randrange(1, 6)):
            cg = random_cigar(rnd, allow_m)
            ql = sum(l for o, l in cg if o in "=XIM")
            tl = sum(l for o, l in cg if o in "=XDM")
            if ql == 0 or tl == 0:
                continue
            text = "".join(f"{l}{o}" for o, l in cg)
            step = rnd.choice([0, 0, 0, 3, 60, 2000, -5, -15])
            if rev:
                if qs - ql < 0:
                    break
                rows.append((q, qs, qs - ql, t, ts, ts + tl, text))
                qs = qs - ql - step
            else:
                rows.append((q, qs, qs + ql, t, ts, ts + tl, text))
                qs = qs + ql + step
            ts = ts + tl + rnd.choice([0, 0, step, step, 7, -3])
            if qs < 0 or ts < 0:
                break
    rnd.shuffle(rows)
    return rows


def product_text(names, lens, rows, fmt, name, d):
    L = ix.lib()
    L.impgx_debug_format_rows.restype = C.c_void_p
    n = len(rows)

    def arr(k, dt):
        return np.array([r[k] for r in rows], dtype=dt)

    cigs = [O.cigar(r[6]) for r in rows]
    off = np.zeros(n + 1, np.uint64)
    for i, cg in enumerate(cigs):
        off[i + 1] = off[i] + len(cg)
    flat = np.ascontiguousarray(np.concatenate(cigs) if n else np.zeros(1, np.uint32), np.uint32)
    nm = (C.c_char_p * len(names))(*[s.encode() for s in names])

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    a = [arr(0, np.uint32), arr(1, np.int32), arr(2, np.int32), arr(3, np.uint32), arr(4, np.int32), arr(5, np.int32)]
    ptr = L.impgx_debug_format_rows(nm, p(lens), C.c_uint32(len(names)), C.c_size_t(n), *[p(x) for x in a], p(off), p(flat),
                                    name.encode(), C.c_int32(d), C.c_int(1 if fmt == "bedpe" else 2), C.c_int(0))
    assert ptr, L.impgx_last_error()
    s = C.string_at(ptr).decode()
    L.impgx_free(C.c_void_p(ptr))
    return s


def test_bedpe_paf_merge_matches_oracle_on_random_chains():
    rnd = random.Random(77)
    n_seqs = 4
    names = [f"s{i}#1#chr{i}" for i in range(n_seqs)]
    lens = np.array([100000 + i for i in range(n_seqs)], np.uint64)
    orc = O.Index.build(np.zeros(0, O.RECORD_DTYPE), np.zeros(0, np.uint32), np.zeros(1, np.uint64), lens, names=names)
    merged_something = 0
    for trial in range(400):
        rows = make_rows(rnd, n_seqs)
        if not rows:
            continue
        for d in (0, 50, 1000, -1):
            for fmt in ("bedpe", "paf"):
                want = orc.format(O.Results.from_tuples(rows), fmt, "reg", d)
                got = product_text(names, lens, rows, fmt, "reg", d)
                assert got == want, (trial, d, fmt, rows)
                merged_something += want.count("\n") < len(rows)
    assert merged_something > 200  # the chains do exercise the merge paths
