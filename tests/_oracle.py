"""ctypes view of oracle/liboracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module. It never touches the product library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

OPS = "=XIDM"


def run(op, length):
    """CigarOp::new(len, op) packing (reference src/impg.rs:81-93)."""
    return (OPS.index(op) << 29) | int(length)


def cigar(text):
    """'10=5I' -> np.uint32 array in the reference packing."""
    out, n = [], 0
    for ch in text:
        if ch.isdigit():
            n = n * 10 + int(ch)
        else:
            out.append(run(ch, n))
            n = 0
    return np.array(out, dtype=np.uint32)


def cigar_str(runs):
    return "".join(f"{int(v) & 0x1FFFFFFF}{OPS[int(v) >> 29]}" for v in runs)


class Record(C.Structure):
    _fields_ = [
        ("query_id", C.c_uint32),
        ("target_id", C.c_uint32),
        ("query_start", C.c_int32),
        ("query_end", C.c_int32),
        ("target_start", C.c_int32),
        ("target_end", C.c_int32),
        ("strand", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("query_id", "<u4"),
        ("target_id", "<u4"),
        ("query_start", "<i4"),
        ("query_end", "<i4"),
        ("target_start", "<i4"),
        ("target_end", "<i4"),
        ("strand", "<u4"),
        ("reserved", "<u4"),
    ]
)
RANGE_DTYPE = np.dtype([("target_id", "<u4"), ("start", "<i4"), ("end", "<i4")])


class Params(C.Structure):
    _fields_ = [
        ("mode", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("min_transitive_len", C.c_int32),
        ("min_distance_between_ranges", C.c_int32),
        ("min_output_length", C.c_int32),
        ("store_cigar", C.c_uint32),
        ("min_identity", C.c_double),
        ("subset_mask", C.c_void_p),
        ("merge_distance", C.c_int32),
        ("merge_strands", C.c_uint32),
        ("mask_offsets", C.c_void_p),
        ("mask_ranges", C.c_void_p),
    ]


MODE_QUERY, MODE_BFS, MODE_DFS = 0, 1, 2
MODE_MULTI_QUERY, MODE_MULTI_BFS, MODE_MULTI_DFS = 3, 4, 5


def make_params(mode=MODE_QUERY, max_depth=2, min_transitive_len=101, min_dist=10, min_output_length=-1,
                store_cigar=False, min_identity=float("nan"), subset_mask=None, merge_distance=0,
                merge_strands=True, masked_regions=None):
    p = Params()
    p.mode = mode
    p.max_depth = max_depth
    p.min_transitive_len = min_transitive_len
    p.min_distance_between_ranges = min_dist
    p.min_output_length = min_output_length
    p.store_cigar = 1 if store_cigar else 0
    p.min_identity = min_identity
    if subset_mask is not None:
        subset_mask = np.ascontiguousarray(subset_mask, dtype=np.uint8)
        p._mask_keepalive = subset_mask
        p.subset_mask = subset_mask.ctypes.data
    else:
        p.subset_mask = None
    p.merge_distance = merge_distance
    p.merge_strands = 1 if merge_strands else 0
    if masked_regions is not None:
        mo = np.ascontiguousarray(masked_regions[0], dtype=np.uint64)
        mr = np.ascontiguousarray(masked_regions[1], dtype=np.int32)
        p._keep_mask = (mo, mr)
        p.mask_offsets, p.mask_ranges = mo.ctypes.data, mr.ctypes.data
    else:
        p.mask_offsets = p.mask_ranges = None
    return p


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


class PartitionParams(C.Structure):
    """impgx_partition_params (include/impgx.h)."""
    _fields_ = [("window_size", C.c_uint64), ("starting_seqs", C.c_void_p), ("n_starting_seqs", C.c_size_t),
                ("selection_mode", C.c_char_p), ("merge_distance", C.c_int32), ("min_missing_size", C.c_int32),
                ("min_boundary_distance", C.c_int32), ("transitive_dfs", C.c_uint32), ("max_depth", C.c_uint32),
                ("min_transitive_len", C.c_int32), ("min_distance_between_ranges", C.c_int32),
                ("rehome_singletons", C.c_uint32), ("min_identity", C.c_double), ("multi_impg", C.c_uint32),
                ("reserved", C.c_uint32)]


def make_partition_params(window_size, merge_distance, starting_seqs=None, selection_mode="longest",
                          min_missing_size=3000, min_boundary_distance=3000, transitive_dfs=False, max_depth=2,
                          min_transitive_len=101, min_distance_between_ranges=10, rehome_singletons=True,
                          min_identity=None, multi_impg=False):
    """Defaults are `impg partition`'s (reference src/main.rs:4765-4880, :4259-4279)."""
    p = PartitionParams()
    p.window_size = window_size
    if starting_seqs is not None and len(starting_seqs):
        a = np.ascontiguousarray(starting_seqs, dtype=np.uint32)
        p._keep = a
        p.starting_seqs, p.n_starting_seqs = a.ctypes.data, len(a)
    else:
        p.starting_seqs, p.n_starting_seqs = None, 0
    p.selection_mode = selection_mode.encode() if selection_mode is not None else None
    p.merge_distance = merge_distance
    p.min_missing_size = min_missing_size
    p.min_boundary_distance = min_boundary_distance
    p.transitive_dfs = 1 if transitive_dfs else 0
    p.max_depth = max_depth
    p.min_transitive_len = min_transitive_len
    p.min_distance_between_ranges = min_distance_between_ranges
    p.rehome_singletons = 1 if rehome_singletons else 0
    p.min_identity = float("nan") if min_identity is None else float(min_identity)
    p.multi_impg = 1 if multi_impg else 0
    p.reserved = 0
    return p


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "oracle.cpp")
        ):
            build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_project.restype = C.c_int
        L.orc_parse_cigar.restype = C.c_long
        L.orc_identity.restype = C.c_double
        L.orc_sorted_ranges_insert.restype = C.c_size_t
        for f in ("orc_index_build", "orc_index_from_paf", "orc_perform_query", "orc_results_from_arrays",
                  "orc_query_batch", "orc_multi_build", "orc_multi_query_batch", "orc_partition", "orc_multi_partition",
                  "orc_partition_format_bed"):
            getattr(L, f).restype = C.c_void_p
        L.orc_index_num_seqs.restype = C.c_uint32
        L.orc_index_num_records.restype = C.c_size_t
        L.orc_index_num_runs.restype = C.c_size_t
        L.orc_index_seq_name.restype = C.c_char_p
        L.orc_index_seq_len.restype = C.c_uint64
        L.orc_index_seq_id.restype = C.c_long
        L.orc_results_len.restype = C.c_size_t
        L.orc_results_cigar_len.restype = C.c_size_t
        L.orc_stab_order.restype = C.c_size_t
        L.orc_format.restype = C.c_void_p
        L.orc_run_batch.restype = C.c_double
        L.orc_run_batch_rows_parallel.restype = C.c_double
        L.orc_refine.restype = C.c_int64
        L.orc_populate_cigar_cache.restype = C.c_uint64
        L.orc_query_with_cache.restype = C.c_void_p
        L.orc_index_attach_cigar_file.restype = C.c_int
        L.orc_partition_error.restype = C.c_char_p
        L.orc_partition_len.restype = C.c_size_t
        L.orc_partition_num_windows.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def project(req, record, ops):
    """project_target_range_through_alignment. record = (t_start,t_end,q_start,q_end,strand_rev)."""
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    out4 = np.zeros(4, dtype=np.int32)
    out_ops = np.zeros(max(1, len(ops)), dtype=np.uint32)
    n_out = C.c_size_t(0)
    ok = lib().orc_project(
        C.c_int32(req[0]), C.c_int32(req[1]), C.c_int32(record[0]), C.c_int32(record[1]), C.c_int32(record[2]),
        C.c_int32(record[3]), C.c_int(1 if record[4] else 0), _p(ops), C.c_size_t(len(ops)), _p(out4), _p(out_ops),
        C.byref(n_out))
    if not ok:
        return None
    return (int(out4[0]), int(out4[1]), out_ops[: n_out.value].copy(), int(out4[2]), int(out4[3]))


def parse_cigar(text):
    b = text.encode() if isinstance(text, str) else text
    out = np.zeros(max(1, len(b)), dtype=np.uint32)
    n = lib().orc_parse_cigar(b, C.c_size_t(len(b)), _p(out), C.c_size_t(len(out)))
    if n < 0:
        raise ValueError("invalid CIGAR")
    return out[:n].copy()


def invert(ops, strand_rev):
    ops = np.array(ops, dtype=np.uint32)
    lib().orc_invert(_p(ops), C.c_size_t(len(ops)), C.c_int(1 if strand_rev else 0))
    return ops


def identity(ops):
    ops = np.ascontiguousarray(ops, dtype=np.uint32)
    return lib().orc_identity(_p(ops), C.c_size_t(len(ops)))


def parse_subsequence_coordinates(name):
    """reference src/main.rs:4642-4659 -> (base, start) or None"""
    buf = C.create_string_buffer(len(name.encode()) + 1)
    st = C.c_int32(0)
    r = lib().orc_parse_subsequence_coordinates(name.encode(), buf, C.c_size_t(len(buf)), C.byref(st))
    return (buf.value.decode(), st.value) if r else None


def subset_matches(list_text, name):
    """SubsetFilter::matches after parse_subset_filter (reference src/subset_filter.rs). -> (bool, entry_count)"""
    n = C.c_size_t(0)
    r = lib().orc_subset_matches(list_text.encode(), name.encode(), C.byref(n))
    return bool(r), n.value


def sorted_ranges_insert(ranges, seq_len, min_dist, new):
    cap = len(ranges) + 2
    buf = np.zeros(2 * cap, dtype=np.int32)
    for i, (s, e) in enumerate(ranges):
        buf[2 * i], buf[2 * i + 1] = s, e
    n = C.c_size_t(len(ranges))
    pieces = np.zeros(2 * cap, dtype=np.int32)
    k = lib().orc_sorted_ranges_insert(_p(buf), C.byref(n), C.c_size_t(cap), C.c_int32(seq_len), C.c_int32(min_dist),
                                       C.c_int32(new[0]), C.c_int32(new[1]), _p(pieces), C.c_size_t(cap))
    return ([(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)],
            [(int(pieces[2 * i]), int(pieces[2 * i + 1])) for i in range(k)])


class Results:
    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def __del__(self):
        if self.h:
            lib().orc_results_free(self.h)
            self.h = None

    def __len__(self):
        return lib().orc_results_len(self.h)

    def columns(self):
        n = len(self)
        nc = lib().orc_results_cigar_len(self.h)
        cols = {
            "q_id": np.zeros(n, np.uint32), "q_first": np.zeros(n, np.int32), "q_last": np.zeros(n, np.int32),
            "t_id": np.zeros(n, np.uint32), "t_first": np.zeros(n, np.int32), "t_last": np.zeros(n, np.int32),
            "cigar_offsets": np.zeros(n + 1, np.uint64), "cigar_runs": np.zeros(max(nc, 1), np.uint32),
        }
        lib().orc_results_copy(self.h, _p(cols["q_id"]), _p(cols["q_first"]), _p(cols["q_last"]), _p(cols["t_id"]),
                               _p(cols["t_first"]), _p(cols["t_last"]), _p(cols["cigar_offsets"]),
                               _p(cols["cigar_runs"]))
        cols["cigar_runs"] = cols["cigar_runs"][:nc]
        return cols

    def tuples(self):
        c = self.columns()
        out = []
        for i in range(len(c["q_id"])):
            a, b = int(c["cigar_offsets"][i]), int(c["cigar_offsets"][i + 1])
            out.append((int(c["q_id"][i]), int(c["q_first"][i]), int(c["q_last"][i]), int(c["t_id"][i]),
                        int(c["t_first"][i]), int(c["t_last"][i]), cigar_str(c["cigar_runs"][a:b])))
        return out

    def drop_first(self):
        lib().orc_results_drop_first(self.h)

    def merge_query(self, d, merge_strands=True):
        lib().orc_merge_query(self.h, C.c_int32(d), C.c_int(1 if merge_strands else 0))

    def merge_2d(self, d):
        lib().orc_merge_2d(self.h, C.c_int32(d))

    def merge_cigar(self, d):
        lib().orc_merge_cigar(self.h, C.c_int32(d))

    @staticmethod
    def from_tuples(rows):
        """rows: (q_id,q_first,q_last,t_id,t_first,t_last[,cigar_text])"""
        n = len(rows)
        qid = np.array([r[0] for r in rows], np.uint32)
        qf = np.array([r[1] for r in rows], np.int32)
        ql = np.array([r[2] for r in rows], np.int32)
        tid = np.array([r[3] for r in rows], np.uint32)
        tf = np.array([r[4] for r in rows], np.int32)
        tl = np.array([r[5] for r in rows], np.int32)
        cigs = [cigar(r[6]) if len(r) > 6 else np.zeros(0, np.uint32) for r in rows]
        off = np.zeros(n + 1, np.uint64)
        for i, cg in enumerate(cigs):
            off[i + 1] = off[i] + len(cg)
        flat = np.concatenate(cigs) if n and off[n] else np.zeros(1, np.uint32)
        flat = np.ascontiguousarray(flat, np.uint32)
        return Results(lib().orc_results_from_arrays(C.c_size_t(n), _p(qid), _p(qf), _p(ql), _p(tid), _p(tf), _p(tl),
                                                     _p(off), _p(flat)))


def _partition_result(h):
    L = lib()
    err = L.orc_partition_error(h).decode()
    if err:
        raise ValueError(err)
    n, nw = L.orc_partition_len(h), L.orc_partition_num_windows(h)
    pnum, seq = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    first, last = np.zeros(n, np.int32), np.zeros(n, np.int32)
    win = np.zeros(max(nw, 1), RANGE_DTYPE)
    L.orc_partition_copy(h, _p(pnum), _p(seq), _p(first), _p(last), _p(win))
    npart, pbp, tbp = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    L.orc_partition_totals(h, C.byref(npart), C.byref(pbp), C.byref(tbp))
    return {"rows": list(zip(pnum.tolist(), seq.tolist(), first.tolist(), last.tolist())),
            "windows": [tuple(int(x) for x in w) for w in win[:nw]], "n_partitions": npart.value,
            "partitioned_bp": pbp.value, "total_bp": tbp.value}


class MultiIndex:
    """The oracle's MultiImpg (reference src/multi_impg.rs): one sub-index per alignment file."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle multi index build failed")
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_multi_free(self.h)
            self.h = None

    @staticmethod
    def build(records, runs, run_offsets, seq_lens, file_of_record, n_files, bidirectional=True):
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        runs = np.ascontiguousarray(runs, dtype=np.uint32)
        run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
        seq_lens = np.ascontiguousarray(seq_lens, dtype=np.uint64)
        fo = np.ascontiguousarray(file_of_record, dtype=np.uint32)
        assert len(fo) == len(records)
        if len(runs) == 0:
            runs = np.zeros(1, np.uint32)
        return MultiIndex(lib().orc_multi_build(_p(records), C.c_size_t(len(records)), _p(runs), _p(run_offsets),
                                                _p(seq_lens), C.c_uint32(len(seq_lens)), _p(fo), C.c_uint32(n_files),
                                                C.c_int(1 if bidirectional else 0)))

    def partition(self, pp, names):
        """partition_alignments over a MultiImpg (its transitive walk answers the windows)."""
        L = lib()
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        h = C.c_void_p(L.orc_multi_partition(self.h, C.byref(pp), arr, C.c_uint32(len(names))))
        try:
            return _partition_result(h)
        finally:
            L.orc_partition_free(h)

    def query_batch(self, ranges, params, bed_merge=False):
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        offs = np.zeros(len(ranges) + 1, np.uint64)
        res = Results(lib().orc_multi_query_batch(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params),
                                                  C.c_int(1 if bed_merge else 0), _p(offs)))
        return res, offs


class Index:
    """The oracle's Impg (reference src/impg.rs)."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index build failed")
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h)
            self.h = None

    @staticmethod
    def build(records, runs, run_offsets, seq_lens, bidirectional=True, names=None):
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        runs = np.ascontiguousarray(runs, dtype=np.uint32)
        run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
        seq_lens = np.ascontiguousarray(seq_lens, dtype=np.uint64)
        if len(runs) == 0:
            runs = np.zeros(1, np.uint32)
        idx = Index(lib().orc_index_build(_p(records), C.c_size_t(len(records)), _p(runs), _p(run_offsets),
                                          _p(seq_lens), C.c_uint32(len(seq_lens)), C.c_int(1 if bidirectional else 0)))
        if names is not None:
            idx.set_names(names)
        return idx

    @staticmethod
    def from_paf(path, bidirectional=True, faithful=False):
        err = C.create_string_buffer(512)
        h = lib().orc_index_from_paf(path.encode(), C.c_int(1 if bidirectional else 0), C.c_int(1 if faithful else 0),
                                     err, C.c_size_t(512))
        if not h:
            raise ValueError(err.value.decode())
        return Index(h)

    def attach_cigar_file(self, path, offsets, lens):
        offsets = np.ascontiguousarray(offsets, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint64)
        if lib().orc_index_attach_cigar_file(self.h, path.encode(), _p(offsets), _p(lens)) != 0:
            raise OSError("cannot open " + path)

    def set_original_coordinates(self, on=True):
        lib().orc_index_set_original_coordinates(self.h, C.c_int(1 if on else 0))

    def set_faithful(self, on):
        lib().orc_index_set_faithful(self.h, C.c_int(1 if on else 0))

    def set_names(self, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        lib().orc_index_set_names(self.h, arr, C.c_uint32(len(names)))

    @property
    def n_seqs(self):
        return lib().orc_index_num_seqs(self.h)

    @property
    def n_records(self):
        return lib().orc_index_num_records(self.h)

    def seq_name(self, i):
        return lib().orc_index_seq_name(self.h, C.c_uint32(i)).decode()

    def seq_len(self, i):
        return lib().orc_index_seq_len(self.h, C.c_uint32(i))

    def seq_id(self, name):
        r = lib().orc_index_seq_id(self.h, name.encode())
        if r < 0:
            raise KeyError(name)
        return r

    def export(self):
        """(records, run_offsets, runs, seq_lens, names) — to build the product from the same data."""
        n = self.n_records
        recs = np.zeros(n, dtype=RECORD_DTYPE)
        offs = np.zeros(n + 1, dtype=np.uint64)
        nr = lib().orc_index_num_runs(self.h)
        runs = np.zeros(max(nr, 1), dtype=np.uint32)
        lib().orc_index_export(self.h, _p(recs), _p(offs), _p(runs))
        ns = self.n_seqs
        lens = np.array([self.seq_len(i) for i in range(ns)], dtype=np.uint64)
        names = [self.seq_name(i) for i in range(ns)]
        return recs, offs, runs[:nr], lens, names

    def stab_order(self, target_id, s, e, cap=1 << 16):
        aln = np.zeros(cap, np.uint64)
        rev = np.zeros(cap, np.uint8)
        k = lib().orc_stab_order(self.h, C.c_uint32(target_id), C.c_int32(s), C.c_int32(e), _p(aln), _p(rev),
                                 C.c_size_t(cap))
        assert k <= cap
        return list(zip(aln[:k].tolist(), rev[:k].tolist()))

    def perform_query(self, target_id, s, e, params, threads=1):
        return Results(lib().orc_perform_query(self.h, C.c_uint32(target_id), C.c_int32(s), C.c_int32(e),
                                               C.byref(params), C.c_int(threads)))

    def query_batch(self, ranges, params, bed_merge=False):
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        offs = np.zeros(len(ranges) + 1, np.uint64)
        res = Results(lib().orc_query_batch(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params),
                                            C.c_int(1 if bed_merge else 0), _p(offs)))
        return res, offs

    def partition(self, pp, threads=1):
        """partition_alignments (-o bed, single file). Returns a dict: rows = [(partition_num, seq, first, last)]
        with the reference's orientation, windows = every window queried, totals and the partitions.bed text."""
        L = lib()
        h = C.c_void_p(L.orc_partition(self.h, C.byref(pp), C.c_int(threads)))
        try:
            out = _partition_result(h)
            ptr = L.orc_partition_format_bed(self.h, h)
            out["bed"] = C.string_at(ptr).decode()
            L.orc_free(C.c_void_p(ptr))
            return out
        finally:
            L.orc_partition_free(h)

    def format(self, results, fmt, name, d, merge_strands=True):
        """fmt: 'bed' | 'bedpe' | 'paf'. Mutates `results` (merges) like the reference writers."""
        code = {"bed": 0, "bedpe": 1, "paf": 2}[fmt]
        ptr = lib().orc_format(self.h, results.h, C.c_int(code), name.encode(), C.c_int32(d),
                               C.c_int(1 if merge_strands else 0))
        s = C.string_at(ptr).decode()
        lib().orc_free(C.c_void_p(ptr))
        return s

    def run_batch(self, ranges, params, threads=1, fmt="bed"):
        """Reference batch driver timing: returns (seconds, n_results, out_bytes, checksum)."""
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        code = {None: -1, "bed": 0, "bedpe": 1, "paf": 2}[fmt]
        nres, nbytes, csum = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        secs = lib().orc_run_batch(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params), C.c_int(threads),
                                   C.c_int(code), C.byref(nres), C.byref(nbytes), C.byref(csum))
        return secs, nres.value, nbytes.value, csum.value

    def refine(self, loci, params):
        """refine_single_range per locus (reference src/commands/refine.rs:144-410); `params` is an
        impg_b200.RefineParams (the same C struct the product takes). Returns the list of records."""
        loci = np.ascontiguousarray(loci, dtype=RANGE_DTYPE)
        n = len(loci)
        rec = np.zeros((n, 8), np.int64)
        eo = np.zeros(n + 1, np.uint64)
        cap = 1 << 16
        while True:
            es, ea, eb = np.zeros(cap, np.uint32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            tot = lib().orc_refine(self.h, _p(loci), C.c_size_t(n), C.byref(params), _p(rec), _p(eo), _p(es), _p(ea), _p(eb),
                                   C.c_size_t(cap))
            if tot < 0:
                raise ValueError(f"refine failed for locus {-1 - tot}")
            if tot <= cap:
                break
            cap = int(tot)
        keys = ("refined_start", "refined_end", "original_start", "original_end", "applied_left_extension",
                "applied_right_extension", "support_count", "original_support_count")
        out = []
        for i in range(n):
            a, b = int(eo[i]), int(eo[i + 1])
            out.append(dict({k: int(rec[i, j]) for j, k in enumerate(keys)},
                            support_entities=list(zip(es[a:b].tolist(), ea[a:b].tolist(), eb[a:b].tolist()))))
        return out

    def refine_parallel(self, loci, params, threads):
        """The loci in parallel like run_refine's par_iter; returns (seconds, records as an (n, 8) int64 array)."""
        loci = np.ascontiguousarray(loci, dtype=RANGE_DTYPE)
        rec = np.zeros((len(loci), 8), np.int64)
        L = lib()
        L.orc_refine_parallel.restype = C.c_double
        secs = L.orc_refine_parallel(self.h, _p(loci), C.c_size_t(len(loci)), C.byref(params), C.c_int(threads), _p(rec))
        if secs < 0:
            raise ValueError("refine failed for a locus")
        return secs, rec

    def populate_cigar_cache(self, target_id, s, e):
        return int(lib().orc_populate_cigar_cache(self.h, C.c_uint32(target_id), C.c_int32(s), C.c_int32(e)))

    def query_with_cache(self, target_id, s, e, cache_range, store_cigar=False, min_identity=None):
        """populate_cigar_cache over `cache_range`, then Impg::query_with_cache(s, e) through that cache."""
        h = lib().orc_query_with_cache(self.h, C.c_uint32(target_id), C.c_int32(s), C.c_int32(e),
                                       C.c_int(1 if store_cigar else 0),
                                       C.c_double(float("nan") if min_identity is None else min_identity),
                                       C.c_int32(cache_range[0]), C.c_int32(cache_range[1]))
        return Results(h)

    def run_batch_rows_parallel(self, ranges, params, threads=1, fmt="bed"):
        """The "CPU-batched" driver: rows in parallel, one thread per row. Returns (seconds, n_results, out_bytes)."""
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        code = {None: -1, "bed": 0}[fmt]
        nres, nbytes = C.c_uint64(0), C.c_uint64(0)
        secs = lib().orc_run_batch_rows_parallel(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params),
                                                 C.c_int(threads), C.c_int(code), C.byref(nres), C.byref(nbytes))
        return secs, nres.value, nbytes.value
