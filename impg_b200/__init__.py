"""impg_b200 — host-side mirror of impg's query interface for the projection
path, over libimpgx's C ABI (include/impgx.h).

The class `Impg` mirrors the reference's `ImpgIndex` trait for this path
(reference src/impg_index.rs:21-121): `query`, `query_transitive_bfs`,
sequence-index accessors, plus the batch entry that replaces the serial BED
loop of `impg query -b` (reference src/main.rs:7435-7456).

There is no CPU fallback: if libimpgx.so is missing the import fails, and
without a CUDA device every compute call raises ImpgxError(NO_DEVICE).
"""
import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libimpgx.so")

OK, E_INVALID, E_NO_DEVICE, E_CUDA, E_NOMEM, E_IO, E_PARSE, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7
MODE_QUERY, MODE_BFS, MODE_DFS = 0, 1, 2
MODE_MULTI_QUERY, MODE_MULTI_BFS, MODE_MULTI_DFS = 3, 4, 5  # MultiImpg semantics (reference src/multi_impg.rs)

RECORD_DTYPE = np.dtype([("query_id", "<u4"), ("target_id", "<u4"), ("query_start", "<i4"), ("query_end", "<i4"),
                         ("target_start", "<i4"), ("target_end", "<i4"), ("strand", "<u4"), ("reserved", "<u4")])
RANGE_DTYPE = np.dtype([("target_id", "<u4"), ("start", "<i4"), ("end", "<i4")])

OPS = "=XIDM"


class ImpgxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[impgx {code}] {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("max_depth", C.c_uint32), ("min_transitive_len", C.c_int32),
                ("min_distance_between_ranges", C.c_int32), ("min_output_length", C.c_int32),
                ("store_cigar", C.c_uint32), ("min_identity", C.c_double), ("subset_mask", C.c_void_p),
                ("merge_distance", C.c_int32), ("merge_strands", C.c_uint32),
                ("mask_offsets", C.c_void_p), ("mask_ranges", C.c_void_p)]


class View(C.Structure):
    _fields_ = [("n_rows", C.c_size_t), ("n_results", C.c_size_t), ("row_offsets", C.c_void_p),
                ("q_id", C.c_void_p), ("q_first", C.c_void_p), ("q_last", C.c_void_p), ("t_id", C.c_void_p),
                ("t_first", C.c_void_p), ("t_last", C.c_void_p), ("cigar_offsets", C.c_void_p),
                ("cigar_runs", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("stab_ranges", C.c_uint64), ("lift_launches", C.c_uint64),
                ("liftovers", C.c_uint64), ("lift_runs", C.c_uint64), ("lift_bytes", C.c_uint64),
                ("lift_touched_bytes", C.c_uint64), ("lift_window_runs", C.c_uint64),
                ("results", C.c_uint64), ("merged", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("lift_ms", C.c_float), ("stab_ms", C.c_float), ("fold_ms", C.c_float), ("merge_ms", C.c_float),
                ("total_ms", C.c_float), ("exchange_ms", C.c_float), ("merge_kernel_ms", C.c_float),
                ("reserved0", C.c_float), ("merge_boxes", C.c_uint64), ("exchange_bytes", C.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PartitionParams(C.Structure):
    """impgx_partition_params (include/impgx.h)."""
    _fields_ = [("window_size", C.c_uint64), ("starting_seqs", C.c_void_p), ("n_starting_seqs", C.c_size_t),
                ("selection_mode", C.c_char_p), ("merge_distance", C.c_int32), ("min_missing_size", C.c_int32),
                ("min_boundary_distance", C.c_int32), ("transitive_dfs", C.c_uint32), ("max_depth", C.c_uint32),
                ("min_transitive_len", C.c_int32), ("min_distance_between_ranges", C.c_int32),
                ("rehome_singletons", C.c_uint32), ("min_identity", C.c_double), ("multi_impg", C.c_uint32),
                ("reserved", C.c_uint32)]


class PartitionView(C.Structure):
    _fields_ = [("n_intervals", C.c_size_t), ("n_partitions", C.c_size_t), ("n_windows", C.c_uint64),
                ("partitioned_bp", C.c_uint64), ("total_bp", C.c_uint64), ("partition_num", C.c_void_p),
                ("seq_id", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p)]


def make_partition_params(window_size, merge_distance, starting_seqs=None, selection_mode="longest",
                          min_missing_size=3000, min_boundary_distance=3000, transitive_dfs=False, max_depth=2,
                          min_transitive_len=101, min_distance_between_ranges=10, rehome_singletons=True,
                          min_identity=None, multi_impg=False):
    """Arguments of partition_alignments (reference src/commands/partition.rs:158-181); defaults are
    `impg partition`'s CLI defaults (src/main.rs:4765-4880, :4259-4279). -d has no default there."""
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


class RefineParams(C.Structure):
    """impgx_refine_params (include/impgx.h): the options of `impg refine` (reference src/main.rs:4411-4450)."""
    _fields_ = [("span_bp", C.c_int32), ("extension_step", C.c_int32), ("max_extension", C.c_double),
                ("support_level", C.c_uint32), ("merge_distance", C.c_int32), ("min_identity", C.c_double),
                ("transitive", C.c_uint32), ("max_depth", C.c_uint32), ("min_transitive_len", C.c_int32),
                ("min_distance_between_ranges", C.c_int32), ("subset_mask", C.c_void_p),
                ("blacklist_offsets", C.c_void_p), ("blacklist_ranges", C.c_void_p)]


class RefineView(C.Structure):
    _fields_ = [("n", C.c_size_t), ("refined_start", C.c_void_p), ("refined_end", C.c_void_p),
                ("original_start", C.c_void_p), ("original_end", C.c_void_p), ("applied_left_extension", C.c_void_p),
                ("applied_right_extension", C.c_void_p), ("support_count", C.c_void_p),
                ("original_support_count", C.c_void_p), ("entity_offsets", C.c_void_p), ("entity_seq", C.c_void_p),
                ("entity_start", C.c_void_p), ("entity_end", C.c_void_p), ("candidates_evaluated", C.c_uint64),
                ("batches", C.c_uint64)]


def make_refine_params(span_bp=1000, max_extension=0.5, extension_step=1000, support_level=0, merge_distance=0,
                       min_identity=None, transitive=0, max_depth=2, min_transitive_len=101,
                       min_distance_between_ranges=10, subset_mask=None, blacklist=None, n_seqs=None):
    """`blacklist`: {seq_id: [(start, end), ...]} as the blacklist BED gives them (needs n_seqs)."""
    p = RefineParams()
    p.span_bp, p.extension_step, p.max_extension = span_bp, extension_step, max_extension
    p.support_level, p.merge_distance = support_level, merge_distance
    p.min_identity = float("nan") if min_identity is None else min_identity
    p.transitive, p.max_depth = transitive, max_depth
    p.min_transitive_len, p.min_distance_between_ranges = min_transitive_len, min_distance_between_ranges
    keep = []
    if subset_mask is not None:
        m = np.ascontiguousarray(subset_mask, np.uint8)
        keep.append(m)
        p.subset_mask = m.ctypes.data
    if blacklist is not None:
        offs, rng = mask_csr(blacklist, n_seqs)
        keep += [offs, rng]
        p.blacklist_offsets, p.blacklist_ranges = offs.ctypes.data, rng.ctypes.data
    p._keep = keep
    return p


class SynthCfg(C.Structure):  # impgx_synth_cfg (csrc/synth_core.h)
    _fields_ = [("genomes", C.c_uint32), ("contigs", C.c_uint32), ("contig_len", C.c_uint32), ("tiles", C.c_uint32),
                ("eq_mean", C.c_uint32), ("rev_permille", C.c_uint32), ("seed", C.c_uint64),
                ("partners", C.c_uint32), ("reserved", C.c_uint32)]


def mask_csr(masked_regions, n_seqs):
    """{seq id: [(start, end), ...]} (sorted, disjoint — a SortedRanges per sequence) -> CSR arrays."""
    offs = np.zeros(n_seqs + 1, np.uint64)
    flat = []
    for s in range(n_seqs):
        rs = masked_regions.get(s, ())
        offs[s + 1] = offs[s] + np.uint64(len(rs))
        for a, b in rs:
            flat += [a, b]
    return offs, np.array(flat if flat else [0, 0], dtype=np.int32)


def make_params(mode=MODE_QUERY, max_depth=2, min_transitive_len=101, min_distance_between_ranges=10,
                min_output_length=None, store_cigar=False, min_identity=None, subset_mask=None, merge_distance=0,
                merge_strands=True, masked_regions=None):
    """Defaults are the reference's CLI defaults (src/main.rs:4259-4410)."""
    p = Params()
    p.mode = mode
    p.max_depth = max_depth
    p.min_transitive_len = min_transitive_len
    p.min_distance_between_ranges = min_distance_between_ranges
    p.min_output_length = -1 if min_output_length is None else min_output_length
    p.store_cigar = 1 if store_cigar else 0
    p.min_identity = float("nan") if min_identity is None else float(min_identity)
    if subset_mask is not None:
        m = np.ascontiguousarray(subset_mask, dtype=np.uint8)
        p._keep = m
        p.subset_mask = m.ctypes.data
    else:
        p.subset_mask = None
    p.merge_distance = merge_distance
    p.merge_strands = 1 if merge_strands else 0
    if masked_regions is not None:  # (offsets, ranges) CSR, see mask_csr
        mo = np.ascontiguousarray(masked_regions[0], dtype=np.uint64)
        mr = np.ascontiguousarray(masked_regions[1], dtype=np.int32)
        p._keep_mask = (mo, mr)
        p.mask_offsets, p.mask_ranges = mo.ctypes.data, mr.ctypes.data
    else:
        p.mask_offsets = p.mask_ranges = None
    return p


_lib = None


def lib():
    """Loads libimpgx.so (built in-tree by `make -C impg_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C impg_b200/csrc` "
                              "(libimpgx has no Python or CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.impgx_last_error.restype = C.c_char_p
        L.impgx_index_num_seqs.restype = C.c_uint32
        L.impgx_index_num_entries.restype = C.c_uint64
        L.impgx_index_device_bytes.restype = C.c_uint64
        L.impgx_index_seq_name.restype = C.c_char_p
        L.impgx_index_seq_len.restype = C.c_uint64
        L.impgx_parse_cigar.restype = C.c_long
        L.impgx_format_bed.restype = C.c_void_p
        L.impgx_format_bed_batch.restype = C.c_void_p
        L.impgx_format_bedpe.restype = C.c_void_p
        L.impgx_format_paf.restype = C.c_void_p
        L.impgx_partitions_format_bed.restype = C.c_void_p
        L.impgx_impg_num_seqs.restype = C.c_uint32
        L.impgx_impg_seq_name.restype = C.c_char_p
        L.impgx_impg_seq_len.restype = C.c_uint64
        L.impgx_impg_num_entries.restype = C.c_uint64
        L.impgx_impg_num_records.restype = C.c_uint64
        L.impgx_debug_host_columns.restype = C.c_long
        L.impgx_debug_host_columns_shard.restype = C.c_long
        _lib = L
    return _lib


_synth = None
SYNTH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libimpgx_synth.so")


def synth_lib():
    """The synthetic workload generator and the CIGAR text writer of the CPU reference arm: bench / test tooling in a
    library of its own (csrc/synth_tool.cpp), so that generating a workload does not load the product."""
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: build it with `make -C impg_b200/csrc`")
        L = C.CDLL(SYNTH_LIB_PATH)
        L.impgx_synth_num_alignments.restype = C.c_uint64
        L.impgx_synth_last_error.restype = C.c_char_p
        _synth = L
    return _synth


def _check_synth(code):
    if code != 0:
        raise ImpgxError(code, synth_lib().impgx_synth_last_error().decode(errors="replace"))


def _check(code):
    if code != 0:
        raise ImpgxError(code, lib().impgx_last_error().decode(errors="replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().impgx_device_count()


def run(op, length):
    return (OPS.index(op) << 29) | int(length)


def cigar_str(runs):
    return "".join(f"{int(v) & 0x1FFFFFFF}{OPS[int(v) >> 29]}" for v in runs)


def parse_cigar(text):
    b = text.encode() if isinstance(text, str) else bytes(text)
    out = np.zeros(max(1, len(b)), dtype=np.uint32)
    n = lib().impgx_parse_cigar(b, C.c_size_t(len(b)), _p(out), C.c_size_t(len(out)))
    if n < 0:
        raise ImpgxError(n, lib().impgx_last_error().decode())
    return out[:n].copy()


class Results:
    """Column view of a result set (AdjustedInterval rows, reference src/impg.rs:225)."""

    def __init__(self, handle, on_device=False):
        self.h = C.c_void_p(handle)
        self.on_device = on_device
        v = View()
        if on_device:
            _check(lib().impgx_results_device_view(self.h, C.byref(v)))
        else:
            _check(lib().impgx_results_view(self.h, C.byref(v)))
        self.view = v
        self.n_rows = v.n_rows
        self.n_results = v.n_results

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_results_free(self.h)
            self.h = None

    def _arr(self, ptr, n, dtype):
        if n == 0 or not ptr:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()

    def columns(self):
        assert not self.on_device
        v, n = self.view, self.n_results
        cols = {
            "row_offsets": self._arr(v.row_offsets, self.n_rows + 1, np.uint64),
            "q_id": self._arr(v.q_id, n, np.uint32), "q_first": self._arr(v.q_first, n, np.int32),
            "q_last": self._arr(v.q_last, n, np.int32), "t_id": self._arr(v.t_id, n, np.uint32),
            "t_first": self._arr(v.t_first, n, np.int32), "t_last": self._arr(v.t_last, n, np.int32),
        }
        if v.cigar_offsets:
            cols["cigar_offsets"] = self._arr(v.cigar_offsets, n + 1, np.uint64)
            nc = int(cols["cigar_offsets"][-1]) if n else 0
            cols["cigar_runs"] = self._arr(v.cigar_runs, nc, np.uint32)
        return cols

    def row_tuples(self, row, cols=None):
        c = cols or self.columns()
        a, b = int(c["row_offsets"][row]), int(c["row_offsets"][row + 1])
        out = []
        for i in range(a, b):
            cg = ""
            if "cigar_offsets" in c:
                cg = cigar_str(c["cigar_runs"][int(c["cigar_offsets"][i]):int(c["cigar_offsets"][i + 1])])
            out.append((int(c["q_id"][i]), int(c["q_first"][i]), int(c["q_last"][i]), int(c["t_id"][i]),
                        int(c["t_first"][i]), int(c["t_last"][i]), cg))
        return out


class Impg:
    """GPU-resident index + query interface (mirrors reference `Impg` / `ImpgIndex`)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_index_free(self.h)
            self.h = None

    # -- construction (Impg::from_multi_alignment_records, reference src/impg.rs:1535)
    @classmethod
    def from_records(cls, records, runs, run_offsets, seq_lens, names=None, bidirectional=True, device=0):
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        runs = np.ascontiguousarray(runs, dtype=np.uint32)
        run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
        seq_lens = np.ascontiguousarray(seq_lens, dtype=np.uint64)
        if len(runs) == 0:
            runs = np.zeros(1, np.uint32)
        h = C.c_void_p()
        _check(lib().impgx_index_build(_p(records), C.c_size_t(len(records)), _p(runs), _p(run_offsets), _p(seq_lens),
                                       C.c_uint32(len(seq_lens)), C.c_int(1 if bidirectional else 0), C.c_int(device),
                                       C.byref(h)))
        idx = cls(h.value)
        if names is not None:
            idx.set_names(names)
        return idx

    @classmethod
    def from_paf(cls, path, bidirectional=True, device=0):
        h = C.c_void_p()
        _check(lib().impgx_index_from_paf(path.encode(), C.c_int(1 if bidirectional else 0), C.c_int(device),
                                          C.byref(h)))
        return cls(h.value)

    # -- SequenceIndex accessors (reference src/seqidx.rs)
    @classmethod
    def from_impg(cls, impg_path, alignment_files, device=0):
        """Impg::load_from_file (reference src/impg.rs:1777-1850) + the CIGARs of the alignment files."""
        arr = (C.c_char_p * len(alignment_files))(*[p.encode() for p in alignment_files])
        h = C.c_void_p()
        _check(lib().impgx_index_from_impg(impg_path.encode(), arr, C.c_size_t(len(alignment_files)), C.c_int(device),
                                           C.byref(h)))
        return cls(h.value)

    def set_original_coordinates(self, on=True):
        """--original-sequence-coordinates for the BED / BEDPE writers (reference src/main.rs:4661-4678)."""
        _check(lib().impgx_index_set_original_coordinates(self.h, C.c_int(1 if on else 0)))

    def set_names(self, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        _check(lib().impgx_index_set_names(self.h, arr, C.c_uint32(len(names))))

    @property
    def n_seqs(self):
        return lib().impgx_index_num_seqs(self.h)

    @property
    def n_entries(self):
        return lib().impgx_index_num_entries(self.h)

    @property
    def device_bytes(self):
        return lib().impgx_index_device_bytes(self.h)

    def seq_name(self, i):
        s = lib().impgx_index_seq_name(self.h, C.c_uint32(i))
        return s.decode() if s else None

    def seq_len(self, i):
        return lib().impgx_index_seq_len(self.h, C.c_uint32(i))

    def seq_id(self, name):
        out = C.c_uint32()
        _check(lib().impgx_index_seq_id(self.h, name.encode(), C.byref(out)))
        return out.value

    # -- batch entry (replaces the BED loop, reference src/main.rs:7435-7456)
    def query_batch(self, ranges, params):
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        h = C.c_void_p()
        _check(lib().impgx_query_batch(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params), C.byref(h)))
        return Results(h.value)

    def query_batch_bed(self, ranges, params):
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        h = C.c_void_p()
        _check(lib().impgx_query_batch_bed(self.h, _p(ranges), C.c_size_t(len(ranges)), C.byref(params), C.byref(h)))
        return Results(h.value)

    def query_batch_bed_device(self, d_ranges_ptr, n, params, stream=0):
        """`d_ranges_ptr`: device pointer (int) to n impgx_range structs already in HBM."""
        h = C.c_void_p()
        _check(lib().impgx_query_batch_bed_device(self.h, C.c_void_p(d_ranges_ptr), C.c_size_t(n), C.byref(params),
                                                  C.c_void_p(stream), C.byref(h)))
        return Results(h.value, on_device=True)

    # -- target-sharded index: this object is ONE shard; queries are collective
    @classmethod
    def from_records_shard(cls, records, runs, run_offsets, seq_lens, owner, rank, n_ranks, names=None,
                           bidirectional=True, device=0):
        records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        runs = np.ascontiguousarray(runs, dtype=np.uint32)
        run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
        seq_lens = np.ascontiguousarray(seq_lens, dtype=np.uint64)
        owner = np.ascontiguousarray(owner, dtype=np.uint32)
        assert len(owner) == len(seq_lens)
        if len(runs) == 0:
            runs = np.zeros(1, np.uint32)
        h = C.c_void_p()
        _check(lib().impgx_index_build_shard(_p(records), C.c_size_t(len(records)), _p(runs), _p(run_offsets),
                                             _p(seq_lens), C.c_uint32(len(seq_lens)),
                                             C.c_int(1 if bidirectional else 0), C.c_int(device), _p(owner),
                                             C.c_uint32(rank), C.c_uint32(n_ranks), C.byref(h)))
        idx = cls(h.value)
        if names is not None:
            idx.set_names(names)
        return idx

    def query_batch_bed_sharded(self, comm, ranges, params):
        """Collective: every rank calls it with the same ranges/params; returns this rank's BED rows."""
        ranges = np.ascontiguousarray(ranges, dtype=RANGE_DTYPE)
        h = C.c_void_p()
        _check(lib().impgx_query_batch_bed_sharded(self.h, comm.h, _p(ranges), C.c_size_t(len(ranges)),
                                                   C.byref(params), C.byref(h)))
        return Results(h.value)

    def query_batch_bed_sharded_device(self, comm, d_ranges_ptr, n, params, stream=0):
        h = C.c_void_p()
        _check(lib().impgx_query_batch_bed_sharded_device(self.h, comm.h, C.c_void_p(d_ranges_ptr), C.c_size_t(n),
                                                          C.byref(params), C.c_void_p(stream), C.byref(h)))
        return Results(h.value, on_device=True)

    # -- refine and its cache entry points (reference src/commands/refine.rs, src/impg.rs:1930-2035)
    def populate_cigar_cache(self, target_id, range_start, range_end):
        """Number of CIGARs Impg::populate_cigar_cache would cache for the range (they are resident in HBM here)."""
        n = C.c_uint64()
        _check(lib().impgx_populate_cigar_cache(self.h, C.c_uint32(target_id), C.c_int32(range_start), C.c_int32(range_end),
                                                C.byref(n)))
        return n.value

    def query_with_cache_batch(self, target_id, orig_start, orig_end, left, right, params):
        """Impg::query_with_cache for every (left, right) flank pair of one locus, as one batch."""
        left = np.ascontiguousarray(left, np.int32)
        right = np.ascontiguousarray(right, np.int32)
        h = C.c_void_p()
        _check(lib().impgx_query_with_cache_batch(self.h, C.c_uint32(target_id), C.c_int32(orig_start), C.c_int32(orig_end),
                                                  _p(left), _p(right), C.c_size_t(len(left)), C.byref(params), C.byref(h)))
        return Results(h.value)

    def refine(self, loci, params):
        """run_refine over RANGE_DTYPE loci: list of dict records (RefineRecord) + (candidates, batches)."""
        loci = np.ascontiguousarray(loci, dtype=RANGE_DTYPE)
        h = C.c_void_p()
        _check(lib().impgx_refine(self.h, _p(loci), C.c_size_t(len(loci)), C.byref(params), C.byref(h)))
        try:
            v = RefineView()
            _check(lib().impgx_refine_view_get(h, C.byref(v)))
            arr = Results._arr
            n = v.n
            cols = {k: arr(None, getattr(v, k), n, np.int32).copy() for k in
                    ("refined_start", "refined_end", "original_start", "original_end", "applied_left_extension",
                     "applied_right_extension")}
            sup = arr(None, v.support_count, n, np.uint64).copy()
            osup = arr(None, v.original_support_count, n, np.uint64).copy()
            eo = arr(None, v.entity_offsets, n + 1, np.uint64).copy()
            ne = int(eo[-1]) if n else 0
            es = arr(None, v.entity_seq, ne, np.uint32).copy()
            ea = arr(None, v.entity_start, ne, np.int32).copy()
            eb = arr(None, v.entity_end, ne, np.int32).copy()
            recs = []
            for i in range(n):
                a, b = int(eo[i]), int(eo[i + 1])
                recs.append(dict({k: int(c[i]) for k, c in cols.items()}, support_count=int(sup[i]),
                                 original_support_count=int(osup[i]),
                                 support_entities=list(zip(es[a:b].tolist(), ea[a:b].tolist(), eb[a:b].tolist()))))
            return recs, (int(v.candidates_evaluated), int(v.batches))
        finally:
            lib().impgx_refine_results_free(h)

    def stats(self):
        s = Stats()
        _check(lib().impgx_index_stats(self.h, C.byref(s)))
        return s.as_dict()

    # -- per-range interface with the reference's argument names
    def query(self, target_id, range_start, range_end, store_cigar=False, min_gap_compressed_identity=None):
        """Impg::query (reference src/impg.rs:1852-1928)."""
        p = make_params(mode=MODE_QUERY, store_cigar=store_cigar, min_identity=min_gap_compressed_identity)
        r = self.query_batch(np.array([(target_id, range_start, range_end)], dtype=RANGE_DTYPE), p)
        return r.row_tuples(0)

    def query_transitive_bfs(self, target_id, range_start, range_end, max_depth=2, min_transitive_len=101,
                             min_distance_between_ranges=10, min_output_length=None, store_cigar=False,
                             min_gap_compressed_identity=None, subset_mask=None):
        """Impg::query_transitive_bfs (reference src/impg.rs:2311-2597)."""
        p = make_params(mode=MODE_BFS, max_depth=max_depth, min_transitive_len=min_transitive_len,
                        min_distance_between_ranges=min_distance_between_ranges, min_output_length=min_output_length,
                        store_cigar=store_cigar, min_identity=min_gap_compressed_identity, subset_mask=subset_mask)
        r = self.query_batch(np.array([(target_id, range_start, range_end)], dtype=RANGE_DTYPE), p)
        return r.row_tuples(0)

    def format_bed_batch(self, results, names, decode=True, length_only=False):
        """BED text of every row of a merged batch (reference src/main.rs:11849-11892 per row), input order.
        length_only: format, free, and return the byte count (timing without a copy into Python)."""
        arr = names if isinstance(names, C.Array) else (C.c_char_p * len(names))(*[n.encode() for n in names])
        ln = C.c_size_t(0)
        ptr = lib().impgx_format_bed_batch(self.h, results.h, arr, C.byref(ln))
        if not ptr:
            raise ImpgxError(E_INVALID, "format failed")
        if length_only:
            lib().impgx_free(C.c_void_p(ptr))
            return ln.value
        s = C.string_at(ptr, ln.value)
        lib().impgx_free(C.c_void_p(ptr))
        return s.decode() if decode else s

    def partition(self, params):
        """`impg partition -o bed` on this index (reference src/commands/partition.rs:158-712): one masked
        transitive query + BED merge on the device per window."""
        h = C.c_void_p()
        _check(lib().impgx_partition(self.h, C.byref(params), C.byref(h)))
        return Partitions(h.value)

    def format_bed(self, results, row, name):
        ptr = lib().impgx_format_bed(self.h, results.h, C.c_size_t(row), name.encode())
        if not ptr:
            raise ImpgxError(E_INVALID, "format_bed failed")
        s = C.string_at(ptr).decode()
        lib().impgx_free(C.c_void_p(ptr))
        return s


def impg_write(paf_paths, out_path, bidirectional=True):
    """`impg index`: the .impg file stock impg would load for these PAF files (reference
    src/impg.rs:1655-1720). Host only."""
    arr = (C.c_char_p * len(paf_paths))(*[p.encode() for p in paf_paths])
    _check(lib().impgx_impg_write(arr, C.c_size_t(len(paf_paths)), C.c_int(1 if bidirectional else 0), out_path.encode()))


class ImpgFile:
    """A parsed .impg index file (reference src/impg.rs:1777-1850): sequence index + the alignments
    with where their CIGAR text lives in the alignment files. Host only."""

    def __init__(self, path):
        h = C.c_void_p()
        _check(lib().impgx_impg_open(path.encode(), C.byref(h)))
        self.h = h
        L = lib()
        self.version = L.impgx_impg_version(h)
        self.bidirectional = bool(L.impgx_impg_bidirectional(h))
        n = L.impgx_impg_num_seqs(h)
        self.names = [L.impgx_impg_seq_name(h, C.c_uint32(i)).decode() for i in range(n)]
        self.lens = np.array([L.impgx_impg_seq_len(h, C.c_uint32(i)) for i in range(n)], np.uint64)
        self.n_entries = L.impgx_impg_num_entries(h)
        self.n_records = L.impgx_impg_num_records(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_impg_close(self.h)
            self.h = None

    def records(self):
        """(records, file_index, data_offset, data_bytes) in alignment-file order."""
        n = self.n_records
        recs = np.zeros(n, RECORD_DTYPE)
        fi, off, ln = np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        _check(lib().impgx_impg_records(self.h, _p(recs), _p(fi), _p(off), _p(ln)))
        return recs, fi, off, ln


class Partitions:
    """Collected partitions (rows of partitions.bed, reference src/commands/partition.rs:1682-1717)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        v = PartitionView()
        _check(lib().impgx_partitions_view(self.h, C.byref(v)))
        n = v.n_intervals
        arr = Results._arr
        self.partition_num = arr(None, v.partition_num, n, np.uint32)
        self.seq_id = arr(None, v.seq_id, n, np.uint32)
        self.start = arr(None, v.start, n, np.int32)
        self.end = arr(None, v.end, n, np.int32)
        self.n_partitions, self.n_windows = v.n_partitions, v.n_windows
        self.partitioned_bp, self.total_bp = v.partitioned_bp, v.total_bp

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_partitions_free(self.h)
            self.h = None

    def rows(self):
        return list(zip(self.partition_num.tolist(), self.seq_id.tolist(), self.start.tolist(), self.end.tolist()))

    def format_bed(self, impg, partition=-1):
        """partitions.bed (partition < 0) or partition<N>.bed of --separate-files."""
        ptr = lib().impgx_partitions_format_bed(impg.h, self.h, C.c_int64(partition))
        if not ptr:
            raise ImpgxError(E_INVALID, "format failed")
        s = C.string_at(ptr).decode()
        lib().impgx_free(C.c_void_p(ptr))
        return s


class Partitioner:
    """partition_alignments as a stepper over any ImpgIndex implementor: `next()` hands out the next
    window and the current masked_regions (CSR), the caller runs the transitive query and `feed`s the
    query intervals back (reference src/commands/partition.rs:295-580)."""

    def __init__(self, seq_lens, params, names=None):
        lens = np.ascontiguousarray(seq_lens, dtype=np.uint64)
        self.n_seqs = len(lens)
        nm = None
        if names is not None:
            nm = (C.c_char_p * len(names))(*[n.encode() for n in names])
        h = C.c_void_p()
        _check(lib().impgx_partitioner_new(_p(lens), nm, C.c_uint32(len(lens)), C.byref(params), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_partitioner_free(self.h)
            self.h = None

    def next(self):
        """(target_id, start, end), (mask_offsets, mask_ranges) or None when no window is left."""
        w = np.zeros(1, RANGE_DTYPE)
        mo, mr = C.c_void_p(), C.c_void_p()
        r = lib().impgx_partitioner_next(self.h, _p(w), C.byref(mo), C.byref(mr))
        if r < 0:
            _check(r)
        if r == 0:
            return None
        offs = Results._arr(None, mo.value, self.n_seqs + 1, np.uint64)
        rng = Results._arr(None, mr.value, max(2 * int(offs[-1]), 2), np.int32)
        return (int(w[0]["target_id"]), int(w[0]["start"]), int(w[0]["end"])), (offs, rng)

    def feed(self, q_id, q_first, q_last):
        a = np.ascontiguousarray(q_id, np.uint32)
        b = np.ascontiguousarray(q_first, np.int32)
        c = np.ascontiguousarray(q_last, np.int32)
        _check(lib().impgx_partitioner_feed(self.h, C.c_size_t(len(a)), _p(a), _p(b), _p(c)))

    def finish(self):
        h = C.c_void_p()
        _check(lib().impgx_partitioner_finish(self.h, C.byref(h)))
        return Partitions(h.value)


def partition_with(index, pparams, answer):
    """partition_alignments (reference src/commands/partition.rs:158-712) over any index that answers a window:
    the stepper hands out the windows and the masked regions, `answer(window, query_params)` returns the merged BED
    rows of the window as columns (q_id, q_first, q_last) — e.g. a collective call on a target-sharded index whose
    per-rank parts were merged. `index` supplies the sequence lengths and names. Needs -d >= 0: a sharded index
    returns merged rows only."""
    if pparams.merge_distance < 0:
        raise ImpgxError(E_UNSUPPORTED, "partition over a sharded index needs a merge distance >= 0 (--no-merge: raw results)")
    if pparams.multi_impg:
        raise ImpgxError(E_UNSUPPORTED, "partition over a sharded index: the MultiImpg walk is not sharded")
    n = index.n_seqs
    lens = np.array([index.seq_len(i) for i in range(n)], dtype=np.uint64)
    names = [index.seq_name(i) for i in range(n)]
    st = Partitioner(lens, pparams, names if all(x is not None for x in names) else None)
    mi = pparams.min_identity
    while True:
        w = st.next()
        if w is None:
            break
        window, mask = w
        # the query of partition.rs:359-391: no CIGARs, no output length filter, no subset, both strands merged
        qp = make_params(mode=MODE_DFS if pparams.transitive_dfs else MODE_BFS, max_depth=pparams.max_depth,
                         min_transitive_len=pparams.min_transitive_len,
                         min_distance_between_ranges=pparams.min_distance_between_ranges, min_output_length=None,
                         store_cigar=False, min_identity=None if mi != mi else mi, merge_distance=pparams.merge_distance,
                         merge_strands=True, masked_regions=mask)
        cols = answer(np.array([window], dtype=RANGE_DTYPE), qp)
        st.feed(cols["q_id"], cols["q_first"], cols["q_last"])
    return st.finish()


class MultiImpg:
    """Mirror of the reference's MultiImpg (src/multi_impg.rs): several alignment files behind
    one ImpgIndex. The reference keeps one sub-index per file and re-sorts the union of their
    hits by (query id, query first, query last, target first, target last); that order does not
    depend on the file partition, so here ONE HBM index holds every file and the
    IMPGX_MODE_MULTI_* modes reproduce MultiImpg's result order and traversal."""

    _MODE = {MODE_QUERY: MODE_MULTI_QUERY, MODE_BFS: MODE_MULTI_BFS, MODE_DFS: MODE_MULTI_DFS}

    def __init__(self, impg):
        self.idx = impg

    @classmethod
    def from_pafs(cls, paths, bidirectional=True, device=0):
        """MultiImpg::load_from_files over PAFs: unified ids by first appearance over the files."""
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        h = C.c_void_p()
        _check(lib().impgx_index_from_pafs(arr, C.c_size_t(len(paths)), C.c_int(1 if bidirectional else 0),
                                           C.c_int(device), C.byref(h)))
        return cls(Impg(h.value))

    @classmethod
    def from_record_sets(cls, parts, seq_lens, names=None, bidirectional=True, device=0):
        """`parts`: one (records, runs, run_offsets) per alignment file, records carrying unified ids."""
        recs = np.concatenate([np.ascontiguousarray(p[0], dtype=RECORD_DTYPE) for p in parts])
        runs = np.concatenate([np.ascontiguousarray(p[1], dtype=np.uint32) for p in parts])
        offs, base = [np.zeros(1, np.uint64)], 0
        for p in parts:
            o = np.ascontiguousarray(p[2], dtype=np.uint64)
            offs.append(o[1:] + np.uint64(base))
            base += int(o[-1])
        return cls(Impg.from_records(recs, runs, np.concatenate(offs), seq_lens, names=names,
                                     bidirectional=bidirectional, device=device))

    def _params(self, params):
        q = Params()
        C.memmove(C.byref(q), C.byref(params), C.sizeof(Params))
        for k in ("_keep", "_keep_mask"):
            if hasattr(params, k):
                setattr(q, k, getattr(params, k))
        q.mode = self._MODE.get(params.mode, params.mode)
        return q

    def query_batch(self, ranges, params):
        return self.idx.query_batch(ranges, self._params(params))

    def query_batch_bed(self, ranges, params):
        return self.idx.query_batch_bed(ranges, self._params(params))

    def query(self, target_id, range_start, range_end, store_cigar=False, min_gap_compressed_identity=None):
        """MultiImpg::query (reference src/multi_impg.rs:630-649)."""
        p = make_params(mode=MODE_MULTI_QUERY, store_cigar=store_cigar, min_identity=min_gap_compressed_identity)
        return self.idx.query_batch(np.array([(target_id, range_start, range_end)], dtype=RANGE_DTYPE), p).row_tuples(0)

    def _transitive(self, mode, target_id, range_start, range_end, **kw):
        p = make_params(mode=mode, **kw)
        return self.idx.query_batch(np.array([(target_id, range_start, range_end)], dtype=RANGE_DTYPE), p).row_tuples(0)

    def query_transitive_bfs(self, target_id, range_start, range_end, **kw):
        """MultiImpg::query_transitive_bfs (reference src/multi_impg.rs:722-755)."""
        return self._transitive(MODE_MULTI_BFS, target_id, range_start, range_end, **kw)

    def query_transitive_dfs(self, target_id, range_start, range_end, **kw):
        """MultiImpg::query_transitive_dfs (reference src/multi_impg.rs:687-720)."""
        return self._transitive(MODE_MULTI_DFS, target_id, range_start, range_end, **kw)


# ---------------------------------------------------------------- sharding
class Comm:
    """Exchange endpoint of one rank of a target-sharded index (include/impgx.h, impgx_comm)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "h", None):
            lib().impgx_comm_free(self.h)
            self.h = None

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        _check(lib().impgx_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def nccl(cls, unique_id, rank, n_ranks, device):
        """One process per GPU; `unique_id` comes from rank 0's Comm.unique_id()."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        h = C.c_void_p()
        _check(lib().impgx_comm_init_nccl(buf, C.c_int(rank), C.c_int(n_ranks), C.c_int(device), C.byref(h)))
        return cls(h.value)

    @classmethod
    def local_group(cls, n_ranks):
        """n_ranks endpoints inside this process, each to be driven by its own thread."""
        arr = (C.c_void_p * n_ranks)()
        _check(lib().impgx_comm_init_local(C.c_int(n_ranks), arr))
        return [cls(arr[r]) for r in range(n_ranks)]

    @property
    def rank(self):
        return lib().impgx_comm_rank(self.h)

    @property
    def size(self):
        return lib().impgx_comm_size(self.h)

    def traffic(self):
        s, r, e = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().impgx_comm_traffic(self.h, C.byref(s), C.byref(r), C.byref(e)))
        return {"bytes_sent": s.value, "bytes_received": r.value, "exchanges": e.value}


def assign_owners(records, run_offsets, n_seqs, n_ranks, bidirectional=True):
    """Balanced sequence -> rank map of a target-sharded index (no GPU needed)."""
    records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
    owner = np.zeros(n_seqs, np.uint32)
    _check(lib().impgx_assign_owners(_p(records), C.c_size_t(len(records)), _p(run_offsets), C.c_uint32(n_seqs),
                                     C.c_int(1 if bidirectional else 0), C.c_uint32(n_ranks), _p(owner)))
    return owner


def shard_records(records, run_offsets, owner, rank, bidirectional=True):
    """Indices of the alignments shard `rank` needs (those with an entry on an owned sequence)."""
    records = np.asarray(records)
    keep = owner[records["target_id"]] == rank
    if bidirectional:
        keep |= (owner[records["query_id"]] == rank) & (records["query_id"] != records["target_id"])
    return np.nonzero(keep)[0]


def merge_shards(parts):
    """impgx_results_merge_shards: per-rank BED rows -> the reference's per-row output."""
    arr = (C.c_void_p * len(parts))(*[p.h for p in parts])
    h = C.c_void_p()
    _check(lib().impgx_results_merge_shards(arr, C.c_int(len(parts)), C.byref(h)))
    return Results(h.value)


def merge_shard_columns(parts):
    """Same reassembly on column dicts (e.g. gathered from other processes): rows of one
    input row are ordered by sequence id, each part holds whole (row, q_id) groups."""
    n_rows = len(parts[0]["row_offsets"]) - 1
    rows = np.concatenate([np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(p["row_offsets"].astype(np.int64)))
                           for p in parts])
    cat = {k: np.concatenate([p[k] for p in parts]) for k in ("q_id", "q_first", "q_last", "t_id", "t_first", "t_last")}
    order = np.lexsort((cat["q_id"], rows))  # stable: keeps each part's order inside a (row, q_id) group
    out = {k: v[order] for k, v in cat.items()}
    ro = np.zeros(n_rows + 1, np.uint64)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=ro[1:])
    out["row_offsets"] = ro
    return out


class ShardedImpg:
    """A target-sharded index inside ONE process: shard r lives on devices[r] (devices may
    repeat: several shards on one GPU), exchanges go through the in-process transport and
    every collective call runs one host thread per shard."""

    def __init__(self, shards, comms, owner):
        self.shards, self.comms, self.owner = shards, comms, owner

    @classmethod
    def from_records(cls, records, runs, run_offsets, seq_lens, devices, names=None, bidirectional=True):
        n = len(devices)
        owner = assign_owners(records, run_offsets, len(seq_lens), n, bidirectional)
        shards = [Impg.from_records_shard(records, runs, run_offsets, seq_lens, owner, r, n, names=names,
                                          bidirectional=bidirectional, device=devices[r]) for r in range(n)]
        return cls(shards, Comm.local_group(n), owner)

    def _collective(self, fn):
        import threading
        n = len(self.shards)
        out, err = [None] * n, [None] * n

        def work(r):
            try:
                out[r] = fn(r)
            except Exception as e:  # noqa: BLE001 - re-raised below
                err[r] = e

        ts = [threading.Thread(target=work, args=(r,)) for r in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None and not (isinstance(e, ImpgxError) and "peer rank" in str(e)):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out

    def query_batch_bed_parts(self, ranges, params):
        return self._collective(lambda r: self.shards[r].query_batch_bed_sharded(self.comms[r], ranges, params))

    def query_batch_bed(self, ranges, params):
        return merge_shards(self.query_batch_bed_parts(ranges, params))

    def partition(self, pparams):
        """`impg partition -o bed` over the sharded index: every window is one collective masked walk."""
        return partition_with(self.shards[0], pparams, lambda w, qp: self.query_batch_bed(w, qp).columns())

    def stats(self):
        return [s.stats() for s in self.shards]


def _format_rows(fn, idx, results, row, name, merge_distance):
    ptr = fn(idx.h, results.h, C.c_size_t(row), name.encode(), C.c_int32(merge_distance))
    if not ptr:
        raise ImpgxError(E_INVALID, "format failed: " + lib().impgx_last_error().decode())
    s = C.string_at(ptr).decode()
    lib().impgx_free(C.c_void_p(ptr))
    return s


def format_bedpe(idx, results, row, name, merge_distance):
    """output_results_bedpe (reference src/main.rs:11894) on a raw result set with CIGARs."""
    return _format_rows(lib().impgx_format_bedpe, idx, results, row, name, merge_distance)


def format_paf(idx, results, row, name, merge_distance):
    """output_results_paf (reference src/main.rs:11989) on a raw result set with CIGARs."""
    return _format_rows(lib().impgx_format_paf, idx, results, row, name, merge_distance)


def parse_subsequence_coordinates(name):
    """parse_subsequence_coordinates (reference src/main.rs:4642-4659): (base, start) or None."""
    buf = C.create_string_buffer(len(name.encode()) + 1)
    st = C.c_int32(0)
    r = lib().impgx_parse_subsequence_coordinates(name.encode(), buf, C.c_size_t(len(buf)), C.byref(st))
    if r < 0:
        _check(r)
    return (buf.value.decode(), st.value) if r == 1 else None


def subset_matches(list_text, name):
    """SubsetFilter::matches for one sequence name (reference src/subset_filter.rs:23-42)."""
    r = lib().impgx_subset_matches(list_text.encode(), name.encode())
    if r < 0:
        _check(r)
    return bool(r)


def subset_mask(impg, list_text):
    """The per-sequence keep mask of --subset-sequence-list for make_params(subset_mask=...)."""
    m = np.zeros(impg.n_seqs, np.uint8)
    lib().impgx_subset_mask.restype = C.c_long
    r = lib().impgx_subset_mask(impg.h, list_text.encode(), _p(m))
    if r < 0:
        _check(int(r))
    return m


def parse_bed_file(path):
    """parse_bed_file (reference src/commands/partition.rs:1719-1753) -> [(seq, (start, end), name)]."""
    L = lib()
    L.impgx_bed_len.restype = C.c_size_t
    for f in ("impgx_bed_seq", "impgx_bed_name"):
        getattr(L, f).restype = C.c_char_p
    for f in ("impgx_bed_start", "impgx_bed_end"):
        getattr(L, f).restype = C.c_int32
    h = C.c_void_p()
    _check(L.impgx_bed_parse(path.encode(), C.byref(h)))
    try:
        n = L.impgx_bed_len(h)
        return [(L.impgx_bed_seq(h, C.c_size_t(i)).decode(),
                 (L.impgx_bed_start(h, C.c_size_t(i)), L.impgx_bed_end(h, C.c_size_t(i))),
                 L.impgx_bed_name(h, C.c_size_t(i)).decode()) for i in range(n)]
    finally:
        L.impgx_bed_free(h)


def parse_target_range(text):
    """parse_target_range (reference src/commands/partition.rs:1755-1768): split on the LAST ':'."""
    seq = C.create_string_buffer(4096)
    name = C.create_string_buffer(4200)
    s, e = C.c_int32(), C.c_int32()
    _check(lib().impgx_parse_target_range(text.encode(), seq, C.c_size_t(4096), C.byref(s), C.byref(e), name,
                                          C.c_size_t(4200)))
    return seq.value.decode(), (s.value, e.value), name.value.decode()


def project_batch(req, records, runs, run_offsets, device=0, want_cigar=True):
    """project_target_range_through_alignment for n independent problems on the GPU
    (reference src/impg.rs:2760-2898). records[i].reserved bit 0 marks a reversed entry."""
    n = len(records)
    records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    rs = np.ascontiguousarray([r[0] for r in req], dtype=np.int32)
    re = np.ascontiguousarray([r[1] for r in req], dtype=np.int32)
    runs = np.ascontiguousarray(runs, dtype=np.uint32)
    if len(runs) == 0:
        runs = np.zeros(1, np.uint32)
    run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
    out4 = np.zeros(4 * max(n, 1), np.int32)
    ok = np.zeros(max(n, 1), np.uint8)
    oo = np.zeros(n + 1, np.uint64)
    cap = int(run_offsets[-1]) + 1
    out_runs = np.zeros(cap, np.uint32)
    _check(lib().impgx_project_batch(C.c_int(device), C.c_size_t(n), _p(rs), _p(re), _p(records), _p(runs),
                                     _p(run_offsets), _p(out4), _p(ok), _p(oo), _p(out_runs) if want_cigar else None,
                                     C.c_size_t(cap)))
    res = []
    for i in range(n):
        if not ok[i]:
            res.append(None)
        else:
            cg = out_runs[int(oo[i]):int(oo[i + 1])].copy() if want_cigar else None
            res.append((int(out4[4 * i]), int(out4[4 * i + 1]), cg, int(out4[4 * i + 2]), int(out4[4 * i + 3])))
    return res


# ---------------------------------------------------------------- synthetic data
def synth_cfg(genomes, contigs, contig_len, tiles, eq_mean, rev_permille=100, seed=1, partners=0):
    """`partners` = k > 0: every genome is aligned (as query) against k others only (sparsified pairs)."""
    c = SynthCfg()
    c.genomes, c.contigs, c.contig_len, c.tiles = genomes, contigs, contig_len, tiles
    c.eq_mean, c.rev_permille, c.seed = eq_mean, rev_permille, seed
    c.partners = partners if 0 < partners < genomes - 1 else 0
    return c


def synth_generate(cfg):
    """gen_synth: returns (records, runs, run_offsets, seq_lens, names)."""
    L = synth_lib()
    n = L.impgx_synth_num_alignments(C.byref(cfg))
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    nr = np.zeros(n, dtype=np.uint32)
    _check_synth(L.impgx_synth_records(C.byref(cfg), C.c_uint64(0), C.c_uint64(n), _p(recs), _p(nr)))
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nr, out=offs[1:], dtype=np.uint64)
    runs = np.zeros(max(int(offs[-1]), 1), dtype=np.uint32)
    _check_synth(L.impgx_synth_runs(C.byref(cfg), C.c_uint64(0), C.c_uint64(n), _p(offs), _p(runs)))
    n_seqs = cfg.genomes * cfg.contigs
    lens = np.full(n_seqs, cfg.contig_len, dtype=np.uint64)
    names = [f"g{g}#1#c{c}" for g in range(cfg.genomes) for c in range(cfg.contigs)]
    return recs, runs[: int(offs[-1])], offs, lens, names


def synth_generate_shard(cfg, n_ranks, rank, bidirectional=True):
    """The part of gen_synth shard `rank` of `n_ranks` needs: records of every alignment are
    generated (cheap) to derive the owner map, runs only for the alignments that have an entry
    on an owned sequence. Returns (records, runs, run_offsets, seq_lens, names, owner)."""
    L = synth_lib()
    n = L.impgx_synth_num_alignments(C.byref(cfg))
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    nr = np.zeros(n, dtype=np.uint32)
    _check_synth(L.impgx_synth_records(C.byref(cfg), C.c_uint64(0), C.c_uint64(n), _p(recs), _p(nr)))
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(nr, out=offs[1:], dtype=np.uint64)
    n_seqs = cfg.genomes * cfg.contigs
    owner = assign_owners(recs, offs, n_seqs, n_ranks, bidirectional)
    keep = shard_records(recs, offs, owner, rank, bidirectional).astype(np.uint64)
    sub_offs = np.zeros(len(keep) + 1, dtype=np.uint64)
    np.cumsum(nr[keep], out=sub_offs[1:], dtype=np.uint64)
    runs = np.zeros(max(int(sub_offs[-1]), 1), dtype=np.uint32)
    _check_synth(L.impgx_synth_runs_subset(C.byref(cfg), _p(keep), C.c_uint64(len(keep)), _p(sub_offs), _p(runs)))
    lens = np.full(n_seqs, cfg.contig_len, dtype=np.uint64)
    names = [f"g{g}#1#c{c}" for g in range(cfg.genomes) for c in range(cfg.contigs)]
    return recs[keep], runs[: int(sub_offs[-1])], sub_offs, lens, names, owner


def synth_generate_contig(cfg, contig=0):
    """The alignments of ONE contig of gen_synth. The synthetic world only aligns contig c of one genome with
    contig c of another, so the trees, hits and transitive closure of a BED row on contig c are the same in
    this sub-world as in the full index. Returns (records, runs, run_offsets, seq_lens, names); sequence ids
    are those of the full world."""
    L = synth_lib()
    n = L.impgx_synth_num_alignments(C.byref(cfg))
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    nr = np.zeros(n, dtype=np.uint32)
    _check_synth(L.impgx_synth_records(C.byref(cfg), C.c_uint64(0), C.c_uint64(n), _p(recs), _p(nr)))
    pairs = cfg.genomes * (cfg.partners if cfg.partners else cfg.genomes - 1)
    keep = ((np.arange(pairs, dtype=np.uint64)[:, None] * np.uint64(cfg.contigs) + np.uint64(contig))
            * np.uint64(cfg.tiles) + np.arange(cfg.tiles, dtype=np.uint64)[None, :]).ravel()
    sub_offs = np.zeros(len(keep) + 1, dtype=np.uint64)
    np.cumsum(nr[keep], out=sub_offs[1:], dtype=np.uint64)
    runs = np.zeros(max(int(sub_offs[-1]), 1), dtype=np.uint32)
    _check_synth(L.impgx_synth_runs_subset(C.byref(cfg), _p(keep), C.c_uint64(len(keep)), _p(sub_offs), _p(runs)))
    n_seqs = cfg.genomes * cfg.contigs
    lens = np.full(n_seqs, cfg.contig_len, dtype=np.uint64)
    names = [f"g{g}#1#c{c}" for g in range(cfg.genomes) for c in range(cfg.contigs)]
    return recs[keep], runs[: int(sub_offs[-1])], sub_offs, lens, names


def synth_bed(cfg, n_rows, seed=2, min_len=1000, max_len=10000):
    out = np.zeros(n_rows, dtype=RANGE_DTYPE)
    _check_synth(synth_lib().impgx_synth_bed(C.byref(cfg), C.c_uint64(seed), C.c_uint64(n_rows), C.c_uint32(min_len),
                                 C.c_uint32(max_len), _p(out)))
    return out


def write_cigar_text(runs, run_offsets, path):
    runs = np.ascontiguousarray(runs, np.uint32)
    run_offsets = np.ascontiguousarray(run_offsets, np.uint64)
    n = len(run_offsets) - 1
    offs = np.zeros(n, np.uint64)
    lens = np.zeros(n, np.uint64)
    _check_synth(synth_lib().impgx_write_cigar_text(_p(runs), _p(run_offsets), C.c_uint64(n), path.encode(), _p(offs), _p(lens)))
    return offs, lens


def host_columns(records, run_offsets, n_seqs, bidirectional=True, owner=None, rank=0):
    """Test hook: the sorted entry columns of an index build (or of one shard), without a GPU."""
    records = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    run_offsets = np.ascontiguousarray(run_offsets, dtype=np.uint64)
    L = lib()
    args = [_p(records), C.c_size_t(len(records)), _p(run_offsets), C.c_uint32(n_seqs), C.c_int(1 if bidirectional else 0)]
    if owner is not None:
        owner = np.ascontiguousarray(owner, dtype=np.uint32)
        E = L.impgx_debug_host_columns_shard(*args, None, None, None, None, None, None, None, None, _p(owner),
                                             C.c_uint32(rank))
        if E < 0:
            raise ImpgxError(E_INVALID, L.impgx_last_error().decode())
        cols = {"e_start": np.zeros(E, np.int32), "e_end": np.zeros(E, np.int32), "e_pmax": np.zeros(E, np.int32),
                "e_vrank": np.zeros(E, np.uint32), "e_query_id": np.zeros(E, np.uint32),
                "e_flags": np.zeros(E, np.uint32), "e_aln": np.zeros(E, np.uint32),
                "tgt_off": np.zeros(n_seqs + 1, np.uint64)}
        L.impgx_debug_host_columns_shard(*args, _p(cols["e_start"]), _p(cols["e_end"]), _p(cols["e_pmax"]),
                                         _p(cols["e_vrank"]), _p(cols["e_query_id"]), _p(cols["e_flags"]),
                                         _p(cols["e_aln"]), _p(cols["tgt_off"]), _p(owner), C.c_uint32(rank))
        return cols
    E = L.impgx_debug_host_columns(*args, None, None, None, None, None, None, None, None)
    if E < 0:
        raise ImpgxError(E_INVALID, L.impgx_last_error().decode())
    cols = {"e_start": np.zeros(E, np.int32), "e_end": np.zeros(E, np.int32), "e_pmax": np.zeros(E, np.int32),
            "e_vrank": np.zeros(E, np.uint32), "e_query_id": np.zeros(E, np.uint32), "e_flags": np.zeros(E, np.uint32),
            "e_aln": np.zeros(E, np.uint32), "tgt_off": np.zeros(n_seqs + 1, np.uint64)}
    L.impgx_debug_host_columns(*args, _p(cols["e_start"]), _p(cols["e_end"]), _p(cols["e_pmax"]), _p(cols["e_vrank"]),
                               _p(cols["e_query_id"]), _p(cols["e_flags"]), _p(cols["e_aln"]), _p(cols["tgt_off"]))
    return cols
