/*
 * impgx.h — C ABI of libimpgx, the B200-native projection engine.
 *
 * This is the drop-in boundary for ONE path of pangenome/impg: batch interval
 * stabbing + per-hit CIGAR liftover + transitive BFS frontier + gap-merge,
 * i.e. what `impg query -b <BED> [-x -m N]` does through `trait ImpgIndex`
 * (reference src/impg_index.rs:21-121). Everything is `extern "C"`, plain
 * pointers and sizes; no C++ or torch types cross this boundary.
 *
 * Each entry point cites the reference interface it replaces (file:line are
 * relative to the reference checkout). INTEGRATION.md shows the Rust-side
 * binding (`GpuImpg: ImpgIndex`) a maintainer would add.
 *
 * Error model: every function returns 0 on success or a negative
 * impgx_status; impgx_last_error() returns a thread-local message. No C++
 * exception, abort or panic crosses the ABI (the reference panics at
 * src/impg.rs:88,506-511,1739 become IMPGX_E_* codes).
 *
 * There is no CPU fallback: without a usable CUDA device every compute entry
 * point fails with IMPGX_E_NO_DEVICE.
 */
#ifndef IMPGX_H
#define IMPGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMPGX_ABI_VERSION 4

typedef enum impgx_status {
  IMPGX_OK = 0,
  IMPGX_E_INVALID = -1,   /* bad argument (NULL, out-of-range id, start>=end …) */
  IMPGX_E_NO_DEVICE = -2, /* no CUDA device / driver; there is no CPU fallback  */
  IMPGX_E_CUDA = -3,      /* a CUDA runtime call or kernel failed               */
  IMPGX_E_NOMEM = -4,     /* host or device allocation failed                   */
  IMPGX_E_IO = -5,        /* file could not be read / written                   */
  IMPGX_E_PARSE = -6,     /* malformed PAF / BED / CIGAR text                   */
  IMPGX_E_UNSUPPORTED = -7 /* feature outside the path (tracepoints, approximate) */
} impgx_status;

/* CIGAR run packing, identical to the reference's CigarOp (src/impg.rs:76-140):
 * val = op << 29 | len, op in {'='=0, 'X'=1, 'I'=2, 'D'=3, 'M'=4}, len < 2^29. */
#define IMPGX_OP_EQ 0u
#define IMPGX_OP_X 1u
#define IMPGX_OP_I 2u
#define IMPGX_OP_D 3u
#define IMPGX_OP_M 4u
#define IMPGX_RUN(op, len) ((((uint32_t)(op)) << 29) | (uint32_t)(len))
#define IMPGX_RUN_OP(v) ((uint32_t)(v) >> 29)
#define IMPGX_RUN_LEN(v) ((uint32_t)(v) & 0x1fffffffu)

/* One alignment = AlignmentRecord (src/alignment_record.rs:12-21) with the
 * file offset replaced by an index into the decoded run stream. */
typedef struct impgx_record {
  uint32_t query_id;
  uint32_t target_id;
  int32_t query_start;
  int32_t query_end;
  int32_t target_start;
  int32_t target_end;
  uint32_t strand; /* 0 = '+', 1 = '-' (bit 63 of strand_and_data_offset) */
  uint32_t reserved;
} impgx_record;

/* One query range = one BED row after name→id (src/main.rs:7435, :11620). */
typedef struct impgx_range {
  uint32_t target_id;
  int32_t start;
  int32_t end;
} impgx_range;

typedef enum impgx_mode {
  IMPGX_MODE_QUERY = 0, /* Impg::query                 src/impg.rs:1852-1928 */
  IMPGX_MODE_BFS = 1,   /* Impg::query_transitive_bfs  src/impg.rs:2311-2597 */
  IMPGX_MODE_DFS = 2,   /* Impg::query_transitive_dfs  src/impg.rs:2057-2309 */
  /* MultiImpg (src/multi_impg.rs): one sub-index per alignment file, queries fanned
   * out to every sub-index that holds the target and the hits re-sorted by
   * (query id, query first, query last, target first, target last) — so the result
   * does not depend on how the alignments are spread over files, and one unified
   * HBM index (impgx_index_from_pafs) answers for all of them. */
  IMPGX_MODE_MULTI_QUERY = 3, /* MultiImpg::query                 src/multi_impg.rs:495-595, :630-649 */
  IMPGX_MODE_MULTI_BFS = 4,   /* MultiImpg::query_transitive_bfs  src/multi_impg.rs:722-755, :796-991 (queue popped at the front) */
  IMPGX_MODE_MULTI_DFS = 5    /* MultiImpg::query_transitive_dfs  src/multi_impg.rs:687-720, :796-991 (popped at the back) */
} impgx_mode;

/* Arguments of ImpgIndex::query / query_transitive_* (src/impg_index.rs:30-80)
 * plus the output-merge options of QueryOpts (src/main.rs:4319-4410). */
typedef struct impgx_params {
  uint32_t mode;                       /* impgx_mode */
  uint32_t max_depth;                  /* -m, u16 in the reference; 0 = unlimited */
  int32_t min_transitive_len;          /* --min-transitive-len (default 101) */
  int32_t min_distance_between_ranges; /* --min-distance-between-ranges (default 10) */
  int32_t min_output_length;           /* -l; < 0 means None */
  uint32_t store_cigar;                /* carry clipped CIGAR runs per result */
  double min_identity;                 /* --min-result-identity; NaN means None */
  const uint8_t *subset_mask;          /* per sequence id, 1 = keep; NULL = no filter */
  int32_t merge_distance;              /* -d; -1 = --no-merge (BED entry point only) */
  uint32_t merge_strands;              /* !--consider-strandness (BED entry point only) */
  /* masked_regions of the transitive queries (src/impg.rs:2331-2340; partition passes the
   * regions already assigned, src/commands/partition.rs:254-270, :359-391) as CSR over ALL
   * sequences: sequence s owns the sorted, disjoint [start, end) pairs
   * mask_ranges[2*mask_offsets[s] .. 2*mask_offsets[s+1]). The visited set of every row starts
   * from it, so only the unmasked pieces of a row's range are output and walked. NULL = None.
   * (A sequence missing from the reference's map gets length 0 there, src/impg.rs:2047-2053;
   * the CSR form always carries every sequence, as partition builds its map.) HOST pointers. */
  const uint64_t *mask_offsets;        /* n_seqs + 1, or NULL */
  const int32_t *mask_ranges;          /* 2 * mask_offsets[n_seqs] */
} impgx_params;

typedef struct impgx_index impgx_index;
typedef struct impgx_results impgx_results;

/* Column view of a result set (AdjustedInterval = (query Interval, cigar,
 * target Interval), src/impg.rs:225). Row r of the batch owns result slots
 * [row_offsets[r], row_offsets[r+1]); within a row the order is the
 * reference's result order (self interval first). For reverse-strand hits
 * q_first > q_last, exactly as the reference reports them. Pointers are owned
 * by the impgx_results object and live until impgx_results_free. */
typedef struct impgx_view {
  size_t n_rows;
  size_t n_results;
  const uint64_t *row_offsets; /* n_rows + 1 */
  const uint32_t *q_id;
  const int32_t *q_first;
  const int32_t *q_last;
  const uint32_t *t_id;
  const int32_t *t_first;
  const int32_t *t_last;
  const uint64_t *cigar_offsets; /* n_results + 1, or NULL without store_cigar */
  const uint32_t *cigar_runs;    /* reference packing, or NULL */
} impgx_view;

/* Counters of the last call on an index (for bench.py's roofline line). */
typedef struct impgx_stats {
  uint64_t kernel_launches; /* kernels of this library launched by the call */
  uint64_t stab_ranges;     /* ranges stabbed over all hops */
  uint64_t lift_launches;   /* liftover kernel launches of the call */
  uint64_t liftovers;       /* hits lifted through a CIGAR */
  uint64_t lift_runs;       /* CIGAR runs read by the liftover kernel */
  uint64_t lift_bytes;      /* algorithmic bytes of the liftover (SURVEY.md 8d formula) */
  uint64_t lift_touched_bytes; /* bytes the kernel's own algorithm touches (entry, checkpoints, blocks read, hit) */
  uint64_t lift_window_runs;   /* sum over hits of r_ov = runs intersecting the request */
  uint64_t results;         /* result intervals before merging */
  uint64_t merged;          /* result intervals after merging (BED entry point) */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  float lift_ms;            /* device time of the liftover launches (CUDA events) */
  float stab_ms;
  float fold_ms;
  float merge_ms;
  float total_ms;
  float exchange_ms;        /* sharded index: device time of the hit / frontier / box exchanges (CUDA events) */
  float merge_kernel_ms;    /* device time of the segment-merge launches alone (inside merge_ms) */
  float reserved0;
  uint64_t merge_boxes;     /* boxes (results before merging) that went through the BED merge */
  uint64_t exchange_bytes;  /* sharded index: bytes this rank sent in the exchanges of the call */
} impgx_stats;

int impgx_abi_version(void);
const char *impgx_last_error(void);
int impgx_device_count(void);

/* Replaces Impg::from_multi_alignment_records (src/impg.rs:1535-1652).
 * `runs`/`run_offsets` hold the decoded CIGAR of every record (record i owns
 * runs[run_offsets[i] .. run_offsets[i+1])), in PAF order; the library never
 * reads PAF text at query time (the reference's per-hit pread+parse,
 * src/impg.rs:495-552, is done once here). Inputs are borrowed for the call.
 * `device` is the CUDA ordinal that will hold the index. */
int impgx_index_build(const impgx_record *records, size_t n_records,
                      const uint32_t *runs, const uint64_t *run_offsets,
                      const uint64_t *seq_lens, uint32_t n_seqs,
                      int bidirectional, int device, impgx_index **out);

/* Convenience: parse_paf_file (src/paf.rs:118-194,306-362) + build. Sequence
 * ids are assigned by first appearance in the file (query column first). */
int impgx_index_from_paf(const char *paf_path, int bidirectional, int device,
                         impgx_index **out);

/* MultiImpg::load_from_files (src/multi_impg.rs:140-216) over PAF files: unified
 * sequence ids by first appearance over the files in the given order; one HBM
 * index holds every file's alignments (query it with the IMPGX_MODE_MULTI_* modes
 * for MultiImpg's result order, or with the plain modes for Impg's). */
int impgx_index_from_pafs(const char *const *paf_paths, size_t n_paths, int bidirectional, int device,
                          impgx_index **out);

void impgx_index_free(impgx_index *idx);

/* SequenceIndex accessors (src/seqidx.rs:22-56). */
uint32_t impgx_index_num_seqs(const impgx_index *idx);
uint64_t impgx_index_num_entries(const impgx_index *idx);
uint64_t impgx_index_device_bytes(const impgx_index *idx);
const char *impgx_index_seq_name(const impgx_index *idx, uint32_t id); /* NULL if unnamed */
uint64_t impgx_index_seq_len(const impgx_index *idx, uint32_t id);
int impgx_index_seq_id(const impgx_index *idx, const char *name, uint32_t *id_out);
int impgx_index_set_names(impgx_index *idx, const char *const *names, uint32_t n);

/* The batch entry: replaces the serial BED loop + perform_query
 * (src/main.rs:7435-7456, :11605-11707). `ranges` is a HOST array. Results
 * are every AdjustedInterval of every row in the reference's order. */
int impgx_query_batch(impgx_index *idx, const impgx_range *ranges, size_t n,
                      const impgx_params *params, impgx_results **out);

/* Same, followed on the device by output_results_bed's two merges
 * (merge_adjusted_intervals_gap_2d then merge_query_adjusted_intervals,
 * src/main.rs:11849-11866, :12858-13011, :12474-12560). Result rows are the
 * BED rows the reference would print (q columns; t columns are the merged
 * boxes' targets and are not part of BED). */
int impgx_query_batch_bed(impgx_index *idx, const impgx_range *ranges, size_t n,
                          const impgx_params *params, impgx_results **out);

/* Device-resident variants: `d_ranges` is a DEVICE pointer to n ranges that
 * are already in HBM; results stay in HBM (impgx_results_device_view). Used to
 * time the path with no host<->device copy in the timed region. `stream` is a
 * cudaStream_t (NULL = default stream). */
int impgx_query_batch_bed_device(impgx_index *idx, const impgx_range *d_ranges,
                                 size_t n, const impgx_params *params,
                                 void *stream, impgx_results **out);

/* ---- target-sharded index (replaces MultiImpg's role of
 * spreading one logical index over several holders, src/multi_impg.rs:495-595).
 * The index is partitioned by target sequence: rank r holds the interval
 * columns, run stream and visited sets of the sequences with owner[seq] == r.
 * A query batch is a COLLECTIVE call: every rank passes the same ranges and
 * params; between transitive hops the ranks exchange lifted hits (all-to-all-v,
 * to the owner of the sequence a hit lands on) and the next frontier
 * (all-gather-v) over an impgx_comm. Rank r returns the BED rows whose
 * sequence it owns; impgx_results_merge_shards reassembles the reference's
 * per-row output (rows of one input row are ordered by sequence id). */
typedef struct impgx_comm impgx_comm;
#define IMPGX_COMM_ID_BYTES 128
/* NCCL transport, one process per GPU: rank 0 creates the id (ncclGetUniqueId)
 * and hands it to the other ranks by any out-of-band means. */
int impgx_comm_unique_id(uint8_t id[IMPGX_COMM_ID_BYTES]);
int impgx_comm_init_nccl(const uint8_t id[IMPGX_COMM_ID_BYTES], int rank, int n_ranks, int device,
                         impgx_comm **out);
/* In-process transport: n_ranks endpoints (out[0..n_ranks)), each driven by its
 * own host thread; the shards may sit on different devices or share one. After
 * a failed collective call the group is unusable and must be recreated. */
int impgx_comm_init_local(int n_ranks, impgx_comm **out);
int impgx_comm_rank(const impgx_comm *c);
int impgx_comm_size(const impgx_comm *c);
/* bytes this rank sent / received in exchanges so far, number of exchange steps */
int impgx_comm_traffic(const impgx_comm *c, uint64_t *sent, uint64_t *received, uint64_t *exchanges);
void impgx_comm_free(impgx_comm *c);

/* Balanced sequence -> rank map (weight = entry bytes + run-stream bytes of the
 * sequence's entries; deterministic, so every rank computes the same map). */
int impgx_assign_owners(const impgx_record *records, size_t n_records, const uint64_t *run_offsets,
                        uint32_t n_seqs, int bidirectional, uint32_t n_ranks, uint32_t *owner_out);
/* As impgx_index_build, keeping only the entries of the sequences owned by
 * `rank` (and the runs of the alignments those entries walk). `records` may be
 * the full list or any subset that contains every alignment touching an owned
 * sequence, in the original relative order. */
int impgx_index_build_shard(const impgx_record *records, size_t n_records, const uint32_t *runs,
                            const uint64_t *run_offsets, const uint64_t *seq_lens, uint32_t n_seqs,
                            int bidirectional, int device, const uint32_t *owner, uint32_t rank,
                            uint32_t n_ranks, impgx_index **out);
/* Collective counterparts of impgx_query_batch_bed / _bed_device. Modes QUERY,
 * BFS and DFS; the MultiImpg walks and raw (unmerged) results need an unsharded index. */
int impgx_query_batch_bed_sharded(impgx_index *shard, impgx_comm *comm, const impgx_range *ranges, size_t n,
                                  const impgx_params *params, impgx_results **out);
int impgx_query_batch_bed_sharded_device(impgx_index *shard, impgx_comm *comm, const impgx_range *d_ranges,
                                         size_t n, const impgx_params *params, void *stream,
                                         impgx_results **out);
/* Host-side reassembly of the per-rank BED rows into one result set. */
int impgx_results_merge_shards(const impgx_results *const *parts, int n_parts, impgx_results **out);

int impgx_results_view(const impgx_results *res, impgx_view *view);        /* host columns */
int impgx_results_device_view(const impgx_results *res, impgx_view *view); /* device columns */
void impgx_results_free(impgx_results *res);

int impgx_index_stats(const impgx_index *idx, impgx_stats *out);

/* Single-hit liftover, the KAT surface of project_target_range_through_alignment
 * (src/impg.rs:2760-2898): runs the liftover kernel on n independent
 * (request, record, runs) problems. out4 = {q_start,q_end,t_start,t_end} per
 * problem, ok[i] = 1 iff the reference returns Some. With out_runs != NULL the
 * clipped run slices are returned (out_run_offsets has n+1 entries). */
int impgx_project_batch(int device, size_t n, const int32_t *req_start,
                        const int32_t *req_end, const impgx_record *records,
                        const uint32_t *runs, const uint64_t *run_offsets,
                        int32_t *out4, uint8_t *ok, uint64_t *out_run_offsets,
                        uint32_t *out_runs, size_t out_runs_cap);

/* BED / range parsing of the query driver (src/commands/partition.rs:1719-1789):
 * >= 3 tab-separated fields, start < end, row name = column 4 unless empty or
 * ".", else "chrom:start-end"; `-r` text is split on the LAST ':'. */
typedef struct impgx_bed impgx_bed;
int impgx_bed_parse(const char *path, impgx_bed **out);
size_t impgx_bed_len(const impgx_bed *bed);
const char *impgx_bed_seq(const impgx_bed *bed, size_t i);
const char *impgx_bed_name(const impgx_bed *bed, size_t i);
int32_t impgx_bed_start(const impgx_bed *bed, size_t i);
int32_t impgx_bed_end(const impgx_bed *bed, size_t i);
void impgx_bed_free(impgx_bed *bed);
int impgx_parse_target_range(const char *text, char *seq_out, size_t seq_cap,
                             int32_t *start, int32_t *end, char *name_out, size_t name_cap);

/* --subset-sequence-list (src/subset_filter.rs:17-178): `list_text` = the contents of the list file (one
 * name per line, '#' comments). A sequence is kept if the list holds its name, its name without ":coords",
 * or its sample / sample+haplotype ("<sample>_hap<N>…" or PanSN "<sample>#<N>#…"). Fills mask_out[n_seqs]
 * for impgx_params.subset_mask; returns the number of list entries (an empty list is IMPGX_E_PARSE, :72-77). */
long impgx_subset_mask(const impgx_index *idx, const char *list_text, uint8_t *mask_out);
/* SubsetFilter::matches on one name (1 / 0): the surface of the reference's own test (:185-206). */
int impgx_subset_matches(const char *list_text, const char *seq_name);

/* --original-sequence-coordinates (transform_coordinates_to_original, src/main.rs:4642-4678): when on, the BED and
 * BEDPE writers report a sequence named "base:start-end" as `base` with `start` added to its coordinates; the
 * PAF writer then fails (it needs the original sequences' lengths from the sequence files). Default off. */
int impgx_index_set_original_coordinates(impgx_index *idx, int on);
/* parse_subsequence_coordinates (src/main.rs:4642-4659): 1 and (base name, start) for "base:start-end", else 0. */
int impgx_parse_subsequence_coordinates(const char *seq_name, char *base_out, size_t base_cap, int32_t *start_out);
/* parse_merge_distance (src/main.rs:47-55): "50000", "50k", "1.5k", "1m", "1M"; rejects "10kb", "3g" (> i32). */
int impgx_parse_merge_distance(const char *text, int32_t *out);

/* Host-side text helpers mirroring the reference's parsers and writers. */
/* parse_cigar_to_delta (src/impg.rs:2935-2950). Returns run count or <0. */
long impgx_parse_cigar(const char *text, size_t len, uint32_t *out, size_t cap);
/* output_results_bed line formatting (src/main.rs:11867-11889) for row `row`
 * of a merged result set. Returns a malloc'ed string (caller frees with
 * impgx_free) or NULL. */
char *impgx_format_bed(const impgx_index *idx, const impgx_results *res,
                       size_t row, const char *name);
/* The same lines for EVERY row of the batch (names[r] = region name of input row r), in input
 * order — what `impg query -b` prints for the whole BED file; formatted on all host cores.
 * *len_out = bytes without the terminating NUL. malloc'ed, impgx_free. */
char *impgx_format_bed_batch(const impgx_index *idx, const impgx_results *res, const char *const *names,
                             size_t *len_out);
/* output_results_bedpe / output_results_paf (src/main.rs:11894-12103) for row
 * `row` of a RAW result set obtained with store_cigar = 1: drops the self
 * interval (src/main.rs:7474,7486), runs merge_adjusted_intervals
 * (src/main.rs:12563-12845, host side: f32 CIGAR surgery) and formats the
 * lines. malloc'ed string, free with impgx_free. */
char *impgx_format_bedpe(const impgx_index *idx, const impgx_results *res,
                         size_t row, const char *name, int32_t merge_distance);
char *impgx_format_paf(const impgx_index *idx, const impgx_results *res,
                       size_t row, const char *name, int32_t merge_distance);
void impgx_free(void *p);

/* ---- `.impg` index files (Impg::serialize_with_forest_map / load_from_file,
 * src/impg.rs:1655-1850; bincode 2 standard config, see csrc/impg_file.cu). An index file holds the
 * sequence index and, per target, the interval entries with (alignment file index, byte offset,
 * byte length) of their CIGAR text — no CIGARs — so opening one needs the alignment files it was
 * built from (uncompressed PAF). Byte parity with stock impg is unpinned: no fixture, bincode not
 * vendored. */
typedef struct impgx_impg impgx_impg; /* a parsed index file, host memory only */
int impgx_impg_open(const char *path, impgx_impg **out);
void impgx_impg_close(impgx_impg *f);
int impgx_impg_version(const impgx_impg *f);       /* 2 = "IMPGIDX2", 1 = legacy "IMPGIDX1" */
int impgx_impg_bidirectional(const impgx_impg *f); /* reversed entries present (src/impg.rs:1588-1605) */
uint32_t impgx_impg_num_seqs(const impgx_impg *f);
const char *impgx_impg_seq_name(const impgx_impg *f, uint32_t id);
uint64_t impgx_impg_seq_len(const impgx_impg *f, uint32_t id);
uint64_t impgx_impg_num_entries(const impgx_impg *f); /* tree intervals, both directions */
uint64_t impgx_impg_num_records(const impgx_impg *f); /* alignments = entries without the REVERSED bit */
/* the alignments in alignment-file order (file index, then byte offset); any output may be NULL */
int impgx_impg_records(const impgx_impg *f, impgx_record *records, uint32_t *file_index, uint64_t *data_offset,
                       uint64_t *data_bytes);
/* Impg::load_from_file + the CIGARs decoded once from alignment_files[alignment_file_index]: the HBM
 * index with the sequence ids of the file (so results are in stock impg's id space). */
int impgx_index_from_impg(const char *impg_path, const char *const *alignment_files, size_t n_files, int device,
                          impgx_index **out);
/* parse the PAF files (ids by first appearance over the files in order) and write the index file
 * stock impg would load for them (`impg index`); host only, no device needed. */
int impgx_impg_write(const char *const *paf_paths, size_t n_paths, int bidirectional, const char *out_path);

/* ---- partition: the second driver of the same kernels ----
 * partition_alignments (src/commands/partition.rs:158-712) for `-o bed`: windows are
 * taken from the sequences still missing, each window runs ONE transitive query with
 * the regions already assigned as masked_regions and store_cigar = false
 * (:359-391); its query intervals are merged (merge_overlaps :939-976), pulled to close
 * sequence ends (extend_to_close_boundaries :1369-1408), cut by the mask which they then
 * join (mask_and_update_regions :978-1366), merged again with distance 0 and become one
 * partition. Windows are inherently sequential (each depends on the mask the previous one
 * left), so the device runs one masked BFS + BED merge per window; the bookkeeping between
 * windows is host code. Only min/max of an interval reach `partitions.bed` (:1509-1542,
 * :1682-1717), so intervals are reported as start < end. */
typedef struct impgx_partition_params {
  uint64_t window_size;                /* -w */
  const uint32_t *starting_seqs;       /* --starting-sequences-file after name->id (:184-247), file order; NULL = none */
  size_t n_starting_seqs;
  const char *selection_mode;          /* "longest" (default when NULL) | "total" | "sample[,sep]" | "haplotype[,sep]" (:715-937) */
  int32_t merge_distance;              /* -d; < 0 = --no-merge */
  int32_t min_missing_size;            /* --min-missing-size (default 3000) */
  int32_t min_boundary_distance;       /* --min-boundary-distance (default 3000) */
  uint32_t transitive_dfs;             /* --transitive-dfs */
  uint32_t max_depth;                  /* -m (default 2; 0 = unlimited) */
  int32_t min_transitive_len;          /* default 101 */
  int32_t min_distance_between_ranges; /* default 10 */
  uint32_t rehome_singletons;          /* !--no-rehome-singletons (rehome_singleton_slivers :45-156) */
  double min_identity;                 /* --min-result-identity; NaN = None */
  uint32_t multi_impg;                 /* the windows use MultiImpg's transitive walk (src/multi_impg.rs:796-991; the
                                          reference's index is a MultiImpg with --index-mode per-file or >= 100 files) */
  uint32_t reserved;
} impgx_partition_params;

typedef struct impgx_partitions impgx_partitions;
typedef struct impgx_partition_view {
  size_t n_intervals;            /* rows of partitions.bed */
  size_t n_partitions;           /* partitions computed (partition_num of the next one) */
  uint64_t n_windows;            /* windows queried */
  uint64_t partitioned_bp;       /* total_partitioned_length (:447) */
  uint64_t total_bp;             /* total_sequence_length (:271-274) */
  const uint32_t *partition_num; /* per interval, non-decreasing */
  const uint32_t *seq_id;
  const int32_t *start;
  const int32_t *end;
} impgx_partition_view;

/* The whole window loop on an index resident in HBM (every window = one masked transitive
 * query + BED merge on the device). Sequence names (impgx_index_set_names / the PAF) are
 * needed for the sample / haplotype selection modes only. */
int impgx_partition(impgx_index *idx, const impgx_partition_params *params, impgx_partitions **out);
int impgx_partitions_view(const impgx_partitions *parts, impgx_partition_view *view);
/* write_single_partition_file (:1682-1717; partition < 0: "name\tstart\tend\tpartition_num")
 * or write_partition_bed of one partition (:1509-1542; "name\tstart\tend"). malloc'ed, impgx_free. */
char *impgx_partitions_format_bed(const impgx_index *idx, const impgx_partitions *parts, int64_t partition);
void impgx_partitions_free(impgx_partitions *parts);

/* The same loop as a stepper over ANY ImpgIndex implementor (what partition_alignments is
 * generic over, :159): next() hands out the next window together with the current
 * masked_regions (CSR over all sequences, valid until the next call on the object), the
 * caller runs query_transitive_bfs/dfs on it and feeds the query intervals back. */
typedef struct impgx_partitioner impgx_partitioner;
int impgx_partitioner_new(const uint64_t *seq_lens, const char *const *names /* NULL unless sample/haplotype */,
                          uint32_t n_seqs, const impgx_partition_params *params, impgx_partitioner **out);
/* returns 1 with *window filled, 0 when no window is left, < 0 on error */
int impgx_partitioner_next(impgx_partitioner *p, impgx_range *window, const uint64_t **mask_offsets,
                           const int32_t **mask_ranges);
/* overlaps of the window in the order the query returned them (first > last on the reverse strand) */
int impgx_partitioner_feed(impgx_partitioner *p, size_t n, const uint32_t *q_id, const int32_t *q_first,
                           const int32_t *q_last);
/* rehomes singletons if asked and returns the collected partitions; the stepper stays valid */
int impgx_partitioner_finish(impgx_partitioner *p, impgx_partitions **out);
void impgx_partitioner_free(impgx_partitioner *p);

/* ---------------------------------------------------------------- refine
 * Replaces Impg::populate_cigar_cache / Impg::query_with_cache (src/impg.rs:1930-2035; trait
 * src/impg_index.rs) and their caller, refine's flank search (src/commands/refine.rs:144-545).
 *
 * The reference decodes the CIGAR of every alignment under the widest candidate interval once
 * (populate_cigar_cache) so that the many overlapping candidate queries of one locus do not re-read and
 * re-parse them (query_with_cache). Here every CIGAR has been resident in HBM since the index was built:
 * there is nothing to populate, and query_with_cache is Impg::query. What the library adds is the batch:
 * the candidates of a whole phase of the flank search — for every locus of the call — are ONE batch of rows. */

/* What populate_cigar_cache(target_id, start, end) would cache: *n_keys = number of distinct alignments
 * (cache keys: alignment file index + data offset) whose tree entries the closed stab of [start, end] visits. */
int impgx_populate_cigar_cache(impgx_index *idx, uint32_t target_id, int32_t start, int32_t end, uint64_t *n_keys);

/* query_with_cache for a grid of candidates of one locus (evaluate_candidate, refine.rs:412-479): candidate k is
 * [max(0, orig_start - left[k]), min(seq_len, orig_end + right[k])) on target_id; candidates that are empty after
 * clamping get zero results (the reference skips them). Results per candidate as Impg::query returns them (self
 * interval first, reference order); params->mode must be IMPGX_MODE_QUERY. */
int impgx_query_with_cache_batch(impgx_index *idx, uint32_t target_id, int32_t orig_start, int32_t orig_end,
                                 const int32_t *left, const int32_t *right, size_t n_candidates,
                                 const impgx_params *params, impgx_results **out);

typedef struct impgx_refine_params {
  int32_t span_bp;                     /* --span-bp (default 1000) */
  int32_t extension_step;              /* --extension-step (default 1000), > 0 */
  double max_extension;                /* --max-extension (default 0.5): <= 1 a fraction of the locus, > 1 bp */
  uint32_t support_level;              /* --pansn-mode: 0 sequence (default), 1 sample, 2 haplotype */
  int32_t merge_distance;              /* -d; < 0 = --no-merge */
  double min_identity;                 /* --min-result-identity; NaN = None */
  uint32_t transitive;                 /* 0: query_with_cache, 1: -x (BFS), 2: --transitive-dfs */
  uint32_t max_depth;                  /* -m */
  int32_t min_transitive_len;
  int32_t min_distance_between_ranges;
  const uint8_t *subset_mask;          /* --subset-sequence-list as a per-sequence mask (impgx_subset_mask) or NULL */
  const uint64_t *blacklist_offsets;   /* --blacklist-bed after name->id: CSR over all sequences (n_seqs + 1) or NULL */
  const int32_t *blacklist_ranges;     /* (start, end) pairs exactly as the BED gives them */
} impgx_refine_params;

typedef struct impgx_refine_results impgx_refine_results;
typedef struct impgx_refine_view {      /* RefineRecord, refine.rs:41-54, one per locus, in input order */
  size_t n;
  const int32_t *refined_start, *refined_end, *original_start, *original_end;
  const int32_t *applied_left_extension, *applied_right_extension;
  const uint64_t *support_count, *original_support_count;
  const uint64_t *entity_offsets;       /* n + 1: support_entities of locus i = entities [off[i], off[i+1]) */
  const uint32_t *entity_seq;           /* sorted by (sequence name, start) */
  const int32_t *entity_start, *entity_end;
  uint64_t candidates_evaluated;        /* candidate ranges queried */
  uint64_t batches;                     /* device batches they were answered in */
} impgx_refine_view;

/* run_refine (refine.rs:81-142) over `n` loci. Errors like the reference: end <= start, unknown target id, or no
 * candidate of a locus evaluable -> IMPGX_E_INVALID. */
int impgx_refine(impgx_index *idx, const impgx_range *loci, size_t n, const impgx_refine_params *params,
                 impgx_refine_results **out);
int impgx_refine_view_get(const impgx_refine_results *res, impgx_refine_view *view);
void impgx_refine_results_free(impgx_refine_results *res);

#ifdef __cplusplus
}
#endif
#endif /* IMPGX_H */
