// index.cuh — the HBM-resident columnar index that replaces the reference's
// per-target coitrees (Impg.trees, src/impg.rs:226,394-404) and its on-disk
// CIGAR text (src/impg.rs:495-552).
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"

namespace impgx {

// Host-side staging of the entry columns (also used by CPU-only tests).
struct HostColumns {
  std::vector<int32_t> e_start, e_end, e_pmax;
  std::vector<uint32_t> e_vrank;   // also stored in e_rec
  std::vector<uint32_t> e_aln;     // alignment ordinal of each entry (host only, tests)
  std::vector<uint32_t> e_qid;     // query sequence of each entry (also stored in e_rec)
  std::vector<EntryRec> e_rec;
  std::vector<uint64_t> tgt_off;   // n_seqs + 1
  std::vector<uint32_t> aln_off;   // n_records + 1, stream offset of each alignment in 32-byte sectors
};

// coitrees 0.4.0 visit order (un-vendored dependency; restated from its
// published algorithm, see DESIGN.md "visit order"): rank[i] = position of
// sorted entry i in the order BasicCOITree::query would visit it if every
// entry overlapped. A stab's hits sorted by rank are in the reference's order.
constexpr size_t SIMPLE_SUBTREE_CUTOFF = 8;
void visit_ranks(size_t n, uint32_t *rank);

// Entry construction + ordering of Impg::from_multi_alignment_records
// (src/impg.rs:1535-1652): forward entry under target_id, reversed entry under
// query_id (skipped for self alignments), per-target order = record order,
// then a stable sort by start (coitrees' radix sort on `first`).
// With `owner` != nullptr only the entries of the targets owned by `rank` are
// built (a target-sharded index); the visit ranks are unchanged
// because a target's entries are never split.
void build_host_columns(const impgx_record *recs, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                        bool bidirectional, HostColumns &out, const uint32_t *owner = nullptr, uint32_t rank = 0);
void assign_owners(const impgx_record *recs, size_t n, const uint64_t *run_offsets, uint32_t n_seqs, bool bidirectional,
                   uint32_t n_ranks, uint32_t *owner);

struct Stats {
  impgx_stats s{};
};

}  // namespace impgx

struct impgx_index {
  int device = 0;
  uint32_t n_seqs = 0;
  uint64_t n_entries = 0, n_records = 0, n_blocks = 0;
  uint64_t device_bytes = 0;
  std::vector<uint64_t> seq_lens;
  std::vector<std::string> names;
  std::unordered_map<std::string, uint32_t> name_to_id;
  // device columns
  int32_t *d_start = nullptr, *d_end = nullptr, *d_pmax = nullptr, *d_seq_len = nullptr;
  uint32_t *d_stream = nullptr, *d_qid = nullptr;
  impgx::EntryRec *d_rec = nullptr;
  uint64_t *d_tgt_off = nullptr;
  // The handle is Send + Sync like the reference's ImpgIndex (src/impg_index.rs:21; refine calls it from rayon
  // workers, src/commands/refine.rs:525): concurrent calls each lease their own scratch arena (and, for the
  // host entry points, their own stream) from these pools, so they overlap on the device. `mu` guards the
  // pools, `last` and `hits_per_row`; it is never held while a query runs.
  std::mutex mu;
  std::vector<std::unique_ptr<impgx::Arena>> arena_pool;
  std::vector<cudaStream_t> stream_pool;
  impgx_stats last{};
  double hits_per_row = 0;  // observed liftovers per row (sizes the row batches)
  // single-launch path for calls of a few rows (small_bfs.cuh): calls still to skip after a row did not fit, and
  // the current back-off (doubles per overflow, cleared by a call that fits)
  uint32_t small_skip = 0, small_penalty = 0;
  bool original_coordinates = false;  // writers: --original-sequence-coordinates (src/main.rs:4661-4678)
  // target-sharded index: this object holds the entries of the
  // sequences with owner[seq] == shard_rank; empty owner = the whole index
  std::vector<uint32_t> owner;
  uint32_t *d_owner = nullptr;
  uint32_t shard_rank = 0, shard_size = 1;
  // the sequences ordered by (owner, id) and the first position of every rank in that order (n_ranks + 1)
  std::vector<uint32_t> q_order, q_first;
  uint32_t *d_qorder = nullptr;

  impgx::DevIndexView view() const {
    impgx::DevIndexView v;
    v.e_start = d_start; v.e_end = d_end; v.e_pmax = d_pmax; v.e_rec = d_rec; v.e_qid = d_qid;
    v.tgt_off = d_tgt_off; v.seq_len = d_seq_len; v.stream = d_stream;
    v.n_seqs = n_seqs; v.n_entries = n_entries;
    return v;
  }
  ~impgx_index();
};
