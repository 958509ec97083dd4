// index_host.cu — host side of the index build (no device code in this file).
#include <algorithm>
#include <numeric>

#include "index.cuh"

namespace impgx {

namespace {

int floor_log2(size_t n) {
  int k = -1;
  while (n) {
    n >>= 1;
    k++;
  }
  return k;
}

// The search tree over the sorted array [s,e) has root s+(e-s)/2. During the
// van-Emde-Boas layout pass a "childless" bottom subtree of at most
// SIMPLE_SUBTREE_CUTOFF nodes is stored as a sorted run (visited in sorted
// order); every other node is visited root, left, right. chunk_depth is the
// depth at which the next bottom subtree starts (pivot depth + 1).
void walk(size_t s, size_t e, int depth, int chunk_depth, uint32_t *rank, uint32_t &ctr) {
  if (s >= e) return;
  const size_t n = e - s;
  if (depth == chunk_depth) {
    if (n <= SIMPLE_SUBTREE_CUTOFF) {
      for (size_t i = s; i < e; i++) rank[i] = ctr++;
      return;
    }
    const int max_depth = depth + floor_log2(n);
    const int pivot = depth + (max_depth - depth) / 2;
    chunk_depth = pivot + 1;
  }
  const size_t root = s + n / 2;
  rank[root] = ctr++;
  walk(s, root, depth + 1, chunk_depth, rank, ctr);
  walk(root + 1, e, depth + 1, chunk_depth, rank, ctr);
}

}  // namespace

void visit_ranks(size_t n, uint32_t *rank) {
  uint32_t ctr = 0;
  walk(0, n, 0, 0, rank, ctr);
}

// Balanced assignment of target sequences to ranks: weight of
// a sequence = bytes of its entry columns + the run-stream regions its entries
// walk; longest-processing-time greedy, ties by id, so every rank computes the
// same map from the same records.
void assign_owners(const impgx_record *recs, size_t n, const uint64_t *run_offsets, uint32_t n_seqs, bool bidirectional,
                   uint32_t n_ranks, uint32_t *owner) {
  REQUIRE(n_ranks >= 1, IMPGX_E_INVALID, "n_ranks must be >= 1");
  std::vector<uint64_t> w(n_seqs, 0);
  for (size_t i = 0; i < n; i++) {
    const impgx_record &r = recs[i];
    REQUIRE(r.query_id < n_seqs && r.target_id < n_seqs, IMPGX_E_INVALID, "record references an unknown sequence id");
    const uint64_t b = 44 + (uint64_t)aln_sectors((uint32_t)(run_offsets[i + 1] - run_offsets[i])) * 32;
    w[r.target_id] += b;
    if (bidirectional && r.query_id != r.target_id) w[r.query_id] += b;
  }
  std::vector<uint32_t> order(n_seqs);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return w[a] > w[b]; });
  std::vector<uint64_t> load(n_ranks, 0);
  for (uint32_t s : order) {
    uint32_t best = 0;
    for (uint32_t r = 1; r < n_ranks; r++)
      if (load[r] < load[best]) best = r;
    owner[s] = best;
    load[best] += w[s] + 1;  // +1: sequences without alignments are spread too
  }
}

void build_host_columns(const impgx_record *recs, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                        bool bidirectional, HostColumns &out, const uint32_t *owner, uint32_t rank) {
  auto owned = [&](uint32_t seq) { return !owner || owner[seq] == rank; };
  for (size_t i = 0; i < n; i++) {
    const impgx_record &r = recs[i];
    REQUIRE(r.query_id < n_seqs && r.target_id < n_seqs, IMPGX_E_INVALID, "record references an unknown sequence id");
    REQUIRE(r.strand <= 1, IMPGX_E_INVALID, "record strand must be 0 or 1");
  }
  // stream offset (32-byte sectors) of each alignment this shard walks
  out.aln_off.resize(n + 1);
  uint64_t sectors = 0;
  for (size_t i = 0; i < n; i++) {
    out.aln_off[i] = (uint32_t)sectors;
    const impgx_record &r = recs[i];
    const bool needed = owned(r.target_id) || (bidirectional && r.query_id != r.target_id && owned(r.query_id));
    if (!needed) continue;
    uint64_t nr = run_offsets[i + 1] - run_offsets[i];
    REQUIRE(nr < (1ull << 30), IMPGX_E_INVALID, "alignment with >= 2^30 CIGAR runs");
    sectors += aln_sectors((uint32_t)nr);
    REQUIRE(sectors < (1ull << 32), IMPGX_E_INVALID, "run stream exceeds 2^32 sectors (137 GB); shard the index");
  }
  out.aln_off[n] = (uint32_t)sectors;

  // count entries per target
  std::vector<uint64_t> cnt(n_seqs + 1, 0);
  for (size_t i = 0; i < n; i++) {
    const impgx_record &r = recs[i];
    if (owned(r.target_id)) cnt[r.target_id + 1]++;
    if (bidirectional && r.query_id != r.target_id && owned(r.query_id)) cnt[r.query_id + 1]++;
  }
  out.tgt_off.assign(n_seqs + 1, 0);
  for (uint32_t s = 0; s < n_seqs; s++) out.tgt_off[s + 1] = out.tgt_off[s] + cnt[s + 1];
  const uint64_t E = out.tgt_off[n_seqs];

  // bucket (record order within a target is preserved)
  struct Tmp {
    int32_t start;
    uint32_t rec;  // record ordinal << 1 | reversed
  };
  std::vector<Tmp> tmp(E);
  {
    std::vector<uint64_t> cur(out.tgt_off.begin(), out.tgt_off.end() - 1);
    for (size_t i = 0; i < n; i++) {
      const impgx_record &r = recs[i];
      if (owned(r.target_id)) tmp[cur[r.target_id]++] = Tmp{r.target_start, (uint32_t)(i << 1)};
      if (bidirectional && r.query_id != r.target_id && owned(r.query_id))
        tmp[cur[r.query_id]++] = Tmp{r.query_start, (uint32_t)(i << 1) | 1u};
    }
  }
  REQUIRE(n < (1ull << 31), IMPGX_E_INVALID, "more than 2^31 records; shard the index");

  out.e_start.resize(E);
  out.e_end.resize(E);
  out.e_pmax.resize(E);
  out.e_vrank.resize(E);
  out.e_aln.resize(E);
  out.e_qid.resize(E);
  out.e_rec.resize(E);
#pragma omp parallel for schedule(dynamic, 16)
  for (long s = 0; s < (long)n_seqs; s++) {
    const uint64_t lo = out.tgt_off[s], hi = out.tgt_off[s + 1];
    if (lo == hi) continue;
    std::stable_sort(tmp.begin() + lo, tmp.begin() + hi, [](const Tmp &a, const Tmp &b) { return a.start < b.start; });
    visit_ranks(hi - lo, out.e_vrank.data() + lo);
    int32_t pm = INT32_MIN;
    for (uint64_t k = lo; k < hi; k++) {
      const uint32_t ri = tmp[k].rec >> 1;
      const bool reversed = tmp[k].rec & 1u;
      const impgx_record &r = recs[ri];
      EntryRec e;
      if (!reversed) {
        e.t_start = r.target_start; e.t_end = r.target_end;
        e.q_start = r.query_start; e.q_end = r.query_end;
        e.query_id = r.query_id;
      } else {
        e.t_start = r.query_start; e.t_end = r.query_end;
        e.q_start = r.target_start; e.q_end = r.target_end;
        e.query_id = r.target_id;
      }
      const uint32_t nr = (uint32_t)(run_offsets[ri + 1] - run_offsets[ri]);
      e.nruns_flags = (nr << 2) | (reversed ? FLAG_REVERSED : 0u) | (r.strand ? FLAG_STRAND : 0u);
      e.aln_off = out.aln_off[ri];
      e.vrank = out.e_vrank[k];
      out.e_aln[k] = ri;
      out.e_rec[k] = e;
      out.e_qid[k] = e.query_id;
      out.e_start[k] = e.t_start;
      out.e_end[k] = e.t_end;
      pm = std::max(pm, e.t_end);
      out.e_pmax[k] = pm;
    }
  }
}

}  // namespace impgx
