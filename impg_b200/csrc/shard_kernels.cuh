// shard_kernels.cuh — routing kernels of the target-sharded pipeline
// (SURVEY.md §8e). A rank owns the sequences with owner[seq] == rank: it stabs
// and lifts the frontier ranges that lie on them, keeps the visited sets of
// (row, owned sequence) and merges the BED rows of (row, owned sequence).
// Lifted hits travel to owner(hit.q_id); the reference's result order travels
// with them as (global frontier index, coitrees visit rank).
#pragma once
#include "bucket_kernels.cuh"

namespace impgx {

constexpr int MAX_RANKS = 64;

// A lifted hit on its way to the rank that owns the sequence it lands on: what
// the fold needs (src/impg.rs:2467-2560) plus the global order key. 32 bytes.
struct __align__(32) RoutedHit {
  uint32_t row, q_id;
  int32_t q_first, q_last;
  uint32_t t_id;
  uint32_t gidx;   // index of the source range in the global frontier of the hop
  uint32_t vrank;  // visit rank of the entry within its target
  uint32_t pad;
};
static_assert(sizeof(RoutedHit) == 32, "RoutedHit must be one sector");

// level-0 frontier of a shard: rows on owned targets (BFS: long enough to expand)
__global__ void k_shard_seed_flags(const Frontier *__restrict__ f, uint64_t n, int32_t min_len,
                                   const uint32_t *__restrict__ owner, uint32_t rank, uint64_t *__restrict__ flag) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    long long d = (long long)f[i].start - (long long)f[i].end;
    flag[i] = ((d < 0 ? -d : d) >= min_len && owner[f[i].seq] == rank) ? 1 : 0;
  }
}
__global__ void k_frontier_rows(const Frontier *__restrict__ f, uint64_t n, uint32_t *__restrict__ gmap) {
  for (uint64_t i = gtid(); i < n; i += gstride()) gmap[i] = f[i].row;
}

// hits of one DFS round as boxes: a row pops at most one range per round, so (round, visit rank) is the position
// of a hit in the reference's result order of its row
__global__ void k_boxes_dfs_round(const Hit *__restrict__ hits, uint64_t n, uint32_t level, uint32_t round, int32_t min_out,
                                  BoxD *__restrict__ out, unsigned long long *__restrict__ n_valid) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const Hit h = hits[i];
    const bool ok = h.row != INVALID_ID && passes_len(h, min_out);
    out[i] = BoxD{h.q_first, h.q_last, h.t_first, h.t_last, h.q_id, h.t_id, h.row, ok ? 1u : 0u,
                  make_ord(level, ((uint64_t)round << 32) | h.vrank)};
    c += ok ? 1 : 0;
  }
  if (c) atomicAdd(n_valid, c);
}

// boxes of rows whose box sequence is not owned here are someone else's (masked seeds: every rank derives them all)
__global__ void k_boxes_keep_owned(BoxD *__restrict__ b, uint64_t n, const uint32_t *__restrict__ owner, uint32_t rank) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (owner[b[i].q_id] != rank) b[i].valid = 0;
}
// position in the uncompacted list of every element that was kept
__global__ void k_compact_indices(const uint64_t *__restrict__ flag, const uint64_t *__restrict__ scan, uint64_t n,
                                  uint32_t *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (flag[i]) out[scan[i]] = (uint32_t)i;
}

// accepted hits -> routed records + destination rank (rejected hits: dest = n_ranks, sorted last)
__global__ void k_route_hits(const Hit *__restrict__ hits, const LiftTask *__restrict__ tasks,
                             const uint32_t *__restrict__ orig, const uint32_t *__restrict__ gmap,
                             const uint32_t *__restrict__ owner, uint64_t n, uint32_t n_ranks,
                             RoutedHit *__restrict__ out, uint32_t *__restrict__ dest, uint32_t *__restrict__ idx,
                             unsigned long long *__restrict__ dest_cnt) {
  __shared__ unsigned int bins[MAX_RANKS + 1];
  for (unsigned k = threadIdx.x; k <= n_ranks; k += blockDim.x) bins[k] = 0;
  __syncthreads();
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const Hit h = hits[i];
    const uint32_t k = tasks[i].range;
    const uint32_t r = orig ? orig[k] : k;
    // hits back onto the frontier's own sequence are never expanded (src/impg.rs:2507): not routed
    const bool go = h.row != INVALID_ID && h.q_id != h.t_id;
    const uint32_t d = go ? owner[h.q_id] : n_ranks;
    out[i] = RoutedHit{h.row, h.q_id, h.q_first, h.q_last, h.t_id, gmap[r], h.vrank, 0u};
    dest[i] = d;
    idx[i] = (uint32_t)i;
    atomicAdd(&bins[d], 1u);
  }
  __syncthreads();
  for (unsigned k = threadIdx.x; k <= n_ranks; k += blockDim.x)
    if (bins[k]) atomicAdd(&dest_cnt[k], (unsigned long long)bins[k]);
}

// received records -> Hit (target coordinates are not needed by the fold) + order key
__global__ void k_routed_to_hits(const RoutedHit *__restrict__ in, uint64_t n, Hit *__restrict__ hits,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const RoutedHit r = in[i];
    hits[i] = Hit{r.row, r.q_id, r.q_first, r.q_last, r.t_id, 0, 0, r.vrank};
    keys[i] = ((uint64_t)r.gidx << 32) | r.vrank;
    idx[i] = (uint32_t)i;
  }
}

// global frontier order: (row, seq); ranges of one (row, seq) come from one rank, already sorted by start
__global__ void k_frontier_gkeys(const Frontier *__restrict__ f, uint64_t n, int seq_bits, uint64_t *__restrict__ keys,
                                 uint32_t *__restrict__ idx) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = ((uint64_t)f[i].row << seq_bits) | f[i].seq;
    idx[i] = (uint32_t)i;
  }
}
__global__ void k_frontier_own_flags(const Frontier *__restrict__ f, const uint32_t *__restrict__ perm, uint64_t n,
                                     const uint32_t *__restrict__ owner, uint32_t rank, uint64_t *__restrict__ flag) {
  for (uint64_t i = gtid(); i < n; i += gstride()) flag[i] = owner[f[perm[i]].seq] == rank ? 1 : 0;
}
// local frontier in global order; gmap = position in the global order
__global__ void k_frontier_take_owned(const Frontier *__restrict__ f, const uint32_t *__restrict__ perm, uint64_t n,
                                      const uint64_t *__restrict__ flag, const uint64_t *__restrict__ scan,
                                      Frontier *__restrict__ out, uint32_t *__restrict__ gmap) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (flag[i]) {
      out[scan[i]] = f[perm[i]];
      gmap[scan[i]] = (uint32_t)i;
    }
}

// ---- what survives the local reduction of the buckets (bucket_kernels.cuh, reduce mode) travels to the owner of
// its query sequence. Destination-major view of the bucket table: v = slot * n_rows + row, slot = position of the
// query sequence in the list of sequences ordered by (owner, id); the buckets one peer receives are contiguous in v.
__global__ void k_send_counts(const uint32_t *__restrict__ out_cnt, uint64_t n_buckets, uint32_t n_rows, uint32_t n_seqs,
                              const uint32_t *__restrict__ qorder, uint32_t *__restrict__ cv) {
  for (uint64_t v = gtid(); v < n_buckets; v += gstride()) {
    const uint32_t slot = (uint32_t)(v / n_rows), row = (uint32_t)(v % n_rows);
    cv[v] = out_cnt[(uint64_t)row * n_seqs + qorder[slot]];
  }
}
__global__ void k_send_copy(const BoxRec *__restrict__ boxes, const uint32_t *__restrict__ beg,
                            const uint32_t *__restrict__ out_cnt, const uint32_t *__restrict__ sv, uint64_t n_buckets,
                            uint32_t n_rows, uint32_t n_seqs, const uint32_t *__restrict__ qorder, BoxD *__restrict__ send) {
  for (uint64_t v = gtid(); v < n_buckets; v += gstride()) {
    const uint32_t slot = (uint32_t)(v / n_rows), row = (uint32_t)(v % n_rows);
    const uint32_t q = qorder[slot];
    const uint64_t b = (uint64_t)row * n_seqs + q;
    const uint32_t c = out_cnt[b];
    if (!c) continue;
    const BoxRec *seg = boxes + beg[b];
    BoxD *dst = send + sv[v];
    for (uint32_t k = 0; k < c; k++) {
      const BoxRec x = seg[k];
      dst[k] = BoxD{x.q_first, x.q_last, x.t_first, x.t_last, q, x.t_id, row,
                    1u | ((x.flags & BOX_MERGED_A) ? BOXD_MERGED_A : 0u), x.ord};
    }
  }
}

// valid boxes -> destination rank = owner of the query sequence (invalid ones go nowhere)
__global__ void k_box_dest(BoxSrc src, uint64_t n, const uint32_t *__restrict__ owner,
                           uint32_t n_ranks, uint32_t *__restrict__ dest, uint32_t *__restrict__ idx,
                           unsigned long long *__restrict__ dest_cnt) {
  __shared__ unsigned int bins[MAX_RANKS + 1];
  for (unsigned k = threadIdx.x; k <= n_ranks; k += blockDim.x) bins[k] = 0;
  __syncthreads();
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    uint32_t q;
    bool ok;
    if (i < src.n_boxes) {
      q = src.boxes[i].q_id;
      ok = src.boxes[i].valid != 0;
    } else {
      const Hit h = src.hits[i - src.n_boxes];
      q = h.q_id;
      ok = h.row != INVALID_ID && passes_len(h, src.min_out);
    }
    const uint32_t d = ok ? owner[q] : n_ranks;
    dest[i] = d;
    idx[i] = (uint32_t)i;
    atomicAdd(&bins[d], 1u);
  }
  __syncthreads();
  for (unsigned k = threadIdx.x; k <= n_ranks; k += blockDim.x)
    if (bins[k]) atomicAdd(&dest_cnt[k], (unsigned long long)bins[k]);
}

}  // namespace impgx
