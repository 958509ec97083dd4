// bfs_kernels.cuh — frontier / visited-set / result-assembly kernels of the
// transitive BFS (reference src/impg.rs:2311-2597).
#pragma once
#include "kernels.cuh"

namespace impgx {

__device__ __forceinline__ uint64_t gtid() { return (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; }
__device__ __forceinline__ uint64_t gstride() { return (uint64_t)gridDim.x * blockDim.x; }

// rows -> level-0 frontier (one range per row) and the seed visited set.
__global__ void k_init_frontier(const impgx_range *__restrict__ ranges, uint32_t n, Frontier *__restrict__ fr) {
  for (uint64_t i = gtid(); i < n; i += gstride()) fr[i] = Frontier{(uint32_t)i, ranges[i].target_id, ranges[i].start, ranges[i].end};
}

// Validation of the request rows on the device (perform_query's bounds checks,
// src/main.rs:11620-11639, and parse_range's start < end).
__global__ void k_validate(const impgx_range *__restrict__ ranges, uint32_t n, const int32_t *__restrict__ seq_len,
                           uint32_t n_seqs, int *__restrict__ bad) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    impgx_range r = ranges[i];
    bool ok = r.target_id < n_seqs && r.start >= 0 && r.start < r.end;
    if (ok) ok = r.end <= seq_len[r.target_id];
    if (!ok) atomicExch(bad, (int)i + 1);
  }
}

// sort key of a lifted hit: (frontier index, coitrees visit rank); rejected hits last
// `orig` maps the processing order of the frontier (sorted by target position
// for cache locality) back to the reference's frontier index; nullptr = identity
__global__ void k_hit_order_keys(const Hit *__restrict__ hits, const LiftTask *__restrict__ tasks,
                                 const uint32_t *__restrict__ orig, uint64_t n, uint32_t n_frontier,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    const uint32_t r = orig ? orig[tasks[i].range] : tasks[i].range;
    keys[i] = h.row == INVALID_ID ? ((uint64_t)n_frontier << 32) : (((uint64_t)r << 32) | h.vrank);
    vals[i] = (uint32_t)i;
  }
}

// ---- MultiImpg hit order (src/multi_impg.rs:582-592): hits of one range sorted by
// (query id, query first, query last, target first, target last), signed compares.
// drop_mode 1 (transitive walk): hits onto the walked sequence itself are skipped
// before they are output (:888-891); drop_mode 2 (query): a hit equal to the self
// interval is a "duplicate self interval" and is dropped (:556-571).
__global__ void k_multi_drop(Hit *__restrict__ hits, const LiftTask *__restrict__ tasks, const Frontier *__restrict__ fr,
                             uint64_t n, int drop_mode, unsigned long long *__restrict__ n_ok) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    if (h.row == INVALID_ID) continue;
    bool drop = false;
    if (drop_mode == 1) drop = h.q_id == h.t_id;
    else if (drop_mode == 2) {
      const Frontier f = fr[tasks[i].range];
      drop = h.q_id == h.t_id && h.q_first == f.start && h.q_last == f.end;
    }
    if (drop) hits[i].row = INVALID_ID;
    else c++;
  }
  if (c) atomicAdd(n_ok, c);
}
// field: 0 t_last, 1 t_first, 2 q_last, 3 q_first
__global__ void k_multi_field_keys(const Hit *__restrict__ hits, const uint32_t *__restrict__ perm, uint64_t n, int field,
                                   uint32_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const Hit h = hits[perm[i]];
    const int32_t v = field == 0 ? h.t_last : (field == 1 ? h.t_first : (field == 2 ? h.q_last : h.q_first));
    keys[i] = (uint32_t)v ^ 0x80000000u;
  }
}
__global__ void k_multi_major_keys(const Hit *__restrict__ hits, const LiftTask *__restrict__ tasks,
                                   const uint32_t *__restrict__ orig, const uint32_t *__restrict__ perm, uint64_t n,
                                   uint32_t n_frontier, int seq_bits, uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const uint32_t j = perm[i];
    const Hit h = hits[j];
    const uint32_t r = orig ? orig[tasks[j].range] : tasks[j].range;
    keys[i] = h.row == INVALID_ID ? ((uint64_t)n_frontier << seq_bits) : (((uint64_t)r << seq_bits) | h.q_id);
  }
}

__global__ void k_locality_keys(const Frontier *__restrict__ fr, uint64_t n, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = ((uint64_t)fr[i].seq << 32) | (uint32_t)fr[i].start;
    vals[i] = (uint32_t)i;
  }
}

template <class T>
__global__ void k_gather(const T *__restrict__ src, const uint32_t *__restrict__ perm, uint64_t n, T *__restrict__ dst) {
  for (uint64_t i = gtid(); i < n; i += gstride()) dst[i] = src[perm[i]];
}

__global__ void k_gather_entry(const LiftTask *__restrict__ tasks, const uint32_t *__restrict__ perm, uint64_t n,
                               uint32_t *__restrict__ entry) {
  for (uint64_t i = gtid(); i < n; i += gstride()) entry[i] = tasks[perm[i]].entry;
}

// fold key: (row, q_id); hits back onto the frontier's own sequence are never
// expanded (src/impg.rs:2507) and sort last
__global__ void k_fold_keys(const Hit *__restrict__ hits, uint64_t n, uint32_t n_rows, uint64_t *__restrict__ keys,
                            uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    keys[i] = h.q_id == h.t_id ? ((uint64_t)n_rows << 32) : (((uint64_t)h.row << 32) | h.q_id);
    vals[i] = (uint32_t)i;
  }
}

// group heads over sorted keys; sentinel keys (>= limit) are not grouped.
__global__ void k_group_heads(const uint64_t *__restrict__ keys, uint64_t n, uint64_t limit,
                              uint64_t *__restrict__ head, uint64_t *__restrict__ n_included) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    uint64_t k = keys[i];
    bool inc = k < limit;
    head[i] = (inc && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
    if (!inc && (i == 0 || keys[i - 1] < limit)) *n_included = i;
  }
}

struct FoldGroup {
  uint64_t key;           // row << 32 | seq
  uint32_t hit_begin, hit_end;
  uint32_t v_begin, v_end;  // existing visited ranges of (row, seq)
  uint64_t list_off;      // scratch list offset (capacity n0 + h)
  uint64_t piece_off;     // piece slots offset (capacity n0 + 2h)
};

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t *a, uint64_t n, uint64_t key) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    uint64_t mid = lo + (hi - lo) / 2;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// head_scan = exclusive scan of head flags (head_scan[n] = group count)
// mask_off != nullptr: a (row, sequence) group that has no visited entry yet starts from the
// masked regions of the sequence (src/impg.rs:2331-2335) instead of an empty list
__global__ void k_make_groups(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ head,
                              const uint64_t *__restrict__ head_scan, uint64_t n, const uint64_t *__restrict__ n_included,
                              const uint64_t *__restrict__ v_keys, uint64_t v_n, FoldGroup *__restrict__ groups,
                              uint64_t *__restrict__ list_cap, uint64_t *__restrict__ piece_cap,
                              const uint64_t *__restrict__ mask_off) {
  const uint64_t ninc = *n_included;
  for (uint64_t i = gtid(); i < ninc; i += gstride()) {
    if (!head[i]) continue;
    const uint64_t g = head_scan[i];
    // end of the group = next head or the end of the included prefix
    uint64_t e = i + 1;
    const uint64_t k = keys[i];
    // groups are contiguous equal keys: binary search the end
    {
      uint64_t lo = i + 1, hi = ninc;
      while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (keys[mid] <= k) lo = mid + 1;
        else hi = mid;
      }
      e = lo;
    }
    FoldGroup fg;
    fg.key = k;
    fg.hit_begin = (uint32_t)i;
    fg.hit_end = (uint32_t)e;
    uint64_t vb = lower_bound_u64(v_keys, v_n, k), ve = lower_bound_u64(v_keys, v_n, k + 1);
    fg.v_begin = (uint32_t)vb;
    fg.v_end = (uint32_t)ve;
    fg.list_off = fg.piece_off = 0;
    groups[g] = fg;
    uint64_t n0 = ve - vb;
    const uint64_t h = e - i;
    if (n0 == 0 && mask_off) {
      const uint32_t seq = (uint32_t)k;
      n0 = mask_off[seq + 1] - mask_off[seq];
    }
    list_cap[g] = n0 + h;
    piece_cap[g] = n0 + 2 * h;
  }
}

__global__ void k_set_group_offsets(FoldGroup *__restrict__ groups, uint64_t n, const uint64_t *__restrict__ list_off,
                                    const uint64_t *__restrict__ piece_off) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    groups[i].list_off = list_off[i];
    groups[i].piece_off = piece_off[i];
  }
}

// SortedRanges::insert with min_distance = 0 (src/impg.rs:270-353) on a small
// array in global memory; returns the number of uncovered pieces written.
__device__ __forceinline__ uint32_t ranges_lower_bound(const int2 *L, uint32_t m, int32_t key) {
  uint32_t lo = 0, hi = m;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (L[mid].x < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void ranges_merge_forward_from(int2 *L, uint32_t &m, uint32_t start_idx) {
  uint32_t write = start_idx, read = start_idx + 1;
  while (read < m) {
    if (L[write].y >= L[read].x) {
      L[write].y = max(L[write].y, L[read].y);
    } else {
      write += 1;
      int2 t = L[write];
      L[write] = L[read];
      L[read] = t;
    }
    read += 1;
  }
  m = write + 1;
}

template <class Emit>
__device__ __forceinline__ void ranges_insert(int2 *L, uint32_t &m, int32_t seq_len, int32_t a, int32_t b, Emit emit) {
  int32_t start = a <= b ? a : b, end = a <= b ? b : a;
  // min_distance == 0: the neighbour-snap branches (|x| < 0) never fire
  if (start < 0) start = 0;
  if (end > seq_len) end = seq_len;
  int32_t current = start;
  uint32_t i = ranges_lower_bound(L, m, start);
  if (i > 0 && L[i - 1].y > start) i -= 1;
  while (i < m && current < end) {
    int2 r = L[i];
    if (r.x > end) break;
    if (current < r.x) emit(current, r.x);
    current = max(current, r.y);
    i += 1;
  }
  if (current < end) emit(current, end);
  uint32_t pos = ranges_lower_bound(L, m, start);
  if (pos > 0 && L[pos - 1].y >= start) {
    L[pos - 1].y = max(L[pos - 1].y, end);
    ranges_merge_forward_from(L, m, pos - 1);
  } else if (pos < m && end >= L[pos].x) {
    L[pos].x = min(start, L[pos].x);
    L[pos].y = max(end, L[pos].y);
    ranges_merge_forward_from(L, m, pos);
  } else {
    for (uint32_t k = m; k > pos; k--) L[k] = L[k - 1];
    L[pos] = make_int2(start, end);
    m += 1;
  }
}

// The order-sensitive sequential fold of one BFS level (src/impg.rs:2471-2560),
// one thread per (row, query sequence): the visited set is keyed by sequence,
// so hits only interact within such a group; inside the group they are
// processed in the reference's order (frontier order, then visit order).
__device__ __forceinline__ void fold_one_group(const FoldGroup &fg, const Hit *__restrict__ hits,
                                               const int32_t *__restrict__ v_start, const int32_t *__restrict__ v_end,
                                               const int32_t *__restrict__ seq_len, int32_t min_dist,
                                               int32_t min_transitive_len, int2 *__restrict__ lists,
                                               Frontier *__restrict__ pieces, const uint64_t *__restrict__ mask_off,
                                               const int2 *__restrict__ mask_rng, uint32_t &list_len_out,
                                               uint32_t &piece_cnt_out) {
  int2 *L = lists + fg.list_off;
  uint32_t m = 0;
  const uint32_t row = (uint32_t)(fg.key >> 32), seq = (uint32_t)fg.key;
  for (uint32_t v = fg.v_begin; v < fg.v_end; v++) L[m++] = make_int2(v_start[v], v_end[v]);
  if (fg.v_begin == fg.v_end && mask_off)
    for (uint64_t k = mask_off[seq]; k < mask_off[seq + 1]; k++) L[m++] = mask_rng[k];
  const int32_t slen = seq_len[seq];
  Frontier *P = pieces + fg.piece_off;
  uint32_t np = 0;
  int2 nxt = fg.hit_begin < fg.hit_end ? *reinterpret_cast<const int2 *>(&hits[fg.hit_begin].q_first) : make_int2(0, 0);
  for (uint32_t h = fg.hit_begin; h < fg.hit_end; h++) {
    // the next hit's interval is requested before this one's dependent walk through the list
    const int32_t a = nxt.x, b = nxt.y;
    if (h + 1 < fg.hit_end) nxt = *reinterpret_cast<const int2 *>(&hits[h + 1].q_first);
    bool should_add = true;
    if (min_dist > 0) {
      const int32_t new_min = min(a, b), new_max = max(a, b);
      const uint32_t idx = ranges_lower_bound(L, m, new_min);
      if (idx > 0) {
        long long d = (long long)new_min - (long long)L[idx - 1].y;
        if ((d < 0 ? -d : d) < min_dist) should_add = false;
      }
      if (should_add && idx < m) {
        long long d = (long long)L[idx].x - (long long)new_max;
        if ((d < 0 ? -d : d) < min_dist) should_add = false;
      }
    }
    if (should_add) {
      ranges_insert(L, m, slen, a, b, [&](int32_t s, int32_t e) {
        long long len = (long long)e - (long long)s;
        if ((len < 0 ? -len : len) >= min_transitive_len) P[np++] = Frontier{row, seq, s, e};
      });
    }
  }
  list_len_out = m;
  piece_cnt_out = np;
}

__global__ void __launch_bounds__(128) k_fold(const FoldGroup *__restrict__ groups, uint64_t n_groups,
                                              const Hit *__restrict__ hits, const int32_t *__restrict__ v_start,
                                              const int32_t *__restrict__ v_end, const int32_t *__restrict__ seq_len,
                                              int32_t min_dist, int32_t min_transitive_len, int2 *__restrict__ lists,
                                              uint32_t *__restrict__ list_len, Frontier *__restrict__ pieces,
                                              uint32_t *__restrict__ piece_cnt, const uint64_t *__restrict__ mask_off,
                                              const int2 *__restrict__ mask_rng) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    uint32_t m = 0, np = 0;
    fold_one_group(groups[g], hits, v_start, v_end, seq_len, min_dist, min_transitive_len, lists, pieces, mask_off,
                   mask_rng, m, np);
    list_len[g] = m;
    piece_cnt[g] = np;
  }
}

// compaction of per-group variable-length outputs
__global__ void k_compact_pieces(const FoldGroup *__restrict__ groups, uint64_t n_groups,
                                 const Frontier *__restrict__ pieces, const uint32_t *__restrict__ piece_cnt,
                                 const uint64_t *__restrict__ out_off, Frontier *__restrict__ out) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const Frontier *P = pieces + groups[g].piece_off;
    Frontier *O = out + out_off[g];
    for (uint32_t k = 0; k < piece_cnt[g]; k++) O[k] = P[k];
  }
}

__global__ void k_compact_lists(const FoldGroup *__restrict__ groups, uint64_t n_groups, const int2 *__restrict__ lists,
                                const uint32_t *__restrict__ list_len, const uint64_t *__restrict__ out_off,
                                uint64_t base, uint64_t *__restrict__ o_keys, int32_t *__restrict__ o_start,
                                int32_t *__restrict__ o_end) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const int2 *L = lists + groups[g].list_off;
    const uint64_t o = base + out_off[g];
    for (uint32_t k = 0; k < list_len[g]; k++) {
      o_keys[o + k] = groups[g].key;
      o_start[o + k] = L[k].x;
      o_end[o + k] = L[k].y;
    }
  }
}

// old visited entries whose (row, seq) was not touched this level are kept
__global__ void k_visited_keep_flags(const uint64_t *__restrict__ v_keys, uint64_t v_n,
                                     const FoldGroup *__restrict__ groups, uint64_t n_groups,
                                     uint64_t *__restrict__ keep) {
  for (uint64_t i = gtid(); i < v_n; i += gstride()) {
    const uint64_t k = v_keys[i];
    uint64_t lo = 0, hi = n_groups;
    while (lo < hi) {
      uint64_t mid = lo + (hi - lo) / 2;
      if (groups[mid].key < k) lo = mid + 1;
      else hi = mid;
    }
    keep[i] = (lo < n_groups && groups[lo].key == k) ? 0 : 1;
  }
}

__global__ void k_visited_copy_kept(const uint64_t *__restrict__ v_keys, const int32_t *__restrict__ v_start,
                                    const int32_t *__restrict__ v_end, uint64_t v_n, const uint64_t *__restrict__ keep,
                                    const uint64_t *__restrict__ keep_scan, uint64_t *__restrict__ o_keys,
                                    int32_t *__restrict__ o_start, int32_t *__restrict__ o_end) {
  for (uint64_t i = gtid(); i < v_n; i += gstride()) {
    if (!keep[i]) continue;
    const uint64_t o = keep_scan[i];
    o_keys[o] = v_keys[i];
    o_start[o] = v_start[i];
    o_end[o] = v_end[i];
  }
}

// ---- seeds under masked_regions: visited[target] starts from the mask, insert(range) returns
// the unmasked pieces (src/impg.rs:2337-2373); each piece is a result and, if long enough, frontier
__global__ void k_seed_mask_caps(const impgx_range *__restrict__ ranges, uint32_t n, const uint64_t *__restrict__ mask_off,
                                 uint64_t *__restrict__ cap) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const uint32_t t = ranges[i].target_id;
    cap[i] = mask_off[t + 1] - mask_off[t] + 1;
  }
}
__global__ void k_seed_masked(const impgx_range *__restrict__ ranges, uint32_t n, const uint64_t *__restrict__ mask_off,
                              const int2 *__restrict__ mask_rng, const int32_t *__restrict__ seq_len,
                              const uint64_t *__restrict__ off, int2 *__restrict__ lists, int2 *__restrict__ pieces,
                              uint32_t *__restrict__ list_len, uint32_t *__restrict__ piece_cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const impgx_range r = ranges[i];
    int2 *L = lists + off[i], *P = pieces + off[i];
    uint32_t m = 0, np = 0;
    for (uint64_t k = mask_off[r.target_id]; k < mask_off[r.target_id + 1]; k++) L[m++] = mask_rng[k];
    ranges_insert(L, m, seq_len[r.target_id], r.start, r.end, [&](int32_t s, int32_t e) { P[np++] = make_int2(s, e); });
    list_len[i] = m;
    piece_cnt[i] = np;
  }
}
__global__ void k_seed_masked_compact(const impgx_range *__restrict__ ranges, uint32_t n, const uint64_t *__restrict__ off,
                                      const int2 *__restrict__ lists, const int2 *__restrict__ pieces,
                                      const uint32_t *__restrict__ list_len, const uint32_t *__restrict__ piece_cnt,
                                      const uint64_t *__restrict__ list_scan, const uint64_t *__restrict__ piece_scan,
                                      uint64_t *__restrict__ v_keys, int32_t *__restrict__ v_start,
                                      int32_t *__restrict__ v_end, Hit *__restrict__ seed_hits,
                                      Frontier *__restrict__ seed_fr) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const uint32_t t = ranges[i].target_id;
    const int2 *L = lists + off[i], *P = pieces + off[i];
    for (uint32_t k = 0; k < list_len[i]; k++) {
      const uint64_t d = list_scan[i] + k;
      v_keys[d] = ((uint64_t)i << 32) | t;
      v_start[d] = L[k].x;
      v_end[d] = L[k].y;
    }
    for (uint32_t k = 0; k < piece_cnt[i]; k++) {
      const uint64_t d = piece_scan[i] + k;
      seed_hits[d] = Hit{(uint32_t)i, t, P[k].x, P[k].y, t, P[k].x, P[k].y, 0u};
      seed_fr[d] = Frontier{(uint32_t)i, t, P[k].x, P[k].y};
    }
  }
}
__global__ void k_fill_seed_slices(uint32_t *__restrict__ entry, CigarSlice *__restrict__ slices, uint64_t n) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    entry[i] = INVALID_ID;
    slices[i] = CigarSlice{0, 1, 0, 0};
  }
}

__global__ void k_seed_visited(const impgx_range *__restrict__ ranges, uint32_t n, uint64_t *__restrict__ keys,
                               int32_t *__restrict__ start, int32_t *__restrict__ end) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = ((uint64_t)i << 32) | ranges[i].target_id;
    start[i] = ranges[i].start;
    end[i] = ranges[i].end;
  }
}

__global__ void k_iota_u32(uint32_t *p, uint64_t n) {
  for (uint64_t i = gtid(); i < n; i += gstride()) p[i] = (uint32_t)i;
}

// frontier pieces: sort keys
__global__ void k_piece_start_keys(const Frontier *__restrict__ p, uint64_t n, uint32_t *__restrict__ keys,
                                   uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = (uint32_t)p[i].start ^ 0x80000000u;  // order-preserving for signed
    vals[i] = (uint32_t)i;
  }
}
__global__ void k_piece_seq_keys(const Frontier *__restrict__ p, const uint32_t *__restrict__ perm, uint64_t n,
                                 uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Frontier f = p[perm[i]];
    keys[i] = ((uint64_t)f.row << 32) | f.seq;
  }
}

// contiguous merge of the sorted next-depth ranges (src/impg.rs:2566-2584):
// same (row, id) and write.end >= read.start. Pieces of one sequence are
// pairwise disjoint (they are what SortedRanges::insert had not covered), so
// the running end is the previous piece's end and heads are independent.
__global__ void k_frontier_heads(const Frontier *__restrict__ p, uint64_t n, uint64_t *__restrict__ head) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    bool h = true;
    if (i > 0) {
      Frontier a = p[i - 1], b = p[i];
      h = !(a.row == b.row && a.seq == b.seq && a.end >= b.start);
    }
    head[i] = h ? 1 : 0;
  }
}
__global__ void k_frontier_merge(const Frontier *__restrict__ p, uint64_t n, const uint64_t *__restrict__ head,
                                 const uint64_t *__restrict__ head_scan, Frontier *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    if (!head[i]) continue;
    Frontier f = p[i];
    uint64_t j = i + 1;
    int32_t e = f.end;
    while (j < n && !head[j]) {
      e = max(e, p[j].end);
      j++;
    }
    f.end = e;
    out[head_scan[i]] = f;
  }
}

// ----------------------------------------------------------- DFS stacks
struct __align__(8) DfsEntry {
  uint32_t row, id;
  int32_t start, end;
  uint32_t depth, pad;
};

__global__ void k_dfs_init_stack(const Frontier *__restrict__ fr, uint64_t n, DfsEntry *__restrict__ st) {
  for (uint64_t i = gtid(); i < n; i += gstride()) st[i] = DfsEntry{fr[i].row, fr[i].seq, fr[i].start, fr[i].end, 0u, 0u};
}

// the stack array is sorted by (row, id, start): the top of a row is the last entry of its segment
// pop_front: MultiImpg's BFS takes the FIRST entry of the sorted queue (src/multi_impg.rs:855-860)
__global__ void k_dfs_pop(const DfsEntry *__restrict__ st, uint64_t n, uint32_t max_depth, int pop_front,
                          uint64_t *__restrict__ popped, Frontier *__restrict__ cand, uint64_t *__restrict__ is_fr,
                          uint32_t *__restrict__ cur_depth) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const DfsEntry e = st[i];
    if (pop_front ? (i > 0 && st[i - 1].row == e.row) : (i + 1 < n && st[i + 1].row == e.row)) continue;  // not the one popped
    popped[i] = 1;
    if (max_depth > 0 && e.depth >= max_depth) continue;  // :2125-2127, popped and dropped
    cand[e.row] = Frontier{e.row, e.id, e.start, e.end};
    is_fr[e.row] = 1;
    cur_depth[e.row] = e.depth;
  }
}

__global__ void k_dfs_keep_flags(const uint64_t *__restrict__ popped, uint64_t n, uint64_t *__restrict__ keep) {
  for (uint64_t i = gtid(); i <= n; i += gstride()) keep[i] = (i < n && !popped[i]) ? 1 : 0;
}
__global__ void k_dfs_copy_kept(const DfsEntry *__restrict__ st, uint64_t n, const uint64_t *__restrict__ popped,
                                const uint64_t *__restrict__ scan, DfsEntry *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (!popped[i]) out[scan[i]] = st[i];
}
__global__ void k_dfs_push(const Frontier *__restrict__ pieces, uint64_t n, const uint32_t *__restrict__ cur_depth,
                           DfsEntry *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const Frontier p = pieces[i];
    out[i] = DfsEntry{p.row, p.seq, p.start, p.end, cur_depth[p.row] + 1, 0u};
  }
}
__global__ void k_dfs_start_keys(const DfsEntry *__restrict__ st, uint64_t n, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = (uint32_t)st[i].start ^ 0x80000000u;
    vals[i] = (uint32_t)i;
  }
}
__global__ void k_dfs_seq_keys(const DfsEntry *__restrict__ st, const uint32_t *__restrict__ perm, uint64_t n,
                               uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const DfsEntry e = st[perm[i]];
    keys[i] = ((uint64_t)e.row << 32) | e.id;
  }
}
// one thread per row: the write/read merge of src/impg.rs:2291-2304 (depth of the survivor kept)
__global__ void __launch_bounds__(128) k_dfs_merge(DfsEntry *__restrict__ st, const uint32_t *__restrict__ begins,
                                                   uint64_t n_groups, uint32_t *__restrict__ cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], e = begins[g + 1];
    uint32_t w = b;
    for (uint32_t r = b + 1; r < e; r++) {
      if (st[w].id == st[r].id && st[w].end >= st[r].start) {
        st[w].end = max(st[w].end, st[r].end);
      } else {
        w++;
        DfsEntry t = st[w];
        st[w] = st[r];
        st[r] = t;
      }
    }
    cnt[g] = w - b + 1;
  }
}
__global__ void k_dfs_compact(const DfsEntry *__restrict__ st, const uint32_t *__restrict__ begins,
                              const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ scan, uint64_t n_groups,
                              DfsEntry *__restrict__ out) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g];
    for (uint32_t k = 0; k < cnt[g]; k++) out[scan[g] + k] = st[b + k];
  }
}

// ----------------------------------------------------------- result assembly
__device__ __forceinline__ bool passes_len(const Hit &h, int32_t min_out) {
  if (min_out < 0) return true;
  long long d = (long long)h.q_last - (long long)h.q_first;
  return (d < 0 ? -d : d) >= min_out;
}

// per-level: pass flag + per-row histogram
__global__ void k_level_pass(const Hit *__restrict__ hits, uint64_t n, int32_t min_out, uint64_t *__restrict__ pass,
                             uint32_t *__restrict__ row_cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    bool ok = passes_len(h, min_out);
    pass[i] = ok ? 1 : 0;
    if (ok) atomicAdd(&row_cnt[h.row], 1u);
  }
}

struct OutCols {
  uint32_t *q_id;
  int32_t *q_first, *q_last;
  uint32_t *t_id;
  int32_t *t_first, *t_last;
  uint64_t *cig_len;  // per result (later scanned into offsets) or nullptr
  uint32_t *src_entry;  // entry index per result (for CIGAR emission) or nullptr
  CigarSlice *src_slice;
};

// hits of one level are sorted by row; rank within (row, level) =
// pass_scan[i] - lvl_row_start[row]
__global__ void k_scatter_level(const Hit *__restrict__ hits, const uint32_t *__restrict__ entry,
                                const CigarSlice *__restrict__ slices, uint64_t n, const uint64_t *__restrict__ pass,
                                const uint64_t *__restrict__ pass_scan, const uint64_t *__restrict__ lvl_row_start,
                                const uint64_t *__restrict__ row_off, const uint32_t *__restrict__ base, OutCols o) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    if (!pass[i]) continue;
    Hit h = hits[i];
    uint64_t d = row_off[h.row] + base[h.row] + (pass_scan[i] - lvl_row_start[h.row]);
    o.q_id[d] = h.q_id;
    o.q_first[d] = h.q_first;
    o.q_last[d] = h.q_last;
    o.t_id[d] = h.t_id;
    o.t_first[d] = h.t_first;
    o.t_last[d] = h.t_last;
    if (o.cig_len) {
      o.cig_len[d] = slices[i].n_ops;
      o.src_entry[d] = entry[i];
      o.src_slice[d] = slices[i];
    }
  }
}

// the self interval (src/impg.rs:1864-1880 / :2345-2363)
__global__ void k_scatter_seed(const impgx_range *__restrict__ ranges, uint32_t n, const uint32_t *__restrict__ seed_cnt,
                               const uint64_t *__restrict__ row_off, OutCols o) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    if (!seed_cnt[i]) continue;
    impgx_range r = ranges[i];
    uint64_t d = row_off[i];
    o.q_id[d] = r.target_id;
    o.q_first[d] = r.start;
    o.q_last[d] = r.end;
    o.t_id[d] = r.target_id;
    o.t_first[d] = r.start;
    o.t_last[d] = r.end;
    if (o.cig_len) {
      o.cig_len[d] = 1;
      o.src_entry[d] = INVALID_ID;
      o.src_slice[d] = CigarSlice{0, 1, 0, 0};
    }
  }
}

__global__ void k_seed_counts(const impgx_range *__restrict__ ranges, uint32_t n, int32_t min_out, int apply_len,
                              uint32_t *__restrict__ seed_cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    bool ok = true;
    if (apply_len && min_out >= 0) ok = (ranges[i].end - ranges[i].start) >= min_out;
    seed_cnt[i] = ok ? 1u : 0u;
  }
}

__global__ void k_add_u32(uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint64_t n) {
  for (uint64_t i = gtid(); i < n; i += gstride()) a[i] += b[i];
}
__global__ void k_u32_to_u64(const uint32_t *__restrict__ a, uint64_t n, uint64_t *__restrict__ o) {
  for (uint64_t i = gtid(); i < n; i += gstride()) o[i] = a[i];
}

// CIGAR emission in final result order (warp per result)
__global__ void __launch_bounds__(256) k_emit_cigar_results(DevIndexView ix, OutCols o, const uint64_t *__restrict__ cig_off,
                                                            uint64_t n, uint32_t *__restrict__ out) {
  const unsigned lane = lane_id();
  uint64_t w = gtid() >> 5;
  const uint64_t nw = gstride() >> 5;
  for (; w < n; w += nw) {
    const CigarSlice s = o.src_slice[w];
    uint32_t *dst = out + cig_off[w];
    const uint32_t e = o.src_entry[w];
    if (e == INVALID_ID) {  // self interval: [len '=']
      if (lane == 0) dst[0] = (uint32_t)(o.q_last[w] - o.q_first[w]);
      continue;
    }
    const EntryRec rec = ix.e_rec[e];
    const uint32_t n_runs = rec.nruns_flags >> 2;
    const bool swap_id = rec.nruns_flags & FLAG_REVERSED;
    const bool backward = swap_id && (rec.nruns_flags & FLAG_STRAND);
    const uint32_t *blk = aln_runs(ix.stream, rec.aln_off, aln_nblk(n_runs));
    for (uint32_t k = lane; k < s.n_ops; k += 32) {
      uint32_t wi = s.first_idx + k;
      uint32_t v = blk[backward ? (n_runs - 1 - wi) : wi];
      uint32_t op = v >> 29;
      int32_t len = (int32_t)(v & 0x1fffffffu);
      if (swap_id) op = op == IMPGX_OP_I ? IMPGX_OP_D : (op == IMPGX_OP_D ? IMPGX_OP_I : op);
      if (k == 0 && s.first_off > 0) len -= s.first_off;
      if (k == s.n_ops - 1 && s.last_rem < 0) len += s.last_rem;
      dst[k] = (op << 29) | (uint32_t)len;
    }
  }
}

}  // namespace impgx
