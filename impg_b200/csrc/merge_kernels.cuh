// merge_kernels.cuh — output_results_bed's two merges on the device
// (src/main.rs:11849-11866): merge_adjusted_intervals_gap_2d (:12858-13011)
// then merge_query_adjusted_intervals (:12474-12560). Results are brought into
// the group order each merge sorts by (radix sorts in engine.cu); the group
// bodies below are literal per-group restatements, one thread per group.
#pragma once
#include "bfs_kernels.cuh"

namespace impgx {

struct ResCols {
  const uint32_t *q_id;
  const int32_t *q_first, *q_last;
  const uint32_t *t_id;
  const int32_t *t_first, *t_last;
};

__global__ void k_fill_rows(const uint64_t *__restrict__ row_off, uint32_t n_rows, uint64_t n, uint32_t *__restrict__ row) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    // last r with row_off[r] <= i
    uint32_t lo = 0, hi = n_rows;
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo) / 2;
      if (row_off[mid + 1] <= i) lo = mid + 1;
      else hi = mid;
    }
    row[i] = lo;
  }
}

__global__ void k_start_keys_i32(const int32_t *__restrict__ v, uint64_t n, uint32_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) keys[i] = (uint32_t)v[i] ^ 0x80000000u;
}

__global__ void k_row_targets(const impgx_range *__restrict__ ranges, uint32_t n, uint32_t *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) out[i] = ranges[i].target_id;
}

__global__ void k_frontier_len_flags(const Frontier *__restrict__ f, uint64_t n, int32_t min_len, uint64_t *__restrict__ flag) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    long long d = (long long)f[i].start - (long long)f[i].end;
    flag[i] = ((d < 0 ? -d : d) >= min_len) ? 1 : 0;
  }
}
__global__ void k_frontier_compact(const Frontier *__restrict__ f, uint64_t n, const uint64_t *__restrict__ flag,
                                   const uint64_t *__restrict__ scan, Frontier *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (flag[i]) out[scan[i]] = f[i];
}

// ---- stage A: 2D gap merge
// sort key 1: q.first (forward) or -q.first (reverse), src/main.rs:12885-12892
__global__ void k_m2d_key1(ResCols r, uint64_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    int32_t f = r.q_first[i];
    bool fwd = f <= r.q_last[i];
    int32_t k = fwd ? f : -f;
    keys[i] = (uint32_t)k ^ 0x80000000u;
    vals[i] = (uint32_t)i;
  }
}
// sort key 2: (row, q_id, t_id, strand) packed
__global__ void k_m2d_key2(ResCols r, const uint32_t *__restrict__ row, const uint32_t *__restrict__ perm, uint64_t n,
                           int seq_bits, uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    uint32_t s = perm[i];
    bool fwd = r.q_first[s] <= r.q_last[s];
    keys[i] = ((((uint64_t)row[s] << seq_bits | r.q_id[s]) << seq_bits | r.t_id[s]) << 1) | (fwd ? 1u : 0u);
  }
}

__global__ void k_heads_u64(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ head) {
  for (uint64_t i = gtid(); i < n; i += gstride()) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void k_group_begins(const uint64_t *__restrict__ head, const uint64_t *__restrict__ head_scan, uint64_t n,
                               uint32_t *__restrict__ begins) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    if (head[i]) begins[head_scan[i]] = (uint32_t)i;
}

struct Box {
  int32_t q_lo, q_hi, t_lo, t_hi;
  uint32_t min_idx;  // smallest input index of the members ("first untaken index", :12948-12958)
  uint32_t q_id, t_id, row;
};

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x) {
  while (parent[x] != x) {
    parent[x] = parent[parent[x]];
    x = parent[x];
  }
  return x;
}

// one thread per (row, q_id, t_id, strand) group; positions are sorted by key1 (stable)
__global__ void __launch_bounds__(128) k_merge2d(ResCols r, const uint32_t *__restrict__ row,
                                                 const uint32_t *__restrict__ perm, const uint32_t *__restrict__ begins,
                                                 uint64_t n_groups, int64_t d, uint32_t *__restrict__ parent,
                                                 Box *__restrict__ box, uint64_t *__restrict__ is_root) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], e = begins[g + 1];
    for (uint32_t a = b; a < e; a++) {
      parent[a] = a;
      is_root[a] = 0;
    }
    const uint32_t i0 = perm[b];
    const bool fwd = r.q_first[i0] <= r.q_last[i0];
    for (uint32_t a = b; a < e; a++) {
      const uint32_t ia = perm[a];
      const int64_t qa_start = fwd ? r.q_first[ia] : r.q_last[ia];
      const int64_t qa_end = fwd ? r.q_last[ia] : r.q_first[ia];
      const int64_t ta_start = r.t_first[ia], ta_end = r.t_last[ia];
      for (uint32_t bb = a + 1; bb < e; bb++) {
        const uint32_t ib = perm[bb];
        const int64_t qb_start = fwd ? r.q_first[ib] : r.q_last[ib];
        if (qb_start < qa_start) continue;
        const int64_t q_gap = qb_start - qa_end;
        if (q_gap > d) break;
        const int64_t tb_start = r.t_first[ib], tb_end = r.t_last[ib];
        int64_t t_gap;
        bool t_forward;
        if (fwd) {
          t_gap = tb_start - ta_end;
          t_forward = tb_start > ta_start;
        } else {
          t_gap = ta_start - tb_end;
          t_forward = tb_end < ta_end;
        }
        if (!t_forward || t_gap > d) continue;
        uint32_t ra = uf_find(parent, a), rb = uf_find(parent, bb);
        if (ra != rb) parent[ra] = rb;
      }
    }
    for (uint32_t a = b; a < e; a++) {
      const uint32_t rt = uf_find(parent, a);
      const uint32_t ia = perm[a];
      const int32_t qf = r.q_first[ia], ql = r.q_last[ia], tf = r.t_first[ia], tl = r.t_last[ia];
      if (!is_root[rt]) {
        is_root[rt] = 1;
        box[rt] = Box{qf, ql, tf, tl, ia, r.q_id[ia], r.t_id[ia], row[ia]};
      } else {
        Box bx = box[rt];
        if (fwd) {
          bx.q_lo = min(bx.q_lo, qf);
          bx.q_hi = max(bx.q_hi, ql);
        } else {
          bx.q_lo = max(bx.q_lo, qf);
          bx.q_hi = min(bx.q_hi, ql);
        }
        bx.t_lo = min(bx.t_lo, tf);
        bx.t_hi = max(bx.t_hi, tl);
        bx.min_idx = min(bx.min_idx, ia);
        box[rt] = bx;
      }
    }
  }
}

__global__ void k_compact_boxes(const Box *__restrict__ box, const uint64_t *__restrict__ is_root,
                                const uint64_t *__restrict__ scan, uint64_t n, Box *__restrict__ out,
                                uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    if (!is_root[i]) continue;
    uint64_t o = scan[i];
    out[o] = box[i];
    keys[o] = box[i].min_idx;
    vals[o] = (uint32_t)o;
  }
}

// raw results -> Box form (when the 2D merge is skipped)
__global__ void k_results_to_boxes(ResCols r, const uint32_t *__restrict__ row, uint64_t n, Box *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride())
    out[i] = Box{r.q_first[i], r.q_last[i], r.t_first[i], r.t_last[i], (uint32_t)i, r.q_id[i], r.t_id[i], row[i]};
}

// ---- stage B: query-axis merge
// sort key 1: (start, !is_forward), src/main.rs:12481-12494
__global__ void k_mq_key1(const Box *__restrict__ b, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Box x = b[i];
    bool fwd = x.q_lo <= x.q_hi;
    int32_t start = fwd ? x.q_lo : x.q_hi;
    keys[i] = ((uint64_t)((uint32_t)start ^ 0x80000000u) << 1) | (fwd ? 0u : 1u);
    vals[i] = (uint32_t)i;
  }
}
__global__ void k_mq_key2(const Box *__restrict__ b, const uint32_t *__restrict__ perm, uint64_t n,
                          uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Box x = b[perm[i]];
    keys[i] = ((uint64_t)x.row << 32) | x.q_id;
  }
}

// one thread per (row, q_id) segment: the literal sweep of :12496-12556 on the
// sorted boxes; merged intervals are written in place at the segment's slots.
__global__ void __launch_bounds__(128) k_sweep(const Box *__restrict__ sorted, const uint32_t *__restrict__ begins,
                                               uint64_t n_groups, int32_t merge_distance, int merge_strands,
                                               Box *__restrict__ out, uint32_t *__restrict__ out_cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], e = begins[g + 1];
    uint32_t w = b;
    Box cur = sorted[b];
    for (uint32_t rd = b + 1; rd < e; rd++) {
      const Box nx = sorted[rd];
      const bool cf = cur.q_lo <= cur.q_hi, nf = nx.q_lo <= nx.q_hi;
      const int32_t cs = cf ? cur.q_lo : cur.q_hi, ce = cf ? cur.q_hi : cur.q_lo;
      const int32_t ns = nf ? nx.q_lo : nx.q_hi, ne = nf ? nx.q_hi : nx.q_lo;
      if (merge_distance < 0 || (!merge_strands && cf != nf) || (int64_t)ns > (int64_t)ce + merge_distance) {
        out[w++] = cur;
        cur = nx;
      } else {
        const int32_t ms = min(cs, ns), me = max(ce, ne);
        bool mf = cf;
        if (merge_strands && cf != nf) {
          const int64_t cl = (int64_t)ce - cs, nl = (int64_t)ne - ns;
          mf = nl > cl ? nf : cf;
        }
        cur.q_lo = mf ? ms : me;
        cur.q_hi = mf ? me : ms;
      }
    }
    out[w++] = cur;
    out_cnt[g] = w - b;
  }
}

__global__ void k_sweep_compact(const Box *__restrict__ out, const uint32_t *__restrict__ begins,
                                const uint32_t *__restrict__ out_cnt, const uint64_t *__restrict__ scan,
                                uint64_t n_groups, OutCols o, uint32_t *__restrict__ row_cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], c = out_cnt[g];
    const uint64_t d0 = scan[g];
    for (uint32_t k = 0; k < c; k++) {
      Box x = out[b + k];
      o.q_id[d0 + k] = x.q_id;
      o.q_first[d0 + k] = x.q_lo;
      o.q_last[d0 + k] = x.q_hi;
      o.t_id[d0 + k] = x.t_id;
      o.t_first[d0 + k] = x.t_lo;
      o.t_last[d0 + k] = x.t_hi;
    }
    if (c) atomicAdd(&row_cnt[out[b].row], c);
  }
}

__global__ void k_boxes_to_cols(const Box *__restrict__ b, uint64_t n, OutCols o, uint32_t *__restrict__ row_cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Box x = b[i];
    o.q_id[i] = x.q_id;
    o.q_first[i] = x.q_lo;
    o.q_last[i] = x.q_hi;
    o.t_id[i] = x.t_id;
    o.t_first[i] = x.t_lo;
    o.t_last[i] = x.t_hi;
    atomicAdd(&row_cnt[x.row], 1u);
  }
}

}  // namespace impgx

// ===================================================================
// Direct BED path: the two merges straight from the lifted hits, without
// materialising the raw per-row result list. Reference order is carried by a
// 64-bit ordinal per result (`ord`): it is only needed to break ties exactly
// as the reference's stable sorts would.
namespace impgx {

struct __align__(8) BoxD {  // 40 bytes: what travels between the ranks of a sharded index
  int32_t q_lo, q_hi, t_lo, t_hi;
  uint32_t q_id, t_id, row, valid;  // valid: bit 0 = a result, BOXD_MERGED_A = a stage-A result already
  uint64_t ord;  // position in the reference's result order of the row: level << 58 | (range << 32 | vrank) or index
};

constexpr uint32_t BOXD_MERGED_A = 2u;

__device__ __forceinline__ uint64_t make_ord(uint32_t level, uint64_t low) { return ((uint64_t)level << 58) | low; }

// Where the boxes of a batch live: the seeds and the ordered levels as BoxD records
// [0, n_boxes), then the raw hits of the last hop exactly as the liftover kernel wrote them
// (box i >= n_boxes is hit i - n_boxes; its ordinal comes from the lift task's range and the
// hit's visit rank), so the bulk of the results is never copied into a second layout.
struct BoxSrc {
  const BoxD *boxes;
  uint64_t n_boxes;
  const Hit *hits;
  const LiftTask *tasks;
  const uint32_t *orig;  // processing index -> frontier index of the hop, or nullptr
  uint32_t level;
  int32_t min_out;
  const uint32_t *gmap;  // sharded index: local frontier index -> index in the GLOBAL frontier, or nullptr
};
__device__ __forceinline__ BoxD load_box(const BoxSrc &s, uint32_t i) {
  if (i < s.n_boxes) return s.boxes[i];
  const uint64_t j = i - s.n_boxes;
  const Hit h = s.hits[j];
  const bool ok = h.row != INVALID_ID && passes_len(h, s.min_out);
  const uint32_t k = s.tasks[j].range;
  uint32_t r = s.orig ? s.orig[k] : k;
  r = s.gmap ? s.gmap[r] : r;
  return BoxD{h.q_first, h.q_last, h.t_first, h.t_last, h.q_id, h.t_id, h.row, ok ? 1u : 0u,
              make_ord(s.level, ((uint64_t)r << 32) | h.vrank)};
}
__global__ void k_gather_src(BoxSrc src, const uint32_t *__restrict__ idx, uint64_t n, BoxD *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) out[i] = load_box(src, idx[i]);
}

// `owner` != nullptr (sharded index): only the rows on targets owned by `rank` are valid here
__global__ void k_boxes_from_seeds(const impgx_range *__restrict__ ranges, uint32_t n, int32_t min_out, int apply_len,
                                   BoxD *__restrict__ out, unsigned long long *__restrict__ n_valid,
                                   const uint32_t *__restrict__ owner, uint32_t rank) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    impgx_range r = ranges[i];
    bool ok = !owner || owner[r.target_id] == rank;
    if (apply_len && min_out >= 0) ok = ok && (r.end - r.start) >= min_out;
    out[i] = BoxD{r.start, r.end, r.start, r.end, r.target_id, r.target_id, (uint32_t)i, ok ? 1u : 0u, make_ord(0, 0)};
    if (ok) atomicAdd(n_valid, 1ull);
  }
}

// hits already in reference order (levels that were sorted for the fold)
__global__ void k_boxes_from_sorted_level(const Hit *__restrict__ hits, uint64_t n, uint32_t level, int32_t min_out,
                                          BoxD *__restrict__ out, unsigned long long *__restrict__ n_valid) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    bool ok = passes_len(h, min_out);
    out[i] = BoxD{h.q_first, h.q_last, h.t_first, h.t_last, h.q_id, h.t_id, h.row, ok ? 1u : 0u, make_ord(level, i)};
    c += ok ? 1 : 0;
  }
  if (c) atomicAdd(n_valid, c);
}

// hits of the last level in task order (frontier range, sorted position); ord from (range, visit rank)
// With a locality-permuted frontier (`orig` != nullptr) the boxes are written
// back in the reference's frontier order (dst_off = offsets of the ranges in
// that order), so the merge sorts see row-grouped input. `gmap` != nullptr
// (sharded index): the order key uses the range's index in the GLOBAL frontier.
__global__ void k_boxes_from_raw_level(const Hit *__restrict__ hits, const LiftTask *__restrict__ tasks,
                                       const uint32_t *__restrict__ orig, const uint64_t *__restrict__ offs,
                                       const uint64_t *__restrict__ dst_off, uint64_t n, uint32_t level, int32_t min_out,
                                       BoxD *__restrict__ out, unsigned long long *__restrict__ n_valid,
                                       const uint32_t *__restrict__ gmap) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    Hit h = hits[i];
    bool ok = h.row != INVALID_ID && passes_len(h, min_out);
    const uint32_t k = tasks[i].range;
    const uint32_t r = orig ? orig[k] : k;
    const uint64_t dst = (orig && dst_off) ? dst_off[r] + (i - offs[k]) : i;
    out[dst] = BoxD{h.q_first, h.q_last, h.t_first, h.t_last, h.q_id, h.t_id, h.row, ok ? 1u : 0u,
                    make_ord(level, ((uint64_t)(gmap ? gmap[r] : r) << 32) | h.vrank)};
    c += ok ? 1 : 0;
  }
  if (c) atomicAdd(n_valid, c);
}

// stage A key: (row, q_id, t_id, strand); invalid boxes sort last
__global__ void k_bd_key_a(const BoxD *__restrict__ b, uint64_t n, int seq_bits, uint64_t invalid_key,
                           uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    BoxD x = b[i];
    bool fwd = x.q_lo <= x.q_hi;
    keys[i] = x.valid ? (((((uint64_t)x.row << seq_bits | x.q_id) << seq_bits | x.t_id) << 1) | (fwd ? 1u : 0u)) : invalid_key;
    vals[i] = (uint32_t)i;
  }
}

// the same key over a BoxSrc; counts the valid boxes
__global__ void k_bd_key_a_src(BoxSrc src, uint64_t n, int seq_bits, uint64_t invalid_key, uint64_t *__restrict__ keys,
                               uint32_t *__restrict__ vals, unsigned long long *__restrict__ n_valid) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    uint32_t row, q, t;
    bool fwd, ok;
    if (i < src.n_boxes) {
      const BoxD x = src.boxes[i];
      row = x.row; q = x.q_id; t = x.t_id;
      fwd = x.q_lo <= x.q_hi;
      ok = x.valid != 0;
    } else {
      const Hit h = src.hits[i - src.n_boxes];
      row = h.row; q = h.q_id; t = h.t_id;
      fwd = h.q_first <= h.q_last;
      ok = h.row != INVALID_ID && passes_len(h, src.min_out);
    }
    keys[i] = ok ? (((((uint64_t)row << seq_bits | q) << seq_bits | t) << 1) | (fwd ? 1u : 0u)) : invalid_key;
    vals[i] = (uint32_t)i;
    c += ok ? 1 : 0;
  }
  if (c) atomicAdd(n_valid, c);
}

// merge_adjusted_intervals_gap_2d (src/main.rs:12858-13011) on an UNSORTED
// group: the reference sorts the group by (q.first | -q.first, input order) and
// scans pairs a < b; the relation tested for a pair does not depend on the
// other members (its `break` only prunes pairs that fail q_gap <= d anyway on
// the forward strand and never fires on the reverse strand, where
// qb.last < qb.first <= qa.first), so the partition is a pure pairwise
// property; `a` is the member with the smaller (sort key, ord).
__global__ void __launch_bounds__(128) k_merge2d_direct(const BoxD *__restrict__ boxes, const uint32_t *__restrict__ perm,
                                                        const uint32_t *__restrict__ begins, uint64_t n_groups, int64_t d,
                                                        uint32_t *__restrict__ parent, BoxD *__restrict__ acc,
                                                        uint64_t *__restrict__ is_root) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], e = begins[g + 1];
    // singleton (the common case), merging disabled (:12859), or a group of stage-A results (merged where
    // they were produced: all members of a group carry the flag or none does)
    if (e - b == 1 || d < 0 || (boxes[perm[b]].valid & BOXD_MERGED_A)) {
      for (uint32_t a = b; a < e; a++) {
        acc[a] = boxes[perm[a]];
        is_root[a] = 1;
      }
      continue;
    }
    for (uint32_t a = b; a < e; a++) {
      parent[a] = a;
      is_root[a] = 0;
    }
    const BoxD first = boxes[perm[b]];
    const bool fwd = first.q_lo <= first.q_hi;
    for (uint32_t i = b; i < e; i++) {
      const BoxD X = boxes[perm[i]];
      const int64_t kx = fwd ? (int64_t)X.q_lo : -(int64_t)X.q_lo;
      for (uint32_t j = i + 1; j < e; j++) {
        const BoxD Y = boxes[perm[j]];
        const int64_t ky = fwd ? (int64_t)Y.q_lo : -(int64_t)Y.q_lo;
        const bool x_first = kx < ky || (kx == ky && X.ord < Y.ord);
        const BoxD &A = x_first ? X : Y;
        const BoxD &B = x_first ? Y : X;
        const int64_t qa_start = fwd ? A.q_lo : A.q_hi, qa_end = fwd ? A.q_hi : A.q_lo;
        const int64_t qb_start = fwd ? B.q_lo : B.q_hi;
        if (qb_start < qa_start) continue;
        if (qb_start - qa_end > d) continue;
        int64_t t_gap;
        bool t_forward;
        if (fwd) {
          t_gap = (int64_t)B.t_lo - A.t_hi;
          t_forward = B.t_lo > A.t_lo;
        } else {
          t_gap = (int64_t)A.t_lo - B.t_hi;
          t_forward = B.t_hi < A.t_hi;
        }
        if (!t_forward || t_gap > d) continue;
        uint32_t ra = uf_find(parent, i), rb = uf_find(parent, j);
        if (ra != rb) parent[ra] = rb;
      }
    }
    for (uint32_t a = b; a < e; a++) {
      const uint32_t rt = uf_find(parent, a);
      const BoxD X = boxes[perm[a]];
      if (!is_root[rt]) {
        is_root[rt] = 1;
        acc[rt] = X;
      } else {
        BoxD bx = acc[rt];
        if (fwd) {
          bx.q_lo = min(bx.q_lo, X.q_lo);
          bx.q_hi = max(bx.q_hi, X.q_hi);
        } else {
          bx.q_lo = max(bx.q_lo, X.q_lo);
          bx.q_hi = min(bx.q_hi, X.q_hi);
        }
        bx.t_lo = min(bx.t_lo, X.t_lo);
        bx.t_hi = max(bx.t_hi, X.t_hi);
        bx.ord = min(bx.ord, X.ord);  // output order = first member in input order (:12948-12958)
        acc[rt] = bx;
      }
    }
  }
}

// stage B key: (row, q_id, start, !is_forward) — src/main.rs:12481-12494
__global__ void k_bd_key_b(const BoxD *__restrict__ acc, const uint64_t *__restrict__ is_root, uint64_t n, int seq_bits,
                           uint64_t invalid_key, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals,
                           unsigned long long *__restrict__ n_roots) {
  unsigned long long c = 0;
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    if (is_root && !is_root[i]) {  // is_root == nullptr: every box is a root (received from the peers)
      keys[i] = invalid_key;
      vals[i] = (uint32_t)i;
      continue;
    }
    c++;
    BoxD x = acc[i];
    bool fwd = x.q_lo <= x.q_hi;
    uint32_t start = (uint32_t)(fwd ? x.q_lo : x.q_hi);
    keys[i] = ((((uint64_t)x.row << seq_bits | x.q_id) << 32 | start) << 1) | (fwd ? 0u : 1u);
    vals[i] = (uint32_t)i;
  }
  if (c) atomicAdd(n_roots, c);
}

// one thread per (row, q_id) segment: restore the reference's tie order (equal
// (start, strand) keys are ordered by ord, as its stable sort leaves them),
// then the literal sweep of src/main.rs:12496-12556.
__global__ void __launch_bounds__(128) k_sweep_direct(BoxD *__restrict__ sorted, const uint64_t *__restrict__ keys,
                                                      const uint32_t *__restrict__ begins, uint64_t n_groups,
                                                      int32_t merge_distance, int merge_strands, BoxD *__restrict__ out,
                                                      uint32_t *__restrict__ out_cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], e = begins[g + 1];
    // tie fix-up: insertion sort by ord inside runs of equal keys
    for (uint32_t i = b + 1; i < e; i++) {
      if (keys[i] != keys[i - 1]) continue;
      BoxD x = sorted[i];
      uint32_t j = i;
      while (j > b && keys[j - 1] == keys[i] && sorted[j - 1].ord > x.ord) {
        sorted[j] = sorted[j - 1];
        j--;
      }
      sorted[j] = x;
    }
    uint32_t w = b;
    BoxD cur = sorted[b];
    for (uint32_t rd = b + 1; rd < e; rd++) {
      const BoxD nx = sorted[rd];
      const bool cf = cur.q_lo <= cur.q_hi, nf = nx.q_lo <= nx.q_hi;
      const int32_t cs = cf ? cur.q_lo : cur.q_hi, ce = cf ? cur.q_hi : cur.q_lo;
      const int32_t ns = nf ? nx.q_lo : nx.q_hi, ne = nf ? nx.q_hi : nx.q_lo;
      if (merge_distance < 0 || (!merge_strands && cf != nf) || (int64_t)ns > (int64_t)ce + merge_distance) {
        out[w++] = cur;
        cur = nx;
      } else {
        const int32_t ms = min(cs, ns), me = max(ce, ne);
        bool mf = cf;
        if (merge_strands && cf != nf) {
          const int64_t cl = (int64_t)ce - cs, nl = (int64_t)ne - ns;
          mf = nl > cl ? nf : cf;
        }
        cur.q_lo = mf ? ms : me;
        cur.q_hi = mf ? me : ms;
      }
    }
    out[w++] = cur;
    out_cnt[g] = w - b;
  }
}

__global__ void k_sweep_compact_direct(const BoxD *__restrict__ out, const uint32_t *__restrict__ begins,
                                       const uint32_t *__restrict__ out_cnt, const uint64_t *__restrict__ scan,
                                       uint64_t n_groups, OutCols o, uint32_t *__restrict__ row_cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], c = out_cnt[g];
    const uint64_t d0 = scan[g];
    for (uint32_t k = 0; k < c; k++) {
      BoxD x = out[b + k];
      o.q_id[d0 + k] = x.q_id;
      o.q_first[d0 + k] = x.q_lo;
      o.q_last[d0 + k] = x.q_hi;
      o.t_id[d0 + k] = x.t_id;
      o.t_first[d0 + k] = x.t_lo;
      o.t_last[d0 + k] = x.t_hi;
    }
    if (c) atomicAdd(&row_cnt[out[b].row], c);
  }
}

// counts in processing order -> counts in reference frontier order
__global__ void k_scatter_counts(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ orig, uint64_t n,
                                 uint64_t *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) out[orig[i]] = counts[i];
}

__global__ void k_keys_shift(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint64_t *__restrict__ out) {
  for (uint64_t i = gtid(); i < n; i += gstride()) out[i] = keys[i] >> shift;
}

// no merging at all (-d < 0 and strands kept apart): boxes in reference order
__global__ void k_bd_key_ord(const BoxD *__restrict__ b, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    keys[i] = b[i].ord;
    vals[i] = (uint32_t)i;
  }
}
__global__ void k_bd_key_row(const BoxD *__restrict__ b, const uint32_t *__restrict__ perm, uint64_t n, uint32_t n_rows,
                             uint64_t *__restrict__ keys) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    BoxD x = b[perm[i]];
    keys[i] = x.valid ? x.row : n_rows;
  }
}
__global__ void k_boxd_to_cols(const BoxD *__restrict__ b, const uint32_t *__restrict__ perm, uint64_t n, OutCols o,
                               uint32_t *__restrict__ row_cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    BoxD x = b[perm[i]];
    o.q_id[i] = x.q_id;
    o.q_first[i] = x.q_lo;
    o.q_last[i] = x.q_hi;
    o.t_id[i] = x.t_id;
    o.t_first[i] = x.t_lo;
    o.t_last[i] = x.t_hi;
    atomicAdd(&row_cnt[x.row], 1u);
  }
}

}  // namespace impgx


// ===================================================================
// Fused BED merge: after ONE global sort of the boxes by (row, q, t, strand)
// every (row, q) segment is merged on chip — stage A (pairwise union-find per
// (t, strand) group, contiguous inside the segment), then stage B (sort of the
// merged boxes by (start, strand, ord) in shared memory + the sweep). A box is
// read from HBM once. One warp per small segment, one CTA per larger one; a
// batch with a segment beyond SEG_MAX falls back to the global path.
namespace impgx {

constexpr int SEG_CLASSES = 5;
// boxes per segment (powers of two: the sort pads to one): two warp classes, then 256-, 128- and 512-thread CTAs
__host__ __device__ constexpr int seg_cap(int c) {
  return c == 0 ? 128 : (c == 1 ? 256 : (c == 2 ? 512 : (c == 3 ? 1024 : 4096)));
}
constexpr int SEG_MAX = seg_cap(SEG_CLASSES - 1);
constexpr int SEG_BYTES = 40;  // shared memory per box

__global__ void k_heads_u64_shift(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint64_t *__restrict__ head) {
  for (uint64_t i = gtid(); i < n; i += gstride()) head[i] = (i == 0 || (keys[i - 1] >> shift) != (keys[i] >> shift)) ? 1 : 0;
}

// segment -> size class list; cls[c] = segments of class c, cls[SEG_CLASSES] = segments too large
// min_class > 0 pushes small segments into a larger class (tests run every kernel variant on small data)
__global__ void k_seg_classify(const uint32_t *__restrict__ begins, uint64_t n_seg, uint32_t *__restrict__ lists,
                               unsigned int *__restrict__ cls, int min_class) {
  for (uint64_t g = gtid(); g < n_seg; g += gstride()) {
    const uint32_t n = begins[g + 1] - begins[g];
    int c = min_class;
    while (c < SEG_CLASSES && n > (uint32_t)seg_cap(c)) c++;
    const unsigned int k = atomicAdd(&cls[c], 1u);
    if (c < SEG_CLASSES) lists[(uint64_t)c * n_seg + k] = (uint32_t)g;
  }
}

__device__ __forceinline__ uint16_t seg_find(uint16_t *p, uint16_t x) {
  while (p[x] != x) {
    p[x] = p[p[x]];
    x = p[x];
  }
  return x;
}

__device__ __forceinline__ int warp_incl_max(int v) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(FULL, v, d);
    if (lane >= (unsigned)d) v = max(v, o);
  }
  return v;
}

// Shared-memory view of one segment (40 bytes per box).
struct SegMem {
  uint64_t *ord;    // reference-order ordinal
  uint64_t *skey;   // stage-B sort keys, then the sorted (start, end) pairs
  int32_t *qlo, *qhi, *tlo, *thi;
  uint32_t *tid;
  uint16_t *parent; // union-find of stage A, then per sorted position: box index | forward << 15
  uint16_t *heads;  // group heads with more than one member
};

template <int T, int CAP>
__global__ void __launch_bounds__(T == 32 ? 256 : T)
    k_merge_segments(const BoxSrc src, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ begins,
                     const uint32_t *__restrict__ list, uint32_t n_list, int64_t d, int merge_strands,
                     BoxD *__restrict__ swept, uint32_t *__restrict__ out_cnt) {
  extern __shared__ __align__(16) unsigned char seg_smem[];
  constexpr int GROUPS = (T == 32) ? 8 : 1;  // segments in flight per CTA
  const int gi = (T == 32) ? (int)(threadIdx.x >> 5) : 0;
  const int lt = (T == 32) ? (int)(threadIdx.x & 31u) : (int)threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  unsigned char *base = seg_smem + (size_t)gi * CAP * SEG_BYTES;
  SegMem m;
  m.ord = reinterpret_cast<uint64_t *>(base);
  m.skey = m.ord + CAP;
  m.qlo = reinterpret_cast<int32_t *>(m.skey + CAP);
  m.qhi = m.qlo + CAP; m.tlo = m.qhi + CAP; m.thi = m.tlo + CAP;
  m.tid = reinterpret_cast<uint32_t *>(m.thi + CAP);
  m.parent = reinterpret_cast<uint16_t *>(m.tid + CAP);
  m.heads = m.parent + CAP;
  __shared__ unsigned int s_cnt[GROUPS][2];  // [0] roots, [1] multi-member group heads
  __shared__ uint64_t s_rk[(T == 32) ? 1 : T / 32];  // cross-warp argmax of the prefilter
  auto sync = [&]() {
    if (T == 32) __syncwarp();
    else __syncthreads();
  };
  const int32_t md = (int32_t)d;

  for (uint32_t li = blockIdx.x * GROUPS + gi; li < n_list; li += gridDim.x * GROUPS) {
    const uint32_t g = list[li];
    const uint32_t b = begins[g], n = begins[g + 1] - b;
    const BoxD first = load_box(src, idx[b]);
    if (n == 1) {
      if (lt == 0) {
        swept[b] = first;
        out_cnt[g] = 1;
      }
      continue;
    }
    // gather the boxes of the segment, four independent loads in flight per thread
    for (uint32_t i0 = lt; i0 < n; i0 += 4 * T) {
      BoxD x[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t i = i0 + u * T;
        if (i < n) x[u] = load_box(src, idx[b + i]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t i = i0 + u * T;
        if (i < n) {
          m.ord[i] = x[u].ord;
          m.qlo[i] = x[u].q_lo; m.qhi[i] = x[u].q_hi; m.tlo[i] = x[u].t_lo; m.thi[i] = x[u].t_hi;
          m.tid[i] = x[u].t_id;
          m.parent[i] = (uint16_t)i;
        }
      }
    }
    if (lt == 0) s_cnt[gi][0] = s_cnt[gi][1] = 0;
    sync();
    // ---- stage A: (t, strand) groups are contiguous (the global sort key ends with t, strand);
    // heads of groups with more than one member are listed first so that every thread gets one
    if (d >= 0) {
      for (uint32_t i = lt; i + 1 < n; i += T) {
        const uint32_t t = m.tid[i];
        const bool fwd = m.qlo[i] <= m.qhi[i];
        const bool head = i == 0 || m.tid[i - 1] != t || (m.qlo[i - 1] <= m.qhi[i - 1]) != fwd;
        if (head && m.tid[i + 1] == t && (m.qlo[i + 1] <= m.qhi[i + 1]) == fwd)
          m.heads[atomicAdd(&s_cnt[gi][1], 1u)] = (uint16_t)i;
      }
      sync();
      const uint32_t nh = s_cnt[gi][1];
      for (uint32_t h = lt; h < nh; h += T) {
        const uint32_t i = m.heads[h];
        const uint32_t t = m.tid[i];
        const bool fwd = m.qlo[i] <= m.qhi[i];
        uint32_t e = i + 2;
        while (e < n && m.tid[e] == t && (m.qlo[e] <= m.qhi[e]) == fwd) e++;
        // pairwise relation of src/main.rs:12895-12946 on the ORIGINAL coordinates
        // (see k_merge2d_direct); `A` is the member with the smaller (sort key, ord)
        for (uint32_t x = i; x < e; x++) {
          const int64_t kx = fwd ? (int64_t)m.qlo[x] : -(int64_t)m.qlo[x];
          for (uint32_t y = x + 1; y < e; y++) {
            const int64_t ky = fwd ? (int64_t)m.qlo[y] : -(int64_t)m.qlo[y];
            const bool x_first = kx < ky || (kx == ky && m.ord[x] < m.ord[y]);
            const uint32_t A = x_first ? x : y, B = x_first ? y : x;
            const int64_t qa_start = fwd ? m.qlo[A] : m.qhi[A], qa_end = fwd ? m.qhi[A] : m.qlo[A];
            const int64_t qb_start = fwd ? m.qlo[B] : m.qhi[B];
            if (qb_start < qa_start) continue;
            if (qb_start - qa_end > d) continue;
            int64_t t_gap;
            bool t_forward;
            if (fwd) {
              t_gap = (int64_t)m.tlo[B] - m.thi[A];
              t_forward = m.tlo[B] > m.tlo[A];
            } else {
              t_gap = (int64_t)m.tlo[A] - m.thi[B];
              t_forward = m.thi[B] < m.thi[A];
            }
            if (!t_forward || t_gap > d) continue;
            const uint16_t ra = seg_find(m.parent, (uint16_t)x), rb = seg_find(m.parent, (uint16_t)y);
            if (ra != rb) m.parent[ra] = rb;
          }
        }
      }
      sync();
      // merged box of a component = min/max over its members (+ the earliest ord), accumulated into the
      // root's slot AFTER every pair was tested on original coordinates; non-root slots stay untouched.
      // One thread per group again: members of a component are folded in sequence.
      for (uint32_t h = lt; h < nh; h += T) {
        const uint32_t i = m.heads[h];
        const uint32_t t = m.tid[i];
        const bool fwd = m.qlo[i] <= m.qhi[i];
        for (uint32_t x = i; x < n && m.tid[x] == t && ((m.qlo[x] <= m.qhi[x]) == fwd); x++) {
          const uint16_t r = seg_find(m.parent, (uint16_t)x);
          if (r == x) continue;
          if (fwd) {
            m.qlo[r] = min(m.qlo[r], m.qlo[x]);
            m.qhi[r] = max(m.qhi[r], m.qhi[x]);
          } else {
            m.qlo[r] = max(m.qlo[r], m.qlo[x]);
            m.qhi[r] = min(m.qhi[r], m.qhi[x]);
          }
          m.tlo[r] = min(m.tlo[r], m.tlo[x]);
          m.thi[r] = max(m.thi[r], m.thi[x]);
          m.ord[r] = min(m.ord[r], m.ord[x]);
        }
      }
      sync();
    }
    // ---- roots -> list of box indices (the union-find is done: `heads` and `parent` become scratch)
    for (uint32_t i = lt; i < n; i += T)
      if (m.parent[i] == i) m.heads[atomicAdd(&s_cnt[gi][0], 1u)] = (uint16_t)i;
    sync();
    uint32_t nr = s_cnt[gi][0];
    uint16_t *live = m.heads, *spare = m.parent;
    // ---- prefilter: drop the boxes the sweep of src/main.rs:12496-12556 cannot see. In the sorted order
    // (start, !forward, input order) a box D that comes after a box J with end(J) >= end(D) never opens a
    // row (its start lies inside the span merged so far), never extends the span and is never "longer than
    // the span merged before it" (that span contains it), so with merge_strands and d >= 0 it changes
    // neither the rows nor their orientation. The hits of one (row, q) segment are mostly near-copies of one
    // interval, so the LONGEST box contains about half of the others: each round takes the longest
    // remaining box as the pivot J (an argmax, no sort), keeps it, and drops what it dominates; the
    // bitonic network below then sorts a third of the compare-exchanges it would otherwise need.
    if (merge_strands && md >= 0) {
      constexpr uint32_t STOP = (T == 32) ? 32u : 64u;  // short enough for a quick sort
      uint32_t n_piv = 0;
      while (nr - n_piv > STOP && n_piv < 6) {
        uint64_t best = 0;  // (length + 1) << 16 | box index; 0 = none
        for (uint32_t c = n_piv + lt; c < nr; c += T) {
          const uint32_t i = live[c];
          const uint64_t len = (uint64_t)((int64_t)max(m.qlo[i], m.qhi[i]) - (int64_t)min(m.qlo[i], m.qhi[i])) + 1;
          best = max(best, (len << 16) | i);
        }
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) best = max(best, __shfl_xor_sync(FULL, best, dlt));
        if (T != 32) {
          if (lane == 0) s_rk[threadIdx.x >> 5] = best;
          __syncthreads();
#pragma unroll
          for (int w = 0; w < T / 32; w++) best = max(best, s_rk[w]);
        }
        // every thread holds the pivot; new list = old pivots, the pivot, the boxes it does not dominate
        const uint32_t bi = (uint32_t)(best & 0xffffu);
        const bool pf = m.qlo[bi] <= m.qhi[bi];
        const int32_t pe = pf ? m.qhi[bi] : m.qlo[bi];
        const uint64_t pk = ((uint64_t)(uint32_t)(pf ? m.qlo[bi] : m.qhi[bi]) << 1) | (pf ? 0u : 1u);
        const uint64_t po = m.ord[bi];
        for (uint32_t c = lt; c < n_piv; c += T) spare[c] = live[c];
        if (lt == 0) {
          spare[n_piv] = (uint16_t)bi;
          s_cnt[gi][0] = n_piv + 1;
        }
        sync();
        uint32_t base = n_piv + 1;  // warp classes: the running count lives in a register
        for (uint32_t c0 = n_piv; c0 < nr; c0 += T) {
          const uint32_t c = c0 + lt;
          uint32_t i = 0;
          bool keep = false;
          if (c < nr) {
            i = live[c];
            const bool fwd = m.qlo[i] <= m.qhi[i];
            const uint64_t k = ((uint64_t)(uint32_t)(fwd ? m.qlo[i] : m.qhi[i]) << 1) | (fwd ? 0u : 1u);
            const bool after = k > pk || (k == pk && m.ord[i] > po);  // sorts after the pivot
            keep = i != bi && !(after && (fwd ? m.qhi[i] : m.qlo[i]) <= pe);
          }
          const unsigned bm = __ballot_sync(FULL, keep);
          uint32_t wb;
          if (T == 32) {
            wb = base;
            base += __popc(bm);
          } else {
            wb = 0;
            if (lane == 0 && bm) wb = atomicAdd(&s_cnt[gi][0], (unsigned)__popc(bm));
            wb = __shfl_sync(FULL, wb, 0);
          }
          if (keep) spare[wb + __popc(bm & lanemask_lt())] = (uint16_t)i;
        }
        if (T == 32 && lt == 0) s_cnt[gi][0] = base;
        sync();
        const uint32_t before = nr - n_piv;
        nr = s_cnt[gi][0];
        n_piv++;
        uint16_t *t = live;
        live = spare;
        spare = t;
        if ((before - (nr - n_piv)) * 10 < before) break;  // the pivot dominated next to nothing: stop
      }
    }
    // ---- stage-B sort keys: (start, !forward, box index); equal (start, strand) are put into ord
    // order afterwards (src/main.rs:12481-12494 is a stable sort of the input order)
    for (uint32_t c = lt; c < nr; c += T) {
      const uint32_t i = live[c];
      const bool fwd = m.qlo[i] <= m.qhi[i];
      const uint32_t st = (uint32_t)(fwd ? m.qlo[i] : m.qhi[i]);
      m.skey[c] = ((uint64_t)st << 17) | ((uint64_t)(fwd ? 0u : 1u) << 16) | i;
    }
    sync();
    uint32_t P = 1;
    while (P < nr) P <<= 1;
    for (uint32_t i = nr + lt; i < P; i += T) m.skey[i] = ~0ull;
    sync();
    {
      // Bitonic network. Warp w owns the compare-exchanges [w * ppw, (w + 1) * ppw) of every step;
      // for j <= ppw both elements of its pairs lie in the warp's own 2 * ppw keys, so those steps
      // only need a warp barrier — the CTA barrier is paid for the few steps that cross warps.
      const uint32_t half = P >> 1;
      const uint32_t ppw = (T == 32) ? half : max(32u, half / (uint32_t)(T / 32));
      const uint32_t wbase = (T == 32) ? 0u : (uint32_t)(threadIdx.x >> 5) * ppw;
      // A short list (after the prefilter) keeps only the first warps busy: the others skip the network
      // and wait at the CTA barrier below; the working warps meet at a named barrier of their own.
      const uint32_t n_work = (T == 32) ? 32u : min((uint32_t)T, ((half + ppw - 1) / ppw) * 32u);
      if (T == 32 || wbase < half) {
        for (uint32_t k = 2; k <= P; k <<= 1) {
          for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t r = lane; r < ppw; r += 32) {
              const uint32_t i = wbase + r;  // i-th compare-exchange of this step: lo has bit j clear
              if (i < half) {
                const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                const uint64_t x = m.skey[lo], y = m.skey[hi];
                const bool up = (lo & k) == 0;
                if ((y < x) == up) {
                  m.skey[lo] = y;
                  m.skey[hi] = x;
                }
              }
            }
            const uint32_t j_next = j > 1 ? (j >> 1) : k;  // first step of the next level has j = k
            if (T != 32 && (j > ppw || j_next > ppw)) asm volatile("bar.sync 1, %0;" ::"r"(n_work) : "memory");
            else __syncwarp();
          }
        }
      }
      sync();
    }
    // ---- ties on (start, strand): restore ord order. The thread at the head of a run of equal
    // keys insertion-sorts that run by ord (runs are short and disjoint, so this is parallel).
    for (uint32_t i = lt; i + 1 < nr; i += T) {
      const uint64_t hk = m.skey[i] >> 16;
      if ((m.skey[i + 1] >> 16) != hk || (i > 0 && (m.skey[i - 1] >> 16) == hk)) continue;
      for (uint32_t a = i + 1; a < nr && (m.skey[a] >> 16) == hk; a++) {
        const uint64_t ka = m.skey[a];
        const uint64_t oa = m.ord[ka & 0xffffu];
        uint32_t j = a;
        while (j > i && m.ord[m.skey[j - 1] & 0xffffu] > oa) {
          m.skey[j] = m.skey[j - 1];
          j--;
        }
        m.skey[j] = ka;
      }
    }
    sync();
    // sorted position -> (start, end) pair in place of the key, box index | forward << 15 in `parent`
    for (uint32_t i = lt; i < nr; i += T) {
      const uint32_t x = (uint32_t)(m.skey[i] & 0xffffu);
      const bool fwd = m.qlo[x] <= m.qhi[x];
      const int32_t st = fwd ? m.qlo[x] : m.qhi[x], en = fwd ? m.qhi[x] : m.qlo[x];
      m.skey[i] = ((uint64_t)(uint32_t)en << 32) | (uint32_t)st;
      m.parent[i] = (uint16_t)(x | (fwd ? 0x8000u : 0u));
    }
    sync();
    auto emit = [&](uint32_t w, uint32_t first_pos, int32_t st, int32_t en, bool fwd) {
      const uint32_t c = m.parent[first_pos] & 0x7fffu;
      swept[b + w] = BoxD{fwd ? st : en, fwd ? en : st, m.tlo[c], m.thi[c], first.q_id, m.tid[c], first.row, 1u, m.ord[c]};
    };
    if (merge_strands) {
      // ---- the sweep of src/main.rs:12496-12556 as scans over the sorted boxes (first warp):
      //   a box starts a new output row iff its start exceeds every earlier end by more than d;
      //   the row spans [start of its first box, max end]; its orientation is that of the last
      //   box that was longer than the span merged before it (else of the first box).
      if (threadIdx.x < 32 || T == 32) {
        int carry_pm = INT32_MIN, carry_lb = -1, carry_os = -1;
        uint32_t carry_cnt = 0;
        for (uint32_t k0 = 0; k0 < nr; k0 += 32) {
          const uint32_t k = k0 + lane;
          const bool valid = k < nr;
          const uint64_t se = valid ? m.skey[k] : 0ull;
          const int32_t st = (int32_t)(uint32_t)se, en = (int32_t)(uint32_t)(se >> 32);
          const int pm_in = max(carry_pm, warp_incl_max(valid ? en : INT32_MIN));
          int pm_ex = __shfl_up_sync(FULL, pm_in, 1);
          if (lane == 0) pm_ex = carry_pm;
          const bool brk = valid && (k == 0 || md < 0 || (int64_t)st > (int64_t)pm_ex + md);
          const int lb_in = max(carry_lb, warp_incl_max(brk ? (int)k : -1));
          int lb_ex = __shfl_up_sync(FULL, lb_in, 1);
          if (lane == 0) lb_ex = carry_lb;
          bool cand = false;
          if (valid && !brk) {
            const int32_t s0 = (int32_t)(uint32_t)m.skey[lb_in];
            cand = ((int64_t)en - st) > ((int64_t)pm_ex - s0);
          }
          const int os_in = max(carry_os, warp_incl_max((brk || cand) ? (int)k : -1));
          int os_ex = __shfl_up_sync(FULL, os_in, 1);
          if (lane == 0) os_ex = carry_os;
          const unsigned bm = __ballot_sync(FULL, brk);
          const uint32_t cnt_ex = carry_cnt + __popc(bm & lanemask_lt());
          if (brk && k > 0) {  // the row that ended just before this box
            const uint64_t fe = m.skey[lb_ex];  // md < 0: nothing merges, a row is its own box
            emit(cnt_ex - 1, (uint32_t)lb_ex, (int32_t)(uint32_t)fe, md < 0 ? (int32_t)(uint32_t)(fe >> 32) : pm_ex,
                 (m.parent[os_ex] & 0x8000u) != 0);
          }
          carry_pm = __shfl_sync(FULL, pm_in, 31);
          carry_lb = __shfl_sync(FULL, lb_in, 31);
          carry_os = __shfl_sync(FULL, os_in, 31);
          carry_cnt += __popc(bm);
        }
        if (lane == 0) {
          const uint64_t fe = m.skey[carry_lb];
          emit(carry_cnt - 1, (uint32_t)carry_lb, (int32_t)(uint32_t)fe, md < 0 ? (int32_t)(uint32_t)(fe >> 32) : carry_pm,
               (m.parent[carry_os] & 0x8000u) != 0);
          out_cnt[g] = carry_cnt;
        }
      }
    } else if (lt == 0) {
      // --consider-strandness: the literal sequential sweep (a strand change always starts a new row)
      uint32_t w = 0, cpos = 0;
      uint64_t se = m.skey[0];
      int32_t cs = (int32_t)(uint32_t)se, ce = (int32_t)(uint32_t)(se >> 32);
      bool cf = (m.parent[0] & 0x8000u) != 0;
      for (uint32_t rd = 1; rd < nr; rd++) {
        se = m.skey[rd];
        const int32_t ns = (int32_t)(uint32_t)se, ne = (int32_t)(uint32_t)(se >> 32);
        const bool nf = (m.parent[rd] & 0x8000u) != 0;
        if (md < 0 || cf != nf || (int64_t)ns > (int64_t)ce + md) {
          emit(w++, cpos, cs, ce, cf);
          cpos = rd; cs = ns; ce = ne; cf = nf;
        } else {
          cs = min(cs, ns);
          ce = max(ce, ne);
        }
      }
      emit(w++, cpos, cs, ce, cf);
      out_cnt[g] = w;
    }
    sync();  // the next segment reuses the shared arrays
  }
}

}  // namespace impgx
