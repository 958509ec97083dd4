// bucket_kernels.cuh — the BED merge of the direct path without a global sort.
//
// output_results_bed (src/main.rs:11849-11866) groups every result of a row by its query sequence:
// merge_adjusted_intervals_gap_2d (:12858-13011) works inside (q, t, strand) groups,
// merge_query_adjusted_intervals (:12474-12560) sweeps the boxes of one q. So the unit of work
// is the (row, q) SEGMENT. Instead of sorting all boxes of a batch by (row, q, t, strand), the hop
// that produces them counts them per (row, q) in a dense table (rows x n_seqs, k_stab_count<BUCKET>),
// an exclusive scan turns the counts into bucket offsets, and the liftover epilogue writes each
// accepted hit as a 32-byte BoxRec straight into its bucket (slot = atomic cursor). A warp / CTA
// then owns one bucket: it reads contiguous memory once, groups by (t, strand) with a shared-memory
// hash table (stage A), and runs the same prefilter / bitonic sort / scan sweep as k_merge_segments
// (stage B). The order of the boxes inside a bucket is arbitrary (atomic slots); every step below
// is a function of the box SET (ties are broken by the reference-order ordinal), so the output is
// deterministic. A bucket beyond SEG_MAX boxes — and only that bucket — goes through the global
// two-sort path (engine.cu, merge_oversized).
#pragma once
#include "merge_kernels.cuh"

namespace impgx {

constexpr int BK_BYTES = 46;  // shared memory per box
constexpr uint32_t NIL32 = 0xffffffffu;
constexpr uint32_t NIL16 = 0xffffu;
constexpr uint32_t SKIP16 = 0xfffeu;  // in `nxt` before stage A: the box is a stage-A result already, it joins no group

// ---- boxes that already exist as BoxD records (seeds, the hops that were ordered for the fold,
// boxes received from peer ranks) join the buckets through a count and a scatter pass
__global__ void k_bucket_count_boxd(const BoxD *__restrict__ b, uint64_t n, uint32_t n_seqs, uint32_t *__restrict__ cnt) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const BoxD x = b[i];
    if (x.valid) atomicAdd(cnt + (uint64_t)x.row * n_seqs + x.q_id, 1u);
  }
}
__global__ void k_bucket_scatter_boxd(const BoxD *__restrict__ b, uint64_t n, uint32_t n_seqs, uint32_t *__restrict__ cursor,
                                      BoxRec *__restrict__ boxes) {
  for (uint64_t i = gtid(); i < n; i += gstride()) {
    const BoxD x = b[i];
    if (!x.valid) continue;
    const uint32_t slot = atomicAdd(cursor + (uint64_t)x.row * n_seqs + x.q_id, 1u);
    uint4 *dst = reinterpret_cast<uint4 *>(boxes + slot);
    dst[0] = make_uint4((uint32_t)x.ord, (uint32_t)(x.ord >> 32), (uint32_t)x.q_lo, (uint32_t)x.q_hi);
    dst[1] = make_uint4(x.t_id, (uint32_t)x.t_lo, (uint32_t)x.t_hi, (x.valid & BOXD_MERGED_A) ? BOX_MERGED_A : 0u);
  }
}

__device__ __forceinline__ void store_seg_out(BoxRec *slot, int32_t qf, int32_t ql, uint32_t tid, int32_t tf, int32_t tl) {
  SegOut *o = reinterpret_cast<SegOut *>(slot);
  o->q_first = qf; o->q_last = ql; o->t_id = tid; o->t_first = tf; o->t_last = tl;
}

constexpr int TINY_MAX = 4;                 // buckets of up to this many boxes: one THREAD per bucket (k_merge_tiny)
constexpr int TINY_CLASS = SEG_CLASSES + 1;  // list / counter index of the tiny buckets

// bucket -> size class list; cls[c] = buckets of class c, cls[SEG_CLASSES] = buckets beyond SEG_MAX
// (listed too, lists[SEG_CLASSES * cap ...]), cls[TINY_CLASS] = buckets of <= TINY_MAX boxes; min_class > 0 pushes
// small buckets into a larger class (tests run every kernel variant on small data)
__global__ void k_bucket_classify(const uint32_t *__restrict__ beg, const uint32_t *__restrict__ cur, uint64_t n_buckets,
                                  uint32_t *__restrict__ lists, uint64_t cap, unsigned int *__restrict__ cls, int min_class) {
  for (uint64_t b = gtid(); b < n_buckets; b += gstride()) {
    const uint32_t n = cur[b] - beg[b];
    if (n == 0) continue;
    int c = min_class;
    if (min_class == 0 && n <= (uint32_t)TINY_MAX) c = TINY_CLASS;
    else
      while (c < SEG_CLASSES && n > (uint32_t)seg_cap(c)) c++;
    const unsigned int k = atomicAdd(&cls[c], 1u);
    lists[(uint64_t)c * cap + k] = (uint32_t)b;
  }
}

// Buckets of at most TINY_MAX boxes (depth-1 queries: the two alignments of a genome pair; sparse worlds), one thread
// each, everything in registers: stage A (pairwise relation on the original coordinates, then the hulls), the sort by
// (start, !forward, ordinal) and the literal sweep of src/main.rs:12496-12556 — or, in reduce mode, the boxes whose end
// exceeds every earlier end. Same semantics as k_merge_buckets, which spends a warp on such a bucket.
__global__ void __launch_bounds__(128) k_merge_tiny(BoxRec *__restrict__ boxes, const uint32_t *__restrict__ beg,
                                                    const uint32_t *__restrict__ cur, const uint32_t *__restrict__ list,
                                                    uint32_t n_list, int64_t d, int merge_strands, int reduce,
                                                    uint32_t *__restrict__ out_cnt,
                                                    const unsigned int *__restrict__ n_list_dev) {
  if (n_list_dev) n_list = *n_list_dev;  // small_bfs.cuh: the list was built by the launch before, nobody read its length
  for (uint64_t li = gtid(); li < n_list; li += gstride()) {
    const uint32_t bk = list[li];
    const uint32_t b = beg[bk];
    const int n = (int)(cur[bk] - b);
    BoxRec *seg = boxes + b;
    BoxRec x[TINY_MAX];
#pragma unroll
    for (int i = 0; i < TINY_MAX; i++)
      if (i < n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(seg + i);
        const uint4 a = src[0], c = src[1];
        x[i].ord = ((uint64_t)a.y << 32) | a.x;
        x[i].q_first = (int32_t)a.z; x[i].q_last = (int32_t)a.w;
        x[i].t_id = c.x; x[i].t_first = (int32_t)c.y; x[i].t_last = (int32_t)c.z; x[i].flags = c.w;
      }
    // ---- stage A: unions from the ORIGINAL coordinates of every pair of one (t, strand) group, then the hulls
    int parent[TINY_MAX] = {0, 1, 2, 3};
    if (d >= 0) {
#pragma unroll
      for (int i = 0; i < TINY_MAX; i++)
#pragma unroll
        for (int j = i + 1; j < TINY_MAX; j++) {
          if (j >= n) continue;
          const bool fi = x[i].q_first <= x[i].q_last, fj = x[j].q_first <= x[j].q_last;
          if (x[i].t_id != x[j].t_id || fi != fj || ((x[i].flags | x[j].flags) & BOX_MERGED_A)) continue;
          const int64_t ki = fi ? (int64_t)x[i].q_first : -(int64_t)x[i].q_first;
          const int64_t kj = fi ? (int64_t)x[j].q_first : -(int64_t)x[j].q_first;
          const bool i_first = ki < kj || (ki == kj && x[i].ord < x[j].ord);
          const BoxRec &A = i_first ? x[i] : x[j];
          const BoxRec &B = i_first ? x[j] : x[i];
          const int64_t qa_start = fi ? A.q_first : A.q_last, qa_end = fi ? A.q_last : A.q_first;
          const int64_t qb_start = fi ? B.q_first : B.q_last;
          if (qb_start < qa_start || qb_start - qa_end > d) continue;
          const int64_t t_gap = fi ? (int64_t)B.t_first - A.t_last : (int64_t)A.t_first - B.t_last;
          const bool t_forward = fi ? B.t_first > A.t_first : B.t_last < A.t_last;
          if (!t_forward || t_gap > d) continue;
          int ri = i, rj = j;
          while (parent[ri] != ri) ri = parent[ri];
          while (parent[rj] != rj) rj = parent[rj];
          if (ri != rj) parent[ri] = rj;
        }
    }
    BoxRec r[TINY_MAX];
    bool is_root[TINY_MAX];
#pragma unroll
    for (int i = 0; i < TINY_MAX; i++) {
      is_root[i] = i < n && parent[i] == i;
      if (i < n) r[i] = x[i];
    }
#pragma unroll
    for (int i = 0; i < TINY_MAX; i++) {
      if (i >= n || parent[i] == i) continue;
      int rt = i;
      while (parent[rt] != rt) rt = parent[rt];
      const bool fwd = x[i].q_first <= x[i].q_last;
#pragma unroll
      for (int k = 0; k < TINY_MAX; k++)  // r[rt] with a compile-time index
        if (k == rt) {
          r[k].q_first = fwd ? min(r[k].q_first, x[i].q_first) : max(r[k].q_first, x[i].q_first);
          r[k].q_last = fwd ? max(r[k].q_last, x[i].q_last) : min(r[k].q_last, x[i].q_last);
          r[k].t_first = min(r[k].t_first, x[i].t_first);
          r[k].t_last = max(r[k].t_last, x[i].t_last);
          r[k].ord = min(r[k].ord, x[i].ord);
        }
    }
    // ---- the roots in the order of the sweep: (start, !forward, ordinal)
    int ordv[TINY_MAX];
    int nr = 0;
#pragma unroll
    for (int i = 0; i < TINY_MAX; i++)
      if (is_root[i]) {
        const bool fi = r[i].q_first <= r[i].q_last;
        const int32_t si = fi ? r[i].q_first : r[i].q_last;
        int pos = 0;  // number of roots that sort before root i
#pragma unroll
        for (int j = 0; j < TINY_MAX; j++)
          if (j != i && is_root[j]) {
            const bool fj = r[j].q_first <= r[j].q_last;
            const int32_t sj = fj ? r[j].q_first : r[j].q_last;
            const bool before = sj < si || (sj == si && ((fj && !fi) || (fj == fi && r[j].ord < r[i].ord)));
            pos += before ? 1 : 0;
          }
#pragma unroll
        for (int k = 0; k < TINY_MAX; k++)
          if (k == pos) ordv[k] = i;
        nr++;
      }
    auto get = [&](int k) -> BoxRec {  // root at sorted position k
      BoxRec o = r[0];
#pragma unroll
      for (int i = 1; i < TINY_MAX; i++)
        if (ordv[k] == i) o = r[i];
      return o;
    };
    uint32_t w = 0;
    const int32_t md = (int32_t)d;
    if (reduce) {
      const bool prune = merge_strands && md >= 0;
      int32_t pm = INT32_MIN;
      for (int k = 0; k < nr; k++) {
        const BoxRec o = get(k);
        const int32_t en = max(o.q_first, o.q_last);
        if (!prune || k == 0 || en > pm) {
          uint4 *dst = reinterpret_cast<uint4 *>(seg + w);
          dst[0] = make_uint4((uint32_t)o.ord, (uint32_t)(o.ord >> 32), (uint32_t)o.q_first, (uint32_t)o.q_last);
          dst[1] = make_uint4(o.t_id, (uint32_t)o.t_first, (uint32_t)o.t_last, BOX_MERGED_A);
          w++;
        }
        pm = max(pm, en);
      }
    } else {
      BoxRec c0 = get(0);
      bool cf = c0.q_first <= c0.q_last;
      int32_t cs = cf ? c0.q_first : c0.q_last, ce = cf ? c0.q_last : c0.q_first;
      uint32_t ct = c0.t_id;
      int32_t ctf = c0.t_first, ctl = c0.t_last;
      for (int k = 1; k < nr; k++) {
        const BoxRec nx = get(k);
        const bool nf = nx.q_first <= nx.q_last;
        const int32_t ns = nf ? nx.q_first : nx.q_last, ne = nf ? nx.q_last : nx.q_first;
        if (md < 0 || (!merge_strands && cf != nf) || (int64_t)ns > (int64_t)ce + md) {
          store_seg_out(seg + w, cf ? cs : ce, cf ? ce : cs, ct, ctf, ctl);
          w++;
          cs = ns; ce = ne; cf = nf; ct = nx.t_id; ctf = nx.t_first; ctl = nx.t_last;
        } else {
          if (merge_strands && cf != nf && ((int64_t)ne - ns) > ((int64_t)ce - cs)) cf = nf;
          cs = min(cs, ns);
          ce = max(ce, ne);
        }
      }
      store_seg_out(seg + w, cf ? cs : ce, cf ? ce : cs, ct, ctf, ctl);
      w++;
    }
    out_cnt[bk] = w;
  }
}

// Shared-memory view of one bucket (46 bytes per box).
struct BkMem {
  uint64_t *ord;    // reference-order ordinal
  uint64_t *skey;   // stage A: hash table of (t, strand) groups (2 x u32 slots per box); then the stage-B sort keys
  int32_t *qlo, *qhi, *tlo, *thi;
  uint32_t *tid;
  uint32_t *chain;  // per group representative: the member inserted last (NIL32: not a representative)
  uint16_t *parent; // union-find of stage A, then per sorted position: box index | forward << 15
  uint16_t *heads;  // representatives of groups with more than one member; then the list of roots
  uint16_t *nxt;    // next member of the (t, strand) group (NIL16: none)
};

// One (row, q) bucket per warp (T == 32, eight buckets in flight per CTA) or per CTA. The merged BED rows
// of bucket `bk` are staged over its own first slots (SegOut records at a BoxRec stride), out_cnt[bk] rows.
//
// A box flagged BOX_MERGED_A is a stage-A result already (it was merged where it was produced; testing the
//   pairwise relation again on merged boxes would union components that no original pair connects): it joins
//   no (t, strand) group here. The members of one group are all flagged or all raw, because a group is
//   produced by one rank (the owner of t) in one bucket.
// reduce == 1 (sharded index, on the rank that PRODUCED the boxes): stage A, then instead of the sweep only
//   the boxes a sweep over ANY superset of this bucket can still see are kept, as BoxRecs over the bucket's
//   first slots: in the sweep's sort order (start, !forward, ord) a box D behind a box J with end(J) >= end(D)
//   never opens a row, never extends one and is never longer than the span merged before it, whatever other
//   boxes the owner of the query sequence receives from the other ranks. What survives is the staircase of
//   strictly growing ends, a handful of boxes per bucket; it travels instead of every hit.
template <int T, int CAP>
__global__ void __launch_bounds__(T == 32 ? 256 : T)
    k_merge_buckets(BoxRec *__restrict__ boxes, const uint32_t *__restrict__ beg, const uint32_t *__restrict__ cur,
                    const uint32_t *__restrict__ list, uint32_t n_list, int64_t d, int merge_strands, int reduce,
                    uint32_t *__restrict__ out_cnt, const unsigned int *__restrict__ n_list_dev) {
  extern __shared__ __align__(16) unsigned char seg_smem[];
  if (n_list_dev) n_list = *n_list_dev;  // small_bfs.cuh: list length left on the device
  constexpr int GROUPS = (T == 32) ? 8 : 1;  // buckets in flight per CTA
  const int gi = (T == 32) ? (int)(threadIdx.x >> 5) : 0;
  const int lt = (T == 32) ? (int)(threadIdx.x & 31u) : (int)threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  unsigned char *base = seg_smem + (size_t)gi * CAP * BK_BYTES;
  BkMem m;
  m.ord = reinterpret_cast<uint64_t *>(base);
  m.skey = m.ord + CAP;
  m.qlo = reinterpret_cast<int32_t *>(m.skey + CAP);
  m.qhi = m.qlo + CAP; m.tlo = m.qhi + CAP; m.thi = m.tlo + CAP;
  m.tid = reinterpret_cast<uint32_t *>(m.thi + CAP);
  m.chain = m.tid + CAP;
  m.parent = reinterpret_cast<uint16_t *>(m.chain + CAP);
  m.heads = m.parent + CAP;
  m.nxt = m.heads + CAP;
  __shared__ unsigned int s_cnt[GROUPS][2];  // [0] roots, [1] multi-member groups
  __shared__ uint64_t s_rk[(T == 32) ? 1 : T / 32];  // cross-warp argmax of the prefilter
  constexpr int BINS = 64;                        // start bins of the staircase filter
  __shared__ int s_bin[GROUPS][BINS];             // max end per start bin, then the exclusive prefix max
  __shared__ int s_mm[(T == 32) ? 1 : T / 32][2];  // cross-warp min / max start
  auto sync = [&]() {
    if (T == 32) __syncwarp();
    else __syncthreads();
  };
  const int32_t md = (int32_t)d;

  for (uint32_t li = blockIdx.x * GROUPS + gi; li < n_list; li += gridDim.x * GROUPS) {
    const uint32_t bk = list[li];
    const uint32_t b = beg[bk], n = cur[bk] - b;
    BoxRec *seg = boxes + b;
    if (n == 1) {
      if (lt == 0 && reduce) {  // the box stays as it is: a group of one
        seg->flags = BOX_MERGED_A;
        out_cnt[bk] = 1;
      }
      if (lt == 0 && !reduce) {
        const uint4 *src = reinterpret_cast<const uint4 *>(seg);
        const uint4 a = src[0], c = src[1];
        store_seg_out(seg, (int32_t)a.z, (int32_t)a.w, c.x, (int32_t)c.y, (int32_t)c.z);
        out_cnt[bk] = 1;
      }
      continue;
    }
    // the boxes of the bucket: contiguous 32-byte records, four independent pairs of 128-bit loads in flight per thread
    for (uint32_t i0 = lt; i0 < n; i0 += 4 * T) {
      uint4 a[4], c[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t i = i0 + u * T;
        if (i < n) {
          const uint4 *src = reinterpret_cast<const uint4 *>(seg + i);
          a[u] = src[0];
          c[u] = src[1];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t i = i0 + u * T;
        if (i < n) {
          m.ord[i] = ((uint64_t)a[u].y << 32) | a[u].x;
          m.qlo[i] = (int32_t)a[u].z; m.qhi[i] = (int32_t)a[u].w;
          m.tid[i] = c[u].x; m.tlo[i] = (int32_t)c[u].y; m.thi[i] = (int32_t)c[u].z;
          m.parent[i] = (uint16_t)i;
          m.chain[i] = NIL32;
          m.nxt[i] = (uint16_t)((c[u].w & BOX_MERGED_A) ? SKIP16 : NIL16);
        }
      }
    }
    if (lt == 0) s_cnt[gi][0] = s_cnt[gi][1] = 0;
    // ---- stage A: (t, strand) groups through a hash table (open addressing, the first box inserted under a key
    // represents its group; every member is chained behind the representative)
    if (d >= 0) {
      uint32_t TS = 2;
      while (TS < 2 * n) TS <<= 1;  // <= 2 * CAP slots: load factor <= 1/2
      uint32_t *tab = reinterpret_cast<uint32_t *>(m.skey);
      for (uint32_t i = lt; i < TS; i += T) tab[i] = 0u;
      sync();
      for (uint32_t i = lt; i < n; i += T) {
        if (m.nxt[i] == SKIP16) continue;
        const uint32_t t = m.tid[i];
        const bool fwd = m.qlo[i] <= m.qhi[i];
        uint32_t h = ((((t << 1) | (fwd ? 1u : 0u)) * 2654435761u) >> 7) & (TS - 1);
        uint32_t rep;
        for (;;) {
          const uint32_t old = atomicCAS(&tab[h], 0u, i + 1);
          if (old == 0u) {
            rep = i;
            break;
          }
          const uint32_t j = old - 1;
          if (m.tid[j] == t && (m.qlo[j] <= m.qhi[j]) == fwd) {
            rep = j;
            break;
          }
          h = (h + 1) & (TS - 1);
        }
        m.nxt[i] = (uint16_t)atomicExch(&m.chain[rep], i);  // NIL32 truncates to NIL16
      }
      sync();
      for (uint32_t i0 = 0; i0 < n; i0 += T) {  // warp-aggregated append
        const uint32_t i = i0 + lt;
        bool is_head = false;
        if (i < n) {
          const uint32_t c = m.chain[i];
          is_head = c != NIL32 && m.nxt[c] != NIL16;
        }
        const unsigned bm = __ballot_sync(FULL, is_head);
        if (bm) {
          uint32_t wb = 0;
          if (lane == (unsigned)(__ffs(bm) - 1)) wb = atomicAdd(&s_cnt[gi][1], (unsigned)__popc(bm));
          wb = __shfl_sync(FULL, wb, __ffs(bm) - 1);
          if (is_head) m.heads[wb + __popc(bm & lanemask_lt())] = (uint16_t)i;
        }
      }
      sync();
      const uint32_t nh = s_cnt[gi][1];
      for (uint32_t h = lt; h < nh; h += T) {
        const uint32_t r = m.heads[h];
        const bool fwd = m.qlo[r] <= m.qhi[r];
        // pairwise relation of src/main.rs:12895-12946 on the ORIGINAL coordinates (see k_merge2d_direct):
        // it is a property of the pair; `A` is the member with the smaller (sort key, ord)
        for (uint32_t x = m.chain[r]; x != NIL16; x = m.nxt[x]) {
          const int64_t kx = fwd ? (int64_t)m.qlo[x] : -(int64_t)m.qlo[x];
          for (uint32_t y = m.nxt[x]; y != NIL16; y = m.nxt[y]) {
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
      // root's slot AFTER every pair was tested on original coordinates; one thread per group again
      for (uint32_t h = lt; h < nh; h += T) {
        const uint32_t r = m.heads[h];
        const bool fwd = m.qlo[r] <= m.qhi[r];
        for (uint32_t x = m.chain[r]; x != NIL16; x = m.nxt[x]) {
          const uint16_t rt = seg_find(m.parent, (uint16_t)x);
          if (rt == x) continue;
          if (fwd) {
            m.qlo[rt] = min(m.qlo[rt], m.qlo[x]);
            m.qhi[rt] = max(m.qhi[rt], m.qhi[x]);
          } else {
            m.qlo[rt] = max(m.qlo[rt], m.qlo[x]);
            m.qhi[rt] = min(m.qhi[rt], m.qhi[x]);
          }
          m.tlo[rt] = min(m.tlo[rt], m.tlo[x]);
          m.thi[rt] = max(m.thi[rt], m.thi[x]);
          m.ord[rt] = min(m.ord[rt], m.ord[x]);
        }
      }
    }
    sync();
    // ---- roots -> list of box indices (the union-find is done: `heads` and `parent` become scratch)
    for (uint32_t i0 = 0; i0 < n; i0 += T) {  // warp-aggregated append
      const uint32_t i = i0 + lt;
      const bool root = i < n && m.parent[i] == i;
      const unsigned bm = __ballot_sync(FULL, root);
      if (bm) {
        uint32_t wb = 0;
        if (lane == (unsigned)(__ffs(bm) - 1)) wb = atomicAdd(&s_cnt[gi][0], (unsigned)__popc(bm));
        wb = __shfl_sync(FULL, wb, __ffs(bm) - 1);
        if (root) m.heads[wb + __popc(bm & lanemask_lt())] = (uint16_t)i;
      }
    }
    sync();
    uint32_t nr = s_cnt[gi][0];
    uint16_t *live = m.heads, *spare = m.parent;
    // ---- prefilter (see k_merge_segments): drop the boxes the sweep of src/main.rs:12496-12556 cannot see
    if (merge_strands && md >= 0) {
      constexpr uint32_t STOP = (T == 32) ? 32u : 64u;  // short enough for a quick sort
      // Staircase filter. The hits of a (row, q) bucket are mostly SHIFTED copies of one interval (start and end
      // move together), which the longest-box pivots below do not dominate. Bin the boxes by start, take the max
      // end per bin and its exclusive prefix max over the bins: a box whose end does not exceed the max end of the
      // EARLIER bins lies behind a box that starts strictly before it and ends no earlier — the sweep cannot see it
      // (same argument as for the pivots). What is left is little more than the staircase of growing ends.
      if (T != 32 && nr > 2 * STOP) {  // the warp classes (<= 256 boxes) go straight to the pivots
        int smin = INT32_MAX, smax = INT32_MIN;
        for (uint32_t c = lt; c < nr; c += T) {
          const uint32_t i = live[c];
          const int st = min(m.qlo[i], m.qhi[i]);
          smin = min(smin, st);
          smax = max(smax, st);
        }
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) {
          smin = min(smin, __shfl_xor_sync(FULL, smin, dlt));
          smax = max(smax, __shfl_xor_sync(FULL, smax, dlt));
        }
        if (T != 32) {
          if (lane == 0) {
            s_mm[threadIdx.x >> 5][0] = smin;
            s_mm[threadIdx.x >> 5][1] = smax;
          }
        }
        for (int k = lt; k < BINS; k += T) s_bin[gi][k] = INT32_MIN;
        sync();
        if (T != 32) {
#pragma unroll
          for (int w = 0; w < T / 32; w++) {
            smin = min(smin, s_mm[w][0]);
            smax = max(smax, s_mm[w][1]);
          }
        }
        const uint64_t width = (uint64_t)((int64_t)smax - (int64_t)smin) + 1;
        for (uint32_t c = lt; c < nr; c += T) {
          const uint32_t i = live[c];
          const uint32_t k = (uint32_t)(((uint64_t)((int64_t)min(m.qlo[i], m.qhi[i]) - (int64_t)smin) * BINS) / width);
          atomicMax(&s_bin[gi][k], max(m.qlo[i], m.qhi[i]));
        }
        sync();
        if (threadIdx.x < 32 || T == 32) {  // exclusive prefix max over the 64 bins, two per lane
          const int v0 = s_bin[gi][lane], v1 = s_bin[gi][lane + 32];
          const int i0 = warp_incl_max(v0);
          const int t0 = __shfl_sync(FULL, i0, 31);
          const int i1 = max(t0, warp_incl_max(v1));
          int e0 = __shfl_up_sync(FULL, i0, 1), e1 = __shfl_up_sync(FULL, i1, 1);
          if (lane == 0) {
            e0 = INT32_MIN;
            e1 = t0;
          }
          s_bin[gi][lane] = e0;
          s_bin[gi][lane + 32] = e1;
        }
        if (lt == 0) s_cnt[gi][0] = 0;
        sync();
        uint32_t wbase_ = 0;
        for (uint32_t c0 = 0; c0 < nr; c0 += T) {
          const uint32_t c = c0 + lt;
          uint32_t i = 0;
          bool keep = false;
          if (c < nr) {
            i = live[c];
            const uint32_t k = (uint32_t)(((uint64_t)((int64_t)min(m.qlo[i], m.qhi[i]) - (int64_t)smin) * BINS) / width);
            keep = max(m.qlo[i], m.qhi[i]) > s_bin[gi][k];
          }
          const unsigned bm = __ballot_sync(FULL, keep);
          uint32_t wb;
          if (T == 32) {
            wb = wbase_;
            wbase_ += __popc(bm);
          } else {
            wb = 0;
            if (lane == 0 && bm) wb = atomicAdd(&s_cnt[gi][0], (unsigned)__popc(bm));
            wb = __shfl_sync(FULL, wb, 0);
          }
          if (keep) spare[wb + __popc(bm & lanemask_lt())] = (uint16_t)i;
        }
        if (T == 32 && lt == 0) s_cnt[gi][0] = wbase_;
        sync();
        nr = s_cnt[gi][0];
        uint16_t *t = live;
        live = spare;
        spare = t;
      }
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
        uint32_t wbase_ = n_piv + 1;  // warp classes: the running count lives in a register
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
            wb = wbase_;
            wbase_ += __popc(bm);
          } else {
            wb = 0;
            if (lane == 0 && bm) wb = atomicAdd(&s_cnt[gi][0], (unsigned)__popc(bm));
            wb = __shfl_sync(FULL, wb, 0);
          }
          if (keep) spare[wb + __popc(bm & lanemask_lt())] = (uint16_t)i;
        }
        if (T == 32 && lt == 0) s_cnt[gi][0] = wbase_;
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
      // Bitonic network (see k_merge_segments): warp w owns the compare-exchanges [w * ppw, (w + 1) * ppw) of
      // every step; steps whose pairs stay inside the warp's own keys only need a warp barrier
      const uint32_t half = P >> 1;
      const uint32_t ppw = (T == 32) ? half : max(32u, half / (uint32_t)(T / 32));
      const uint32_t wbase = (T == 32) ? 0u : (uint32_t)(threadIdx.x >> 5) * ppw;
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
    // ---- ties on (start, strand): restore ord order (runs of equal keys are short and disjoint)
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
    // sorted position -> (start, end) pair in place of the key, box index | forward << 15 in `spare`
    // (`live` still holds the box list the keys were built from; it is dead from here on)
    uint16_t *pos = spare;
    for (uint32_t i = lt; i < nr; i += T) {
      const uint32_t x = (uint32_t)(m.skey[i] & 0xffffu);
      const bool fwd = m.qlo[x] <= m.qhi[x];
      const int32_t st = fwd ? m.qlo[x] : m.qhi[x], en = fwd ? m.qhi[x] : m.qlo[x];
      m.skey[i] = ((uint64_t)(uint32_t)en << 32) | (uint32_t)st;
      pos[i] = (uint16_t)(x | (fwd ? 0x8000u : 0u));
    }
    sync();
    auto emit = [&](uint32_t w, uint32_t first_pos, int32_t st, int32_t en, bool fwd) {
      const uint32_t c = pos[first_pos] & 0x7fffu;
      store_seg_out(seg + w, fwd ? st : en, fwd ? en : st, m.tid[c], m.tlo[c], m.thi[c]);
    };
    if (reduce) {
      // survivors (first warp): with merge_strands and d >= 0 the boxes whose end exceeds every earlier end,
      // otherwise every stage-A result (a strand-aware or non-merging sweep can see all of them)
      if (threadIdx.x < 32 || T == 32) {
        const bool prune = merge_strands && md >= 0;
        int carry_pm = INT32_MIN;
        uint32_t carry_cnt = 0;
        for (uint32_t k0 = 0; k0 < nr; k0 += 32) {
          const uint32_t k = k0 + lane;
          const bool valid = k < nr;
          const int32_t en = valid ? (int32_t)(uint32_t)(m.skey[k] >> 32) : INT32_MIN;
          const int pm_in = max(carry_pm, warp_incl_max(en));
          int pm_ex = __shfl_up_sync(FULL, pm_in, 1);
          if (lane == 0) pm_ex = carry_pm;
          const bool keep = valid && (!prune || k == 0 || en > pm_ex);
          const unsigned bm = __ballot_sync(FULL, keep);
          if (keep) {
            const uint32_t c = pos[k] & 0x7fffu;
            const uint64_t o = m.ord[c];
            uint4 *dst = reinterpret_cast<uint4 *>(seg + carry_cnt + __popc(bm & lanemask_lt()));
            dst[0] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), (uint32_t)m.qlo[c], (uint32_t)m.qhi[c]);
            dst[1] = make_uint4(m.tid[c], (uint32_t)m.tlo[c], (uint32_t)m.thi[c], BOX_MERGED_A);
          }
          carry_pm = __shfl_sync(FULL, pm_in, 31);
          carry_cnt += __popc(bm);
        }
        if (lane == 0) out_cnt[bk] = carry_cnt;
      }
    } else if (merge_strands) {
      // ---- the sweep of src/main.rs:12496-12556 as scans over the sorted boxes (first warp), see k_merge_segments
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
                 (pos[os_ex] & 0x8000u) != 0);
          }
          carry_pm = __shfl_sync(FULL, pm_in, 31);
          carry_lb = __shfl_sync(FULL, lb_in, 31);
          carry_os = __shfl_sync(FULL, os_in, 31);
          carry_cnt += __popc(bm);
        }
        if (lane == 0) {
          const uint64_t fe = m.skey[carry_lb];
          emit(carry_cnt - 1, (uint32_t)carry_lb, (int32_t)(uint32_t)fe, md < 0 ? (int32_t)(uint32_t)(fe >> 32) : carry_pm,
               (pos[carry_os] & 0x8000u) != 0);
          out_cnt[bk] = carry_cnt;
        }
      }
    } else if (lt == 0) {
      // --consider-strandness: the literal sequential sweep (a strand change always starts a new row)
      uint32_t w = 0, cpos = 0;
      uint64_t se = m.skey[0];
      int32_t cs = (int32_t)(uint32_t)se, ce = (int32_t)(uint32_t)(se >> 32);
      bool cf = (pos[0] & 0x8000u) != 0;
      for (uint32_t rd = 1; rd < nr; rd++) {
        se = m.skey[rd];
        const int32_t ns = (int32_t)(uint32_t)se, ne = (int32_t)(uint32_t)(se >> 32);
        const bool nf = (pos[rd] & 0x8000u) != 0;
        if (md < 0 || cf != nf || (int64_t)ns > (int64_t)ce + md) {
          emit(w++, cpos, cs, ce, cf);
          cpos = rd; cs = ns; ce = ne; cf = nf;
        } else {
          cs = min(cs, ns);
          ce = max(ce, ne);
        }
      }
      emit(w++, cpos, cs, ce, cf);
      out_cnt[bk] = w;
    }
    sync();  // the next bucket reuses the shared arrays
  }
}

// merged rows of every bucket -> the output columns, in (row, q) order = bucket order
__global__ void k_bucket_compact(const BoxRec *__restrict__ boxes, const uint32_t *__restrict__ beg,
                                 const uint32_t *__restrict__ out_cnt, const uint32_t *__restrict__ out_off,
                                 uint64_t n_buckets, uint32_t n_seqs, OutCols o) {
  for (uint64_t b = gtid(); b < n_buckets; b += gstride()) {
    const uint32_t c = out_cnt[b];
    if (!c) continue;
    const uint32_t q = (uint32_t)(b % n_seqs);
    const BoxRec *seg = boxes + beg[b];
    const uint64_t d0 = out_off[b];
    for (uint32_t k = 0; k < c; k++) {
      const SegOut x = *reinterpret_cast<const SegOut *>(seg + k);
      o.q_id[d0 + k] = q;
      o.q_first[d0 + k] = x.q_first;
      o.q_last[d0 + k] = x.q_last;
      o.t_id[d0 + k] = x.t_id;
      o.t_first[d0 + k] = x.t_first;
      o.t_last[d0 + k] = x.t_last;
    }
  }
}
__global__ void k_bucket_row_offsets(const uint32_t *__restrict__ out_off, uint32_t n_rows, uint32_t n_seqs,
                                     uint64_t *__restrict__ row_off) {
  for (uint64_t r = gtid(); r <= n_rows; r += gstride()) row_off[r] = out_off[r * n_seqs];
}

// ---- buckets beyond SEG_MAX boxes: their boxes as BoxD records for the global two-sort path ...
__global__ void k_oversized_sizes(const uint32_t *__restrict__ list, uint32_t n_list, const uint32_t *__restrict__ beg,
                                  const uint32_t *__restrict__ cur, uint64_t *__restrict__ sizes) {
  for (uint64_t i = gtid(); i < n_list; i += gstride()) sizes[i] = cur[list[i]] - beg[list[i]];
}
__global__ void k_oversized_keep(const uint32_t *__restrict__ list, uint32_t n_list, const uint32_t *__restrict__ beg,
                                 const uint32_t *__restrict__ cur, uint32_t *__restrict__ out_cnt) {
  for (uint64_t i = gtid(); i < n_list; i += gstride()) out_cnt[list[i]] = cur[list[i]] - beg[list[i]];
}
__global__ void k_oversized_to_boxd(const uint32_t *__restrict__ list, uint32_t n_list, const uint64_t *__restrict__ offs,
                                    const BoxRec *__restrict__ boxes, const uint32_t *__restrict__ beg, uint32_t n_seqs,
                                    BoxD *__restrict__ out) {
  // one warp per bucket
  uint64_t w = gtid() >> 5;
  const uint64_t nw = gstride() >> 5;
  const unsigned lane = lane_id();
  for (; w < n_list; w += nw) {
    const uint32_t bk = list[w];
    const uint32_t row = bk / n_seqs, q = bk % n_seqs;
    const uint64_t o0 = offs[w], n = offs[w + 1] - o0;
    const BoxRec *seg = boxes + beg[bk];
    for (uint64_t i = lane; i < n; i += 32) {
      const BoxRec x = seg[i];
      out[o0 + i] = BoxD{x.q_first, x.q_last, x.t_first, x.t_last, q, x.t_id, row,
                         1u | ((x.flags & BOX_MERGED_A) ? BOXD_MERGED_A : 0u), x.ord};
    }
  }
}
// ... and the rows the global sweep produced for them back into the buckets' staging slots
// (groups of stage B are exactly the (row, q) segments)
__global__ void k_groups_to_buckets(const BoxD *__restrict__ swept, const uint32_t *__restrict__ begins,
                                    const uint32_t *__restrict__ cnt, uint64_t n_groups, uint32_t n_seqs,
                                    const uint32_t *__restrict__ beg, BoxRec *__restrict__ boxes, uint32_t *__restrict__ out_cnt) {
  for (uint64_t g = gtid(); g < n_groups; g += gstride()) {
    const uint32_t b = begins[g], c = cnt[g];
    const BoxD f = swept[b];
    const uint64_t bk = (uint64_t)f.row * n_seqs + f.q_id;
    BoxRec *seg = boxes + beg[bk];
    for (uint32_t k = 0; k < c; k++) {
      const BoxD x = swept[b + k];
      store_seg_out(seg + k, x.q_lo, x.q_hi, x.t_id, x.t_lo, x.t_hi);
    }
    out_cnt[bk] = c;
  }
}

}  // namespace impgx
