// small_bfs.cuh — the whole walk of a row in ONE launch, for calls of a few rows.
//
// `impg partition` and `impg refine` call query_transitive_bfs once per window / flank
// (src/commands/partition.rs:359-391, src/commands/refine.rs:493-531): a frontier of a few ranges, a few
// hundred hits. The batched pipeline of engine.cu spends ~65 launches and ~20 size readbacks on such a call
// (count -> scan -> read the size -> allocate -> fill, per stage); the device work itself is microseconds.
// Here ONE CTA owns a row and runs every stage of Impg::query_transitive_bfs (src/impg.rs:2316-2593) /
// query_transitive_dfs (:2057-2309) / Impg::query (:1852-1925) on capacity-bounded buffers, with the sizes kept in shared memory:
//   seed (masked or not) -> per hop: stab count, scan, fill, endpoint liftover, order by (range, visit rank),
//   append results, fold into the visited set (thread per touched sequence, the same fold_one_group as the
//   batched path), rebuild the visited set, sort + join the next frontier.
// The results leave either in reference order (raw) or — BED — as BoxRecs grouped by query sequence with the
// bucket lists k_merge_buckets / k_merge_tiny consume (their list lengths are read from device memory), and
// k_small_finish lays the rows of all CTAs out contiguously. The host reads one header + the rows: no size
// readback in between. A row that exceeds a capacity (SB_CAP ranges / hits per hop / results / visited ranges)
// or holds an invalid range sets a status; the caller then runs the batched path, which is exact for any size.
#pragma once
#include "bucket_kernels.cuh"

namespace impgx {

constexpr int SB_THREADS = 512;
constexpr uint32_t SB_CAP = 8192;           // ranges, hits of one hop, results, visited ranges of ONE row
constexpr uint32_t SB_LISTS = 3 * SB_CAP;   // fold scratch (lists / pieces) of one row
constexpr uint32_t SB_MAX_ROWS = 64;  // one CTA each: they all run at once on 148 SMs
constexpr int SB_PHASES = 10;  // stab, fill, liftover, order, results, group, fold, visited, frontier, buckets
constexpr size_t SB_SMEM = (size_t)SB_CAP * 12;  // sort keys (u64) + values (u32)

enum : uint32_t { SB_OK = 0, SB_OVERFLOW = 1, SB_INVALID = 2 };

struct SbParams {
  uint32_t n_rows, max_depth;
  int32_t min_transitive_len, min_dist, min_out;
  int query_mode;            // 1: Impg::query (closed stab, no clipping, no walk)
  int dfs;                   // 1: query_transitive_dfs (src/impg.rs:2057-2309): one popped range per round
  int bed;                   // 1: BoxRecs in (row, q) buckets for the bucket merge; 0: results in reference order
  const uint64_t *mask_off;  // masked_regions CSR over all sequences, or nullptr
  const int2 *mask_rng;
  const uint8_t *subset;     // subset filter or nullptr
};

struct __align__(8) SbOut {
  uint32_t q_id;
  int32_t q_first, q_last;
  uint32_t t_id;
  int32_t t_first, t_last;
};

// scratch of one row
struct SbRowMem {
  Frontier *fr[2];
  Window *win;
  uint32_t *cnt;  // SB_CAP + 1
  LiftTask *tasks;
  Hit *hits, *ordered, *sorted, *res;
  uint32_t *gvr;  // visit rank per hit slot (~0 = rejected), written by every CTA of the row's cluster
  uint64_t *vkey[2], *tkey;
  int32_t *vstart[2], *vend[2], *tstart, *tend;
  FoldGroup *grp;
  uint32_t *llen, *pcnt, *loff, *poff;
  int2 *lists;
  Frontier *pieces, *pc;
  DfsEntry *stk[2];  // DFS: the stack, sorted by (sequence, start); the top is the last entry
};
__host__ __device__ inline size_t sb_carve(SbRowMem *m, char *base) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char *p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  const size_t C = SB_CAP;
  SbRowMem t;
  t.fr[0] = (Frontier *)take(C * sizeof(Frontier));
  t.fr[1] = (Frontier *)take(C * sizeof(Frontier));
  t.win = (Window *)take(C * sizeof(Window));
  t.cnt = (uint32_t *)take((C + 1) * 4);
  t.tasks = (LiftTask *)take(C * sizeof(LiftTask));
  t.hits = (Hit *)take(C * sizeof(Hit));
  t.ordered = (Hit *)take(C * sizeof(Hit));
  t.sorted = (Hit *)take(C * sizeof(Hit));
  t.res = (Hit *)take(C * sizeof(Hit));
  t.gvr = (uint32_t *)take(C * 4);
  t.vkey[0] = (uint64_t *)take(C * 8);
  t.vkey[1] = (uint64_t *)take(C * 8);
  t.tkey = (uint64_t *)take(C * 8);
  t.vstart[0] = (int32_t *)take(C * 4);
  t.vstart[1] = (int32_t *)take(C * 4);
  t.vend[0] = (int32_t *)take(C * 4);
  t.vend[1] = (int32_t *)take(C * 4);
  t.tstart = (int32_t *)take(C * 4);
  t.tend = (int32_t *)take(C * 4);
  t.grp = (FoldGroup *)take(C * sizeof(FoldGroup));
  t.llen = (uint32_t *)take((C + 1) * 4);
  t.pcnt = (uint32_t *)take((C + 1) * 4);
  t.loff = (uint32_t *)take((C + 1) * 4);
  t.poff = (uint32_t *)take((C + 1) * 4);
  t.lists = (int2 *)take((size_t)SB_LISTS * sizeof(int2));
  t.pieces = (Frontier *)take((size_t)SB_LISTS * sizeof(Frontier));
  t.pc = (Frontier *)take(C * sizeof(Frontier));
  t.stk[0] = (DfsEntry *)take(C * sizeof(DfsEntry));
  t.stk[1] = (DfsEntry *)take(C * sizeof(DfsEntry));
  if (m) *m = t;
  return off;
}
inline size_t sb_row_bytes() { return sb_carve(nullptr, nullptr); }

// A row is walked by a thread-block CLUSTER: CTA 0 (the leader) runs the walk, the other CTAs join it for the
// liftover of a hop — the one stage that is bound by memory latency per thread, so more threads in flight is the
// only lever — and otherwise wait at the cluster barrier. The leader posts the request here before the barrier.
enum : uint32_t { SB_OP_EXIT = 0, SB_OP_LIFT = 1 };
struct SbCtl {
  uint32_t op, n_hits, fb, clip;
};

// memory of one call
struct SbCall {
  char *rows;            // n_rows x row_bytes
  size_t row_bytes;
  uint32_t *row_target;  // n_rows
  uint32_t *status;      // n_rows
  uint32_t *n_res;       // n_rows: results (raw) of the row
  uint32_t *n_bk;        // n_rows: buckets of the row (bed)
  SbCtl *ctl;            // n_rows: what the leader CTA of a row asks of the other CTAs of its cluster
  unsigned long long *stats;  // [0] ranges stabbed, [1] hits lifted, [2 + k] clock cycles of phase k (zeroed by the host)
  // bed: bucket b of row r is slot r * SB_CAP + b
  BoxRec *boxes;
  uint32_t *bk_beg, *bk_cur, *bk_q, *out_cnt;
  uint32_t *lists;       // (SEG_CLASSES + 2) x (n_rows * SB_CAP)
  unsigned int *cls;     // SEG_CLASSES + 2 list lengths (zeroed by the host)
  // output: hdr[0..n_rows] = row offsets, hdr[n_rows + 1] = worst status, hdr[n_rows + 2] = ranges, [n_rows + 3] = hits
  uint32_t *hdr;
  SbOut *out;            // contiguous rows of all rows (follows hdr in memory)
};

// ---- CTA-wide helpers (every thread of the CTA calls them; n is CTA-uniform)
struct SbSync {
  uint32_t warp_sum[SB_THREADS / 32];
  uint32_t total;
};

// exclusive prefix of one value per thread; *total = the sum (CTA-uniform)
__device__ __forceinline__ uint32_t sb_block_scan(uint32_t v, SbSync &sh, uint32_t *total) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, x, d);
    if (lane >= (unsigned)d) x += y;
  }
  __syncthreads();  // the previous use of sh is over
  if (lane == 31) sh.warp_sum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SB_THREADS / 32 ? sh.warp_sum[lane] : 0u;
    uint32_t s = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, s, d);
      if (lane >= (unsigned)d) s += y;
    }
    if (lane < SB_THREADS / 32) sh.warp_sum[lane] = s - w;
    if (lane == 31) sh.total = s;
  }
  __syncthreads();
  *total = sh.total;
  return sh.warp_sum[warp] + x - v;
}

// in-place exclusive scan of a[0..n) (n <= SB_CAP); returns the sum. Ends with a barrier.
__device__ __forceinline__ uint32_t sb_scan_array(uint32_t *a, uint32_t n, SbSync &sh) {
  const uint32_t per = (n + SB_THREADS - 1) / SB_THREADS;
  const uint32_t b = threadIdx.x * per, e = min(n, b + per);
  uint32_t s = 0;
  for (uint32_t i = b; i < e; i++) s += a[i];
  uint32_t total;
  uint32_t base = sb_block_scan(s, sh, &total);
  for (uint32_t i = b; i < e; i++) {
    const uint32_t v = a[i];
    a[i] = base;
    base += v;
  }
  __syncthreads();
  return total;
}

// bitonic sort of (key, val) pairs in shared memory, ascending by key; n <= SB_CAP. Ends with a barrier.
__device__ __forceinline__ void sb_sort(uint64_t *key, uint32_t *val, uint32_t n) {
  uint32_t P = 2;
  while (P < n) P <<= 1;
  for (uint32_t i = n + threadIdx.x; i < P; i += SB_THREADS) {
    key[i] = ~0ull;
    val[i] = 0;
  }
  __syncthreads();
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < P / 2; t += SB_THREADS) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t l = i | j;
        const bool up = (i & k) == 0;
        const uint64_t a = key[i], b = key[l];
        if ((a > b) == up) {
          key[i] = b;
          key[l] = a;
          const uint32_t va = val[i];
          val[i] = val[l];
          val[l] = va;
        }
      }
      __syncthreads();
    }
}

// the same network over keys that carry their payload in the low bits (half the shared-memory traffic)
__device__ __forceinline__ void sb_sort_keys(uint64_t *key, uint32_t n) {
  uint32_t P = 2;
  while (P < n) P <<= 1;
  for (uint32_t i = n + threadIdx.x; i < P; i += SB_THREADS) key[i] = ~0ull;
  __syncthreads();
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = threadIdx.x; t < P / 2; t += SB_THREADS) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t l = i | j;
        const bool up = (i & k) == 0;
        const uint64_t a = key[i], b = key[l];
        if ((a > b) == up) {
          key[i] = b;
          key[l] = a;
        }
      }
      __syncthreads();
    }
}
constexpr uint32_t SB_IDX_BITS = 13;  // SB_CAP = 2^13: an index into any per-row array
constexpr uint64_t SB_IDX_MASK = (1ull << SB_IDX_BITS) - 1;
static_assert(SB_CAP == (1u << SB_IDX_BITS), "index bits");

__device__ __forceinline__ uint32_t sb_lower_bound_u64(const uint64_t *a, uint32_t n, uint64_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (a[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ uint32_t sb_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t sb_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// release / acquire at cluster scope: what a CTA wrote to global memory before the barrier is visible to the
// other CTAs of the cluster after it
__device__ __forceinline__ void sb_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The share of one CTA in the liftover of a hop: hit slot t -> hits[t], gvr[t].
__device__ __forceinline__ void sb_lift_share(const DevIndexView &ix, const SbRowMem &m, const Frontier *F, uint32_t H,
                                              const LiftParams &lp, uint32_t row, uint32_t first, uint32_t stride) {
  for (uint32_t t = first; t < H; t += stride) {
    const uint2 tv = __ldcg(reinterpret_cast<const uint2 *>(m.tasks + t));
    const LiftTask task{tv.x, tv.y};
    const uint4 fv = __ldcg(reinterpret_cast<const uint4 *>(F + task.range));
    Frontier f;
    f.row = fv.x; f.seq = fv.y; f.start = (int32_t)fv.z; f.end = (int32_t)fv.w;
    const EndsHit eh = lift_ends_hit<false>(ix, f, task.entry, lp);
    uint4 *dst = reinterpret_cast<uint4 *>(m.hits + t);
    dst[0] = make_uint4(eh.ok ? row : INVALID_ID, eh.query_id, (uint32_t)eh.f_q, (uint32_t)eh.l_q);
    dst[1] = make_uint4(f.seq, (uint32_t)eh.f_t, (uint32_t)eh.l_t, eh.vrank);
    m.gvr[t] = eh.ok ? eh.vrank : 0xffffffffu;
  }
}
__device__ __forceinline__ Hit sb_load_hit_cg(const Hit *p) {
  const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(p)), b = __ldcg(reinterpret_cast<const uint4 *>(p) + 1);
  Hit h;
  h.row = a.x; h.q_id = a.y; h.q_first = (int32_t)a.z; h.q_last = (int32_t)a.w;
  h.t_id = b.x; h.t_first = (int32_t)b.y; h.t_last = (int32_t)b.z; h.vrank = b.w;
  return h;
}

__global__ void __launch_bounds__(SB_THREADS, 1)
    k_small_bfs(DevIndexView ix, const impgx_range *__restrict__ ranges, SbParams p, SbCall c) {
  extern __shared__ __align__(16) unsigned char sb_smem[];
  uint64_t *skey = reinterpret_cast<uint64_t *>(sb_smem);
  uint32_t *sval = reinterpret_cast<uint32_t *>(skey + SB_CAP);
  __shared__ SbSync sh;
  __shared__ uint32_t s_u[4];
  const uint32_t CL = sb_cluster_size(), crank = sb_cluster_rank();
  const uint32_t row = blockIdx.x / CL, tid = threadIdx.x;
  const unsigned lane = tid & 31u, warp = tid >> 5;
  constexpr uint32_t NW = SB_THREADS / 32, T = SB_THREADS;
  SbRowMem m;
  sb_carve(&m, c.rows + (size_t)row * c.row_bytes);
  LiftParams lp;
  lp.clip = p.query_mode ? 0 : 1;
  lp.min_output_len = -1;
  lp.use_identity = 0;
  lp.min_identity = 0.0;
  lp.subset = p.subset;
  lp.row_target = c.row_target;
  if (crank != 0) {
    // a helper CTA: liftover shares until the leader says the walk is over
    for (;;) {
      sb_cluster_sync();
      const volatile SbCtl *ctl = c.ctl + row;
      if (ctl->op == SB_OP_EXIT) return;
      sb_lift_share(ix, m, m.fr[ctl->fb], ctl->n_hits, lp, row, crank * T + tid, CL * T);
      sb_cluster_sync();
    }
  }
  const impgx_range r = ranges[row];
  // every way out of the leader passes here: the helpers are released first
  auto release_cluster = [&]() {
    if (tid == 0) c.ctl[row].op = SB_OP_EXIT;
    sb_cluster_sync();
  };
  auto leave = [&](uint32_t status) {
    if (tid == 0) {
      c.status[row] = status;
      c.n_res[row] = 0;
      c.n_bk[row] = 0;
    }
    release_cluster();
  };
  {
    // perform_query's bounds checks (k_validate); the batched path words the error
    bool valid = r.target_id < ix.n_seqs && r.start >= 0 && r.start < r.end;
    if (valid) valid = r.end <= ix.seq_len[r.target_id];
    if (!valid) {
      leave(SB_INVALID);
      return;
    }
  }
  const uint32_t tgt = r.target_id;
  const uint64_t row_key = (uint64_t)row << 32;
  uint32_t nF = 0, nV = 0, nR = 0;
  int fb = 0, vb = 0;

  // ---- seed (src/impg.rs:2337-2373 / :1864-1880)
  if (p.query_mode || !p.mask_off) {
    if (tid == 0) {
      c.row_target[row] = tgt;
      const bool out_ok = !p.query_mode || p.min_out < 0 || (r.end - r.start) >= p.min_out;
      if (out_ok) m.res[0] = Hit{row, tgt, r.start, r.end, tgt, r.start, r.end, 0u};
      long long len = (long long)r.end - (long long)r.start;
      const bool walk = p.query_mode || p.min_transitive_len <= 0 || len >= p.min_transitive_len;
      if (walk) m.fr[0][0] = Frontier{row, tgt, r.start, r.end};
      m.vkey[0][0] = row_key | tgt;
      m.vstart[0][0] = r.start;
      m.vend[0][0] = r.end;
      s_u[0] = walk ? 1u : 0u;
      s_u[1] = 1u;
      s_u[2] = out_ok ? 1u : 0u;
      s_u[3] = SB_OK;
    }
  } else {
    // visited[target] = mask[target] + range: the unmasked pieces are the self intervals and the frontier
    const uint64_t m0 = p.mask_off[tgt], m1 = p.mask_off[tgt + 1];
    if (m1 - m0 + 1 > SB_CAP) {
      leave(SB_OVERFLOW);
      return;
    }
    const uint32_t n0 = (uint32_t)(m1 - m0);
    for (uint32_t k = tid; k < n0; k += T) m.lists[k] = p.mask_rng[m0 + k];
    __syncthreads();
    if (tid == 0) {
      c.row_target[row] = tgt;
      uint32_t mm = n0, np = 0, nf = 0;
      int2 *L = m.lists;
      ranges_insert(L, mm, ix.seq_len[tgt], r.start, r.end, [&](int32_t s, int32_t e) {
        m.res[np++] = Hit{row, tgt, s, e, tgt, s, e, 0u};
        long long len = (long long)e - (long long)s;
        if (p.min_transitive_len <= 0 || (len < 0 ? -len : len) >= p.min_transitive_len)
          m.fr[0][nf++] = Frontier{row, tgt, s, e};
      });
      for (uint32_t k = 0; k < mm; k++) {
        m.vkey[0][k] = row_key | tgt;
        m.vstart[0][k] = L[k].x;
        m.vend[0][k] = L[k].y;
      }
      s_u[0] = nf;
      s_u[1] = mm;
      s_u[2] = np;
      s_u[3] = SB_OK;
    }
  }
  __syncthreads();
  nF = s_u[0];
  nV = s_u[1];
  nR = s_u[2];
  __syncthreads();

  unsigned long long st_ranges = 0, st_hits = 0;
  // cycles per phase (thread 0; IMPGX_TRACE=2 prints them)
  long long ph[SB_PHASES] = {0}, t_prev = clock64();
  auto lap = [&](int k) {
    const long long t = clock64();
    ph[k] += t - t_prev;
    t_prev = t;
  };
  uint32_t depth = 0;
  // DFS: the seeds that are walked on are the initial stack (they are in (sequence, start) order)
  uint32_t n_stack = 0, cur_depth = 0;
  int sbuf = 0;
  if (p.dfs) {
    for (uint32_t i = tid; i < nF; i += T) {
      const Frontier f = m.fr[0][i];
      m.stk[0][i] = DfsEntry{row, f.seq, f.start, f.end, 0u, 0u};
    }
    n_stack = nF;
    __syncthreads();
  }
  for (;;) {
    if (p.dfs) {
      // pop the top; an entry at max_depth is dropped (:2125-2127)
      if (n_stack == 0) break;
      const DfsEntry top = m.stk[sbuf][n_stack - 1];
      n_stack--;
      if (p.max_depth > 0 && top.depth >= p.max_depth) continue;
      if (tid == 0) m.fr[fb][0] = Frontier{row, top.id, top.start, top.end};
      nF = 1;
      cur_depth = top.depth;
    } else if (!(nF > 0 && (p.query_mode ? depth == 0 : (p.max_depth == 0 || depth < p.max_depth)))) {
      break;
    }
    const bool last = !p.dfs && (p.query_mode || (p.max_depth != 0 && depth + 1 >= p.max_depth));
    const bool closed = p.query_mode != 0;
    const Frontier *F = m.fr[fb];
    st_ranges += nF;
    if (tid == 0) s_u[0] = 0;
    __syncthreads();
    // ---- stab: windows and hit counts, one warp per range
    for (uint32_t i = warp; i < nF; i += NW) {
      const Frontier f = F[i];
      const Window wd = closed ? stab_window<true>(ix, f.seq, f.start, f.end) : stab_window<false>(ix, f.seq, f.start, f.end);
      uint32_t cn = 0;
      for (uint64_t j = wd.lb + lane; j < wd.ub; j += 32) {
        const int32_t e = ix.e_end[j];
        cn += (closed ? e >= f.start : e > f.start) ? 1u : 0u;
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) cn += __shfl_xor_sync(FULL, cn, d);
      if (lane == 0) {
        m.win[i] = wd;
        m.cnt[i] = cn;
        atomicMax(&s_u[0], cn);
      }
    }
    __syncthreads();
    const uint32_t max_per_range = s_u[0];
    lap(0);
    const uint32_t H = sb_scan_array(m.cnt, nF, sh);
    if (H > SB_CAP) {
      leave(SB_OVERFLOW);
      return;
    }
    st_hits += H;
    // ---- fill: the hits of range i at cnt[i]..., in index order (the visit rank orders them below)
    for (uint32_t i = warp; i < nF; i += NW) {
      const int32_t rs = F[i].start;
      const Window wd = m.win[i];
      uint32_t base = m.cnt[i];
      for (uint64_t j0 = wd.lb; j0 < wd.ub; j0 += 32) {
        const uint64_t j = j0 + lane;
        bool hit = false;
        if (j < wd.ub) {
          const int32_t e = ix.e_end[j];
          hit = closed ? e >= rs : e > rs;
        }
        const unsigned b = __ballot_sync(FULL, hit);
        if (hit) m.tasks[base + __popc(b & lanemask_lt())] = LiftTask{(uint32_t)j, i};
        base += __popc(b);
      }
    }
    __syncthreads();
    lap(1);
    // ---- endpoint liftover, one thread per hit, over every CTA of the cluster; order key = (range, visit rank)
    if (tid == 0) c.ctl[row] = SbCtl{SB_OP_LIFT, H, (uint32_t)fb, 0u};
    sb_cluster_sync();
    sb_lift_share(ix, m, F, H, lp, row, tid, CL * T);
    sb_cluster_sync();
    // Ranges with few hits each (the common case): the rank of a hit among the accepted hits of its range is
    // counted directly (vrank is unique within a target), O(hits of the range) shared-memory reads per hit and no
    // sorting network. A range with many hits takes the bitonic sort instead.
    const bool enumerate = max_per_range <= 512;
    uint32_t *vr = sval;                                   // visit rank per hit slot, ~0 = rejected
    uint32_t *okpre = reinterpret_cast<uint32_t *>(skey);  // accepted hits before the slot
    uint32_t okc = 0;
    for (uint32_t t = tid; t < H; t += T) {
      const uint32_t v = __ldcg(m.gvr + t);
      const bool ok = v != 0xffffffffu;
      if (enumerate) {
        vr[t] = v;
        okpre[t] = ok ? 1u : 0u;
      } else {
        // (range: 13 bits, visit rank: 32 bits, hit slot: 13 bits)
        skey[t] = ok ? (((uint64_t)m.tasks[t].range << (32 + SB_IDX_BITS)) | ((uint64_t)v << SB_IDX_BITS) | t) : ~0ull;
      }
      okc += ok ? 1u : 0u;
    }
    uint32_t n_ok;
    sb_block_scan(okc, sh, &n_ok);
    lap(2);
    if (enumerate) {
      sb_scan_array(okpre, H, sh);
      for (uint32_t t = tid; t < H; t += T) {
        const uint32_t mine = vr[t];
        if (mine == 0xffffffffu) continue;
        const uint32_t r = m.tasks[t].range;
        const uint32_t lo = m.cnt[r], hi = r + 1 < nF ? m.cnt[r + 1] : H;
        uint32_t rank = 0;
        for (uint32_t u = lo; u < hi; u++) rank += vr[u] < mine ? 1u : 0u;
        m.ordered[okpre[lo] + rank] = sb_load_hit_cg(m.hits + t);
      }
    } else {
      sb_sort_keys(skey, H);
      for (uint32_t k = tid; k < n_ok; k += T) m.ordered[k] = sb_load_hit_cg(m.hits + (uint32_t)(skey[k] & SB_IDX_MASK));
    }
    __syncthreads();
    lap(3);
    // ---- results of the hop, in reference order, filtered by min_output_length
    {
      uint32_t *pass = reinterpret_cast<uint32_t *>(skey);  // flag, then position among the passing hits
      for (uint32_t k = tid; k < n_ok; k += T) pass[k] = passes_len(m.ordered[k], p.min_out) ? 1u : 0u;
      __syncthreads();
      uint32_t last_flag = 0;
      if (n_ok) last_flag = pass[n_ok - 1];
      __syncthreads();
      const uint32_t n_pass = sb_scan_array(pass, n_ok, sh);
      if (nR + n_pass > SB_CAP) {
        leave(SB_OVERFLOW);
        return;
      }
      for (uint32_t k = tid; k < n_ok; k += T) {
        const bool is_pass = k + 1 < n_ok ? pass[k + 1] != pass[k] : last_flag != 0;
        if (is_pass) m.res[nR + pass[k]] = m.ordered[k];
      }
      nR += n_pass;
    }
    __syncthreads();
    lap(4);
    if (last) break;

    // ---- fold (src/impg.rs:2467-2560): hits grouped by query sequence, stable in reference order; hits back onto
    // the range's own sequence are not expanded (:2507)
    uint32_t incc = 0;
    for (uint32_t k = tid; k < n_ok; k += T) {
      const Hit h = m.ordered[k];
      const bool inc = h.q_id != h.t_id;
      skey[k] = inc ? (((uint64_t)h.q_id << SB_IDX_BITS) | k) : ~0ull;
      incc += inc ? 1u : 0u;
    }
    uint32_t n_inc;
    sb_block_scan(incc, sh, &n_inc);
    sb_sort_keys(skey, n_ok);
    for (uint32_t k = tid; k < n_inc; k += T) m.sorted[k] = m.ordered[(uint32_t)(skey[k] & SB_IDX_MASK)];
    __syncthreads();
    // group heads -> group starts (cnt is free again)
    uint32_t G = 0;
    for (uint32_t k0 = 0; k0 < n_inc; k0 += T) {
      const uint32_t k = k0 + tid;
      const bool head = k < n_inc && (k == 0 || m.sorted[k - 1].q_id != m.sorted[k].q_id);
      uint32_t round;
      const uint32_t pos = sb_block_scan(head ? 1u : 0u, sh, &round);
      if (head) m.cnt[G + pos] = k;
      G += round;
    }
    __syncthreads();
    const uint64_t *VK = m.vkey[vb];
    for (uint32_t g = tid; g < G; g += T) {
      FoldGroup fg;
      fg.hit_begin = m.cnt[g];
      fg.hit_end = g + 1 < G ? m.cnt[g + 1] : n_inc;
      const uint32_t q = m.sorted[fg.hit_begin].q_id;
      fg.key = row_key | q;
      fg.v_begin = sb_lower_bound_u64(VK, nV, fg.key);
      fg.v_end = sb_lower_bound_u64(VK, nV, fg.key + 1);
      fg.list_off = fg.piece_off = 0;
      m.grp[g] = fg;
      uint64_t n0 = fg.v_end - fg.v_begin;
      const uint64_t h = fg.hit_end - fg.hit_begin;
      if (n0 == 0 && p.mask_off) n0 = p.mask_off[q + 1] - p.mask_off[q];
      m.loff[g] = (uint32_t)min(n0 + h, (uint64_t)0x7fffffffu / SB_CAP);
      m.poff[g] = (uint32_t)min(n0 + 2 * h, (uint64_t)0x7fffffffu / SB_CAP);
    }
    __syncthreads();
    const uint32_t l_tot = sb_scan_array(m.loff, G, sh);
    const uint32_t p_tot = sb_scan_array(m.poff, G, sh);
    if (l_tot > SB_LISTS || p_tot > SB_LISTS) {
      leave(SB_OVERFLOW);
      return;
    }
    lap(5);
    // the lists the fold edits live in the (idle) sort buffer when they fit: shared-memory instead of L2 latency
    // on every step of the sequential insertions
    int2 *LS = l_tot <= SB_SMEM / sizeof(int2) ? reinterpret_cast<int2 *>(sb_smem) : m.lists;
    for (uint32_t g = tid; g < G; g += T) {
      FoldGroup fg = m.grp[g];
      fg.list_off = m.loff[g];
      fg.piece_off = m.poff[g];
      m.grp[g] = fg;
      uint32_t ll = 0, np = 0;
      fold_one_group(fg, m.sorted, m.vstart[vb], m.vend[vb], ix.seq_len, p.min_dist, p.min_transitive_len, LS,
                     m.pieces, p.mask_off, p.mask_rng, ll, np);
      m.llen[g] = ll;
      m.pcnt[g] = np;
    }
    __syncthreads();
    lap(6);
    // ---- new visited set: untouched old entries + the groups' lists, sorted by (sequence, start)
    uint32_t kept = 0;
    for (uint32_t i0 = 0; i0 < nV; i0 += T) {
      const uint32_t i = i0 + tid;
      bool keep = false;
      uint64_t k = 0;
      if (i < nV) {
        k = VK[i];
        uint32_t lo = 0, hi = G;
        while (lo < hi) {
          const uint32_t mid = lo + (hi - lo) / 2;
          if (m.grp[mid].key < k) lo = mid + 1;
          else hi = mid;
        }
        keep = !(lo < G && m.grp[lo].key == k);
      }
      uint32_t round;
      const uint32_t pos = sb_block_scan(keep ? 1u : 0u, sh, &round);
      if (keep) {
        m.tkey[kept + pos] = k;
        m.tstart[kept + pos] = m.vstart[vb][i];
        m.tend[kept + pos] = m.vend[vb][i];
      }
      kept += round;
    }
    __syncthreads();
    for (uint32_t g = tid; g < G; g += T) m.loff[g] = m.llen[g];
    __syncthreads();
    const uint32_t n_new = sb_scan_array(m.loff, G, sh);
    if (kept + n_new > SB_CAP) {
      leave(SB_OVERFLOW);
      return;
    }
    for (uint32_t g = tid; g < G; g += T) {
      const FoldGroup fg = m.grp[g];
      const int2 *L = LS + fg.list_off;
      const uint32_t o = kept + m.loff[g];
      for (uint32_t k = 0; k < m.llen[g]; k++) {
        m.tkey[o + k] = fg.key;
        m.tstart[o + k] = L[k].x;
        m.tend[o + k] = L[k].y;
      }
    }
    __syncthreads();
    const uint32_t nV2 = kept + n_new;
    for (uint32_t i = tid; i < nV2; i += T) {
      skey[i] = ((uint64_t)(uint32_t)m.tkey[i] << 32) | ((uint32_t)m.tstart[i] ^ 0x80000000u);
      sval[i] = i;
    }
    __syncthreads();
    sb_sort(skey, sval, nV2);
    for (uint32_t i = tid; i < nV2; i += T) {
      const uint32_t s = sval[i];
      m.vkey[vb ^ 1][i] = m.tkey[s];
      m.vstart[vb ^ 1][i] = m.tstart[s];
      m.vend[vb ^ 1][i] = m.tend[s];
    }
    __syncthreads();
    vb ^= 1;
    nV = nV2;
    lap(7);
    // ---- next frontier: the uncovered pieces sorted by (sequence, start), touching ones joined (:2566-2584)
    for (uint32_t g = tid; g < G; g += T) m.poff[g] = m.pcnt[g];
    __syncthreads();
    const uint32_t n_pieces = sb_scan_array(m.poff, G, sh);
    if (n_pieces > SB_CAP) {
      leave(SB_OVERFLOW);
      return;
    }
    for (uint32_t g = tid; g < G; g += T) {
      const Frontier *P = m.pieces + m.grp[g].piece_off;
      Frontier *O = m.pc + m.poff[g];
      for (uint32_t k = 0; k < m.pcnt[g]; k++) O[k] = P[k];
    }
    __syncthreads();
    if (p.dfs) {
      // push the pieces with depth + 1, re-sort the stack by (sequence, start), join touching entries keeping the
      // first one's depth (:2289-2304). Entries of one sequence are disjoint (each is a part that had not been
      // visited when it was pushed), so an entry joins its predecessor exactly when the two touch.
      if (n_stack + n_pieces > SB_CAP) {
        leave(SB_OVERFLOW);
        return;
      }
      DfsEntry *S = m.stk[sbuf], *D = m.stk[sbuf ^ 1];
      for (uint32_t i = tid; i < n_pieces; i += T) {
        const Frontier f = m.pc[i];
        S[n_stack + i] = DfsEntry{row, f.seq, f.start, f.end, cur_depth + 1, 0u};
      }
      __syncthreads();
      const uint32_t ms = n_stack + n_pieces;
      for (uint32_t i = tid; i < ms; i += T) {
        skey[i] = ((uint64_t)S[i].id << 32) | ((uint32_t)S[i].start ^ 0x80000000u);
        sval[i] = i;
      }
      __syncthreads();
      sb_sort(skey, sval, ms);
      uint32_t n_new = 0;
      for (uint32_t i0 = 0; i0 < ms; i0 += T) {
        const uint32_t i = i0 + tid;
        bool head = false;
        if (i < ms) {
          head = true;
          if (i > 0) {
            const DfsEntry a = S[sval[i - 1]], b = S[sval[i]];
            head = !(a.id == b.id && a.end >= b.start);
          }
        }
        uint32_t round;
        const uint32_t pos = sb_block_scan(head ? 1u : 0u, sh, &round);
        if (head) {
          DfsEntry e = S[sval[i]];
          for (uint32_t j = i + 1; j < ms; j++) {
            const DfsEntry a = S[sval[j - 1]], b = S[sval[j]];
            if (!(a.id == b.id && a.end >= b.start)) break;
            e.end = max(e.end, b.end);
          }
          D[n_new + pos] = e;
        }
        n_new += round;
      }
      __syncthreads();
      sbuf ^= 1;
      n_stack = n_new;
      lap(8);
      continue;
    }
    for (uint32_t i = tid; i < n_pieces; i += T) {
      const Frontier f = m.pc[i];
      skey[i] = ((uint64_t)f.seq << 32) | ((uint32_t)f.start ^ 0x80000000u);
      sval[i] = i;
    }
    __syncthreads();
    sb_sort(skey, sval, n_pieces);
    Frontier *PS = m.pieces;  // the group slots are consumed
    for (uint32_t i = tid; i < n_pieces; i += T) PS[i] = m.pc[sval[i]];
    __syncthreads();
    Frontier *NF = m.fr[fb ^ 1];
    uint32_t n_next = 0;
    for (uint32_t i0 = 0; i0 < n_pieces; i0 += T) {
      const uint32_t i = i0 + tid;
      bool head = false;
      if (i < n_pieces) {
        head = true;
        if (i > 0) {
          const Frontier a = PS[i - 1], b = PS[i];
          head = !(a.seq == b.seq && a.end >= b.start);
        }
      }
      uint32_t round;
      const uint32_t pos = sb_block_scan(head ? 1u : 0u, sh, &round);
      if (head) {
        Frontier f = PS[i];
        int32_t e = f.end;
        for (uint32_t j = i + 1; j < n_pieces; j++) {
          const Frontier a = PS[j - 1], b = PS[j];
          if (!(a.seq == b.seq && a.end >= b.start)) break;
          e = max(e, b.end);
        }
        f.end = e;
        NF[n_next + pos] = f;
      }
      n_next += round;
    }
    __syncthreads();
    fb ^= 1;
    nF = n_next;
    depth++;
    lap(8);
  }

  if (tid == 0) {
    atomicAdd(&c.stats[0], st_ranges);
    atomicAdd(&c.stats[1], st_hits);
  }
  if (!p.bed) {
    if (tid == 0)
      for (int k = 0; k < SB_PHASES; k++) atomicAdd(&c.stats[2 + k], (unsigned long long)ph[k]);
    if (tid == 0) {
      c.status[row] = SB_OK;
      c.n_res[row] = nR;
      c.n_bk[row] = 0;
    }
    release_cluster();
    return;
  }
  release_cluster();  // the bucket stage is the leader's alone
  // ---- BED: the results as BoxRecs grouped by query sequence (ord = position in the reference's result order),
  // one bucket per query sequence, listed by size class for the bucket merge kernels
  const uint32_t slot0 = row * SB_CAP;
  uint32_t nb = 0;
  // Grouping through a shared-memory hash table of the query sequences (no sort of the results): count per
  // sequence, sort the DISTINCT sequences (a few dozen to a few hundred), scan, scatter through per-bucket cursors.
  // The order inside a bucket is arbitrary, which the bucket merge allows (it works on the box set, ties by ord).
  constexpr uint32_t HT = 4096, HD = 2048;  // slots, distinct sequences the table path takes
  uint32_t *tabk = reinterpret_cast<uint32_t *>(sb_smem);          // q + 1, 0 = empty
  uint32_t *tabc = tabk + HT;                                      // count, then cursor
  uint64_t *dkey = reinterpret_cast<uint64_t *>(tabc + HT);        // (q << 12 | slot) of the distinct sequences
  uint32_t *rslot = reinterpret_cast<uint32_t *>(dkey + HD);       // table slot of every result (SB_CAP entries)
  static_assert((size_t)HT * 8 + (size_t)HD * 8 + (size_t)SB_CAP * 4 <= SB_SMEM, "hash bucketing layout");
  for (uint32_t i = tid; i < 2 * HT; i += T) tabk[i] = 0u;  // keys and counts
  if (tid == 0) s_u[0] = s_u[1] = 0;  // distinct sequences, table gave up
  __syncthreads();
  for (uint32_t k = tid; k < nR; k += T) {
    const uint32_t q = m.res[k].q_id;
    uint32_t h = (q * 2654435761u) >> 20;  // 12 bits
    uint32_t probes = 0;
    for (;;) {
      const uint32_t old = atomicCAS(&tabk[h], 0u, q + 1);
      if (old == 0u) {
        const uint32_t d = atomicAdd(&s_u[0], 1u);
        if (d < HD) dkey[d] = ((uint64_t)q << 12) | h;
      }
      if (old == 0u || old == q + 1) break;
      h = (h + 1) & (HT - 1);
      if (++probes >= HT) {
        s_u[1] = 1;
        break;
      }
    }
    rslot[k] = h;
    atomicAdd(&tabc[h], 1u);
  }
  __syncthreads();
  const uint32_t n_distinct = s_u[0];
  const bool table_ok = s_u[1] == 0 && n_distinct <= HD;
  __syncthreads();
  if (table_ok) {
    sb_sort_keys(dkey, n_distinct);
    nb = n_distinct;
    // bucket b = the b-th distinct sequence in ascending order; its boxes start at the scan of the counts
    uint32_t *bcount = c.bk_cur + slot0;  // scratch until the ends are written below
    for (uint32_t b = tid; b < nb; b += T) bcount[b] = tabc[(uint32_t)(dkey[b] & (HT - 1))];
    __syncthreads();
    sb_scan_array(bcount, nb, sh);
    for (uint32_t b = tid; b < nb; b += T) {
      const uint32_t hs = (uint32_t)(dkey[b] & (HT - 1));
      c.bk_beg[slot0 + b] = slot0 + bcount[b];
      c.bk_q[slot0 + b] = (uint32_t)(dkey[b] >> 12);
      tabc[hs] = slot0 + bcount[b];  // cursor
    }
    __syncthreads();
    for (uint32_t k = tid; k < nR; k += T) {
      const Hit h = m.res[k];
      const uint32_t pos = atomicAdd(&tabc[rslot[k]], 1u);
      uint4 *dst = reinterpret_cast<uint4 *>(c.boxes + pos);
      dst[0] = make_uint4(k, 0u, (uint32_t)h.q_first, (uint32_t)h.q_last);
      dst[1] = make_uint4(h.t_id, (uint32_t)h.t_first, (uint32_t)h.t_last, 0u);
    }
  } else {
    // thousands of distinct query sequences in one row: sort the results by (sequence, ordinal)
    for (uint32_t k = tid; k < nR; k += T) skey[k] = ((uint64_t)m.res[k].q_id << SB_IDX_BITS) | k;
    __syncthreads();
    sb_sort_keys(skey, nR);
    for (uint32_t j = tid; j < nR; j += T) {
      const uint32_t k = (uint32_t)(skey[j] & SB_IDX_MASK);
      const Hit h = m.res[k];
      uint4 *dst = reinterpret_cast<uint4 *>(c.boxes + slot0 + j);
      dst[0] = make_uint4(k, 0u, (uint32_t)h.q_first, (uint32_t)h.q_last);
      dst[1] = make_uint4(h.t_id, (uint32_t)h.t_first, (uint32_t)h.t_last, 0u);
    }
    for (uint32_t j0 = 0; j0 < nR; j0 += T) {
      const uint32_t j = j0 + tid;
      const bool head = j < nR && (j == 0 || (skey[j - 1] >> SB_IDX_BITS) != (skey[j] >> SB_IDX_BITS));
      uint32_t round;
      const uint32_t pos = sb_block_scan(head ? 1u : 0u, sh, &round);
      if (head) {
        c.bk_beg[slot0 + nb + pos] = slot0 + j;
        c.bk_q[slot0 + nb + pos] = (uint32_t)(skey[j] >> SB_IDX_BITS);
      }
      nb += round;
    }
  }
  __syncthreads();
  bool over = false;
  for (uint32_t b = tid; b < nb; b += T) {
    const uint32_t beg = c.bk_beg[slot0 + b];
    const uint32_t end = b + 1 < nb ? c.bk_beg[slot0 + b + 1] : slot0 + nR;
    c.bk_cur[slot0 + b] = end;
    c.out_cnt[slot0 + b] = 0;
    const uint32_t n = end - beg;
    int cl = 0;
    if (n <= (uint32_t)TINY_MAX) cl = TINY_CLASS;
    else
      while (cl < SEG_CLASSES && n > (uint32_t)seg_cap(cl)) cl++;
    if (cl == SEG_CLASSES) {
      over = true;  // a bucket beyond SEG_MAX boxes: the batched path has the global merge for it
      continue;
    }
    const unsigned int k = atomicAdd(&c.cls[cl], 1u);
    c.lists[(size_t)cl * ((size_t)p.n_rows * SB_CAP) + k] = slot0 + b;
  }
  const int any_over = __syncthreads_or(over ? 1 : 0);
  lap(9);
  if (tid == 0) {
    for (int k = 0; k < SB_PHASES; k++) atomicAdd(&c.stats[2 + k], (unsigned long long)ph[k]);
    c.status[row] = any_over ? SB_OVERFLOW : SB_OK;
    c.n_res[row] = nR;
    c.n_bk[row] = nb;
  }
}

// Rows of every CTA laid out contiguously behind the header (one CTA; the rows of a call are few).
__global__ void __launch_bounds__(SB_THREADS) k_small_finish(SbParams p, SbCall c) {
  __shared__ SbSync sh;
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t T = SB_THREADS;
  uint32_t running = 0, worst = SB_OK;
  for (uint32_t row = 0; row < p.n_rows; row++) {
    if (tid == 0) c.hdr[row] = running;
    worst = max(worst, c.status[row]);
    if (c.status[row] != SB_OK) continue;
    if (!p.bed) {
      SbRowMem m;
      sb_carve(&m, c.rows + (size_t)row * c.row_bytes);
      const uint32_t n = c.n_res[row];
      for (uint32_t k = tid; k < n; k += T) {
        const Hit h = m.res[k];
        c.out[running + k] = SbOut{h.q_id, h.q_first, h.q_last, h.t_id, h.t_first, h.t_last};
      }
      running += n;
    } else {
      const uint32_t slot0 = row * SB_CAP, nb = c.n_bk[row];
      for (uint32_t b = tid; b < nb; b += T) c.bk_cur[slot0 + b] = c.out_cnt[slot0 + b];  // bk_cur is consumed: scan buffer
      __syncthreads();
      const uint32_t total = sb_scan_array(c.bk_cur + slot0, nb, sh);
      for (uint32_t b = tid; b < nb; b += T) {
        const uint32_t cn = c.out_cnt[slot0 + b], q = c.bk_q[slot0 + b];
        const BoxRec *seg = c.boxes + c.bk_beg[slot0 + b];
        SbOut *o = c.out + running + c.bk_cur[slot0 + b];
        for (uint32_t k = 0; k < cn; k++) {
          const SegOut x = *reinterpret_cast<const SegOut *>(seg + k);
          o[k] = SbOut{q, x.q_first, x.q_last, x.t_id, x.t_first, x.t_last};
        }
      }
      running += total;
      __syncthreads();
    }
  }
  if (tid == 0) {
    c.hdr[p.n_rows] = running;
    c.hdr[p.n_rows + 1] = worst;
    c.hdr[p.n_rows + 2] = (uint32_t)min(c.stats[0], 0xffffffffull);
    c.hdr[p.n_rows + 3] = (uint32_t)min(c.stats[1], 0xffffffffull);
  }
}

}  // namespace impgx
