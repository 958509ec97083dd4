// kernels.cuh — hand-written sm_100a kernels of the projection path.
//
//   k_build_blocks   K0  raw CIGAR runs -> padded 32-run blocks + position checkpoints
//   k_stab_count     K1a warp-per-range interval stabbing, count pass
//   k_stab_fill      K1b same scan, emits (entry, range) lift tasks via ballot/prefix
//   k_liftover       K2  warp-per-hit CIGAR walk (thread-per-run, checkpointed start)
//
// All arithmetic is integer (i32 coordinates, u32 packed runs); the work is
// HBM-bound gather/scan, tensor cores are deliberately unused.
#pragma once
#include "common.cuh"

namespace impgx {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming (read-once) loads: bypass L1 allocation so scans do not evict the
// entry/checkpoint lines other warps are about to reuse.
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t *p) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Loads of the liftover's software pipeline: volatile, so they are issued where they are written.
__device__ __forceinline__ uint2 ld_pipe_u2(const void *p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_pipe_u4(const void *p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Warp-cooperative partition point: pred is monotone (true...true false...false)
// over [lo, hi); returns the first index where it is false. 32-ary search: each
// round every lane probes one position, a ballot narrows the range 32-fold.
template <class Pred>
__device__ __forceinline__ uint64_t warp_partition_point(uint64_t lo, uint64_t hi, Pred pred) {
  const unsigned lane = lane_id();
  while (hi - lo > 32) {
    uint64_t step = (hi - lo + 31) / 32;  // probe positions lo + (lane+1)*step - 1
    uint64_t pos = lo + (uint64_t)(lane + 1) * step - 1;
    bool in = pos < hi;
    bool t = in ? pred(pos) : false;
    unsigned b = __ballot_sync(FULL, t);
    unsigned k = __popc(b);  // lanes 0..k-1 true (monotone)
    uint64_t nlo = lo + (uint64_t)k * step;           // everything before probe k is true
    uint64_t nhi = lo + (uint64_t)(k + 1) * step - 1; // probe k is false (or out of range)
    if (nhi > hi) nhi = hi;
    lo = nlo;
    hi = nhi;
    if (lo >= hi) return hi;
  }
  uint64_t pos = lo + lane;
  bool t = pos < hi ? pred(pos) : false;
  unsigned b = __ballot_sync(FULL, t);
  return lo + __popc(b);
}

__device__ __forceinline__ int warp_incl_scan(int v) {
  const unsigned lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(FULL, v, d);
    if (lane >= (unsigned)d) v += o;
  }
  return v;
}

// ------------------------------------------------------------------ K0
// One warp per alignment: write its region of the stream — checkpoints first,
// then the raw runs as zero-padded 8-run (32-byte) blocks. A warp iteration
// covers 32 runs = 4 blocks; lanes 0, 8, 16, 24 own the checkpoints.
__global__ void __launch_bounds__(256) k_build_blocks(const uint32_t *__restrict__ raw,
                                                      const uint64_t *__restrict__ run_off,
                                                      const uint32_t *__restrict__ aln_off, uint64_t n_aln,
                                                      uint32_t *__restrict__ stream, int *__restrict__ bad_op) {
  const unsigned lane = lane_id();
  uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_aln; w += nw) {
    const uint64_t off = run_off[w], n = run_off[w + 1] - off;
    const uint32_t nblk = aln_nblk((uint32_t)n);
    Checkpoint *ck = const_cast<Checkpoint *>(aln_ck(stream, aln_off[w]));
    uint32_t *runs = const_cast<uint32_t *>(aln_runs(stream, aln_off[w], nblk));
    const uint64_t padded = (uint64_t)nblk * RUNS_PER_BLOCK;
    uint32_t t_acc = 0, q_acc = 0;
    for (uint64_t i0 = 0; i0 < padded; i0 += 32) {
      const uint64_t i = i0 + lane;
      const uint32_t v = i < n ? raw[off + i] : 0u;
      if (i < padded) runs[i] = v;
      const uint32_t op = v >> 29, len = v & 0x1fffffffu;
      if (op > IMPGX_OP_M && bad_op) atomicExch(bad_op, 1);  // not one of = X I D M: the build is rejected
      const int td = op == IMPGX_OP_I ? 0 : (int)len;
      const int qd = op == IMPGX_OP_D ? 0 : (int)len;
      const int ts = warp_incl_scan(td), qs = warp_incl_scan(qd);
      if ((lane % RUNS_PER_BLOCK) == 0 && i < padded)
        ck[i / RUNS_PER_BLOCK] = Checkpoint{t_acc + (uint32_t)(ts - td), q_acc + (uint32_t)(qs - qd)};
      t_acc += (uint32_t)__shfl_sync(FULL, ts, 31);
      q_acc += (uint32_t)__shfl_sync(FULL, qs, 31);
    }
    if (lane == 0) {
      ck[nblk] = Checkpoint{t_acc, q_acc};
      for (uint32_t k = nblk + 1; k < aln_ck_sectors(nblk) * 4; k++) ck[k] = Checkpoint{t_acc, q_acc};  // padding
    }
  }
}

// ------------------------------------------------------------------ K1
// Candidate window of one frontier range on its target's sorted entry columns.
// closed = true  reproduces coitrees' closed-interval visit used by Impg::query
//                (src/impg.rs:1897; iv.first <= end && start <= iv.last);
// closed = false is the visit followed by the BFS clip test that drops empty
//                overlaps (src/impg.rs:2398-2403), i.e. a half-open test.
struct Window {
  uint64_t lb, ub;
};

template <bool CLOSED>
__device__ __forceinline__ Window stab_window(const DevIndexView &ix, uint32_t seq, int32_t rs, int32_t re) {
  uint64_t lo = ix.tgt_off[seq], hi = ix.tgt_off[seq + 1];
  const int32_t *st = ix.e_start, *pm = ix.e_pmax;
  uint64_t ub = warp_partition_point(lo, hi, [&](uint64_t i) { return CLOSED ? st[i] <= re : st[i] < re; });
  uint64_t lb = warp_partition_point(lo, ub, [&](uint64_t i) { return CLOSED ? pm[i] < rs : pm[i] <= rs; });
  return Window{lb, ub};
}

// ---- TMA staging of interval tiles -------------------------------------------
// A warp stages the candidate window of the e_end column into its slice of
// shared memory with a 1-D bulk async copy (cp.async.bulk -> UBLKCP in SASS):
// one elected lane arms the warp's mbarrier with the byte count and issues the
// copy; every lane then waits on the barrier phase and scans the tile from
// shared memory. Tiles start on 16-byte boundaries of the column (the column
// is allocated with 16 bytes of slack so the last tile may round up).
constexpr int STAB_TILE = 1024;  // int32 entries per warp tile (4 KB)
constexpr int STAB_WARPS = 8;    // warps per CTA

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_tile(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}

// Scans the window [lb, ub) of `col` tile by tile through shared memory and
// calls f(i, value) for every index (all lanes of the warp participate; the
// callback may use warp collectives: `i` may be >= ub for padding lanes, which
// must be ignored via the `live` flag).
template <class F>
__device__ __forceinline__ void stab_scan_tiles(const int32_t *__restrict__ col, uint64_t lb, uint64_t ub, int32_t *tile,
                                                uint64_t *bar, uint32_t &phase, F f) {
  const unsigned lane = lane_id();
  for (uint64_t t0 = lb & ~(uint64_t)3; t0 < ub; t0 += STAB_TILE) {
    const uint64_t rem = ub - t0;
    const uint32_t n = (uint32_t)(rem < (uint64_t)STAB_TILE ? ((rem + 3) & ~(uint64_t)3) : (uint64_t)STAB_TILE);
    __syncwarp();  // every lane is done reading the previous tile
    if (lane == 0) bulk_load_tile(tile, col + t0, n * 4u, bar);
    mbar_wait(bar, phase);
    phase ^= 1u;
    for (uint32_t k0 = 0; k0 < n; k0 += 32) {
      const uint32_t k = k0 + lane;
      const uint64_t i = t0 + k;
      const bool live = k < n && i >= lb && i < ub;
      f(i, live ? tile[k] : 0, live);
    }
  }
}

// BUCKET: the hop feeds the direct BED path — every hit also counts into its (row, query sequence)
// bucket (dense table rows x n_seqs), which sizes the buckets the liftover epilogue writes into.
template <bool CLOSED, bool BUCKET>
__global__ void __launch_bounds__(32 * STAB_WARPS) k_stab_count(DevIndexView ix, const Frontier *__restrict__ fr, uint64_t n,
                                                                Window *__restrict__ win, uint32_t *__restrict__ counts,
                                                                uint32_t *__restrict__ bucket_cnt) {
  __shared__ __align__(16) int32_t tiles[STAB_WARPS][STAB_TILE];
  __shared__ uint64_t bars[STAB_WARPS];
  const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
  if (lane == 0) mbar_init(&bars[wib], 1);
  __syncwarp();
  uint32_t phase = 0;
  uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    Frontier f = fr[w];
    Window wd = stab_window<CLOSED>(ix, f.seq, f.start, f.end);
    uint32_t c = 0;
    uint32_t *brow = BUCKET ? bucket_cnt + (uint64_t)f.row * ix.n_seqs : nullptr;
    stab_scan_tiles(ix.e_end, wd.lb, wd.ub, tiles[wib], &bars[wib], phase, [&](uint64_t i, int32_t e, bool live) {
      const bool hit = live && (CLOSED ? e >= f.start : e > f.start);
      c += hit ? 1u : 0u;
      if (BUCKET && hit) atomicAdd(brow + ld_stream_u32(ix.e_qid + i), 1u);
    });
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(FULL, c, d);
    if (lane == 0) {
      win[w] = wd;
      counts[w] = c;
    }
  }
}

template <bool CLOSED>
__global__ void __launch_bounds__(32 * STAB_WARPS) k_stab_fill(DevIndexView ix, const Frontier *__restrict__ fr, uint64_t n,
                                                               const Window *__restrict__ win,
                                                               const uint64_t *__restrict__ offsets,
                                                               LiftTask *__restrict__ tasks) {
  __shared__ __align__(16) int32_t tiles[STAB_WARPS][STAB_TILE];
  __shared__ uint64_t bars[STAB_WARPS];
  const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
  if (lane == 0) mbar_init(&bars[wib], 1);
  __syncwarp();
  uint32_t phase = 0;
  uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    const int32_t rs = fr[w].start;
    Window wd = win[w];
    uint64_t base = offsets[w];
    stab_scan_tiles(ix.e_end, wd.lb, wd.ub, tiles[wib], &bars[wib], phase, [&](uint64_t i, int32_t e, bool live) {
      const bool hit = live && (CLOSED ? e >= rs : e > rs);
      const unsigned b = __ballot_sync(FULL, hit);
      if (hit) tasks[base + __popc(b & lanemask_lt())] = LiftTask{(uint32_t)i, (uint32_t)w};
      base += __popc(b);
    });
  }
}

// ------------------------------------------------------------------ K2
// Result of one liftover, uniform across the warp.
struct LiftOut {
  bool ok;
  int32_t q_start, q_end, t_start, t_end;
  // clipped CIGAR slice in walk order (for store_cigar / identity)
  uint32_t first_idx, last_idx;  // walk-order op indices [first, last)
  int32_t first_off, last_rem;   // first_op_offset (>0 trims), last_op_remaining (<0 trims)
  // gap-compressed identity terms over the clipped slice (src/impg.rs:2952-2973)
  int32_t matches, mismatches, n_ins, n_del;
  uint32_t runs_read;            // runs loaded (for the roofline accounting)
};

// project_target_range_through_alignment (src/impg.rs:2760-2898) for entry `rec`
// and request [rs, re), by one warp. The walk starts at the last 32-run block
// boundary whose target position is < rs (checkpointed), so the cost is
// O(runs overlapping the request) instead of O(runs from the alignment start).
//
// Entry orientation (src/impg.rs:144-156,547-550): a REVERSED entry swaps
// I<->D, and on the '-' strand also walks the stored runs backwards.
__device__ __forceinline__ LiftOut lift_one(const EntryRec &rec, const uint32_t *__restrict__ stream, int32_t rs,
                                            int32_t re) {
  const unsigned lane = lane_id();
  const uint32_t n = rec.nruns_flags >> 2;
  const bool rev_strand = rec.nruns_flags & FLAG_STRAND;
  const bool swap_id = rec.nruns_flags & FLAG_REVERSED;
  const bool backward = swap_id && rev_strand;
  const int32_t dir = rev_strand ? -1 : 1;
  const uint32_t nblk = aln_nblk(n);
  const Checkpoint *ck = aln_ck(stream, rec.aln_off);
  const uint32_t *blk = aln_runs(stream, rec.aln_off, nblk);
  const int32_t last_target_pos = min(rec.t_end, re);

  LiftOut o;
  o.ok = false;
  o.q_start = o.q_end = o.t_start = o.t_end = -1;
  o.first_idx = o.last_idx = 0;
  o.first_off = o.last_rem = 0;
  o.matches = o.mismatches = o.n_ins = o.n_del = 0;
  o.runs_read = 0;
  if (nblk == 0) return o;

  // totals (walk-space): target axis of the walk is the query axis of the
  // stored alignment when roles are swapped
  const Checkpoint tot = ck[nblk];
  const int64_t w_tot = swap_id ? tot.q_off : tot.t_off;
  const int64_t wq_tot = swap_id ? tot.t_off : tot.q_off;
  const int64_t rel = (int64_t)rs - (int64_t)rec.t_start;  // request start relative to walk start

  // ---- locate the starting checkpoint (8-run granularity)
  uint32_t pb;  // physical block where the walk starts
  if (!backward) {
    // largest b in [0,nblk) with P(b) < rel, else 0
    uint64_t c = warp_partition_point(0, nblk, [&](uint64_t b) {
      Checkpoint k = ck[b];
      return (int64_t)(swap_id ? k.q_off : k.t_off) < rel;
    });
    pb = c > 0 ? (uint32_t)c - 1 : 0;
  } else {
    // walk block j starts at W - P(nblk-j); want largest j with that < rel,
    // i.e. the first pb with P(pb+1) > W - rel, else nblk-1
    const int64_t x = w_tot - rel;
    uint64_t c = warp_partition_point(0, nblk, [&](uint64_t b) {
      Checkpoint k = ck[b + 1];
      return (int64_t)(swap_id ? k.q_off : k.t_off) <= x;
    });
    pb = c < nblk ? (uint32_t)c : nblk - 1;
  }

  // ---- walk state at the start of that block
  int64_t tcons, qcons;  // walk-space target / query bases consumed so far
  {
    Checkpoint k = backward ? ck[pb + 1] : ck[pb];
    int64_t pt = swap_id ? k.q_off : k.t_off, pq = swap_id ? k.t_off : k.q_off;
    tcons = backward ? w_tot - pt : pt;
    qcons = backward ? wq_tot - pq : pq;
  }
  int32_t target_pos = (int32_t)((int64_t)rec.t_start + tcons);
  int32_t query_pos = rev_strand ? (int32_t)((int64_t)rec.q_end - qcons) : (int32_t)((int64_t)rec.q_start + qcons);

  // the warp scans 32 runs per iteration: forward from run `lo`, or backward
  // below run `hi` (exclusive)
  uint32_t lo = pb * RUNS_PER_BLOCK;
  uint32_t hi = min(n, (pb + 1) * RUNS_PER_BLOCK);
  bool found = false;
  int32_t last_rem = 0, first_off = 0;
  for (;;) {
    // lanes -> ops in walk order
    const uint32_t cnt = backward ? min(32u, hi) : min(32u, n - lo);
    const bool valid = lane < cnt;
    const uint32_t pi = backward ? (hi - 1 - lane) : (lo + lane);  // physical run index
    uint32_t v = valid ? blk[pi] : 0u;
    o.runs_read += cnt;
    uint32_t op = v >> 29;
    const int32_t len = (int32_t)(v & 0x1fffffffu);
    if (swap_id) op = op == IMPGX_OP_I ? IMPGX_OP_D : (op == IMPGX_OP_D ? IMPGX_OP_I : op);
    const int32_t td = (op == IMPGX_OP_I) ? 0 : len;
    const int32_t qd = (op == IMPGX_OP_D) ? 0 : len;  // |query_delta|
    const int32_t ts_incl = warp_incl_scan(td), qs_incl = warp_incl_scan(qd);
    const int32_t tp = target_pos + (ts_incl - td);
    const int32_t qp = query_pos + (qs_incl - qd) * dir;
    // loop break: first op whose starting target position is past the end
    const unsigned brk = __ballot_sync(FULL, valid && tp > last_target_pos);
    const unsigned live = brk ? ((1u << (__ffs(brk) - 1)) - 1u) : FULL;  // lanes before the break
    const bool act = valid && ((live >> lane) & 1u);

    bool ov = false;
    int32_t pqs = 0, pts = 0, pqe = 0, pte = 0, foff = 0, lrem = 0;
    bool sets_rem = false;
    if (act) {
      if (td == 0) {  // insertion arm (also any zero-length op)
        ov = tp >= rs;
        pqs = qp; pts = tp;
        pqe = qp + qd * dir; pte = tp;
      } else if (qd == 0) {  // deletion arm
        int32_t os = max(tp, rs), oe = min(tp + td, last_target_pos);
        ov = os < oe;
        pqs = qp; pts = os; foff = os - tp;
        pqe = qp; pte = oe; lrem = oe - (tp + td);
        sets_rem = true;
      } else {  // match / mismatch arm
        int32_t os = max(tp, rs), oe = min(tp + td, re);
        ov = os < oe;
        pqs = qp + (os - tp) * dir; pts = os; foff = os - tp;
        pqe = pqs + (oe - os) * dir; pte = oe; lrem = oe - (tp + td);
        sets_rem = true;
      }
    }
    const unsigned ovm = __ballot_sync(FULL, ov);
    if (ovm) {
      // walk-order index of lane 0 in this iteration
      const uint32_t widx0 = backward ? (n - hi) : lo;
      if (!found) {
        const int fl = __ffs(ovm) - 1;
        o.q_start = __shfl_sync(FULL, pqs, fl);
        o.t_start = __shfl_sync(FULL, pts, fl);
        first_off = __shfl_sync(FULL, foff, fl);
        o.first_idx = widx0 + fl;
        found = true;
      }
      const int ll = 31 - __clz(ovm);
      o.q_end = __shfl_sync(FULL, pqe, ll);
      o.t_end = __shfl_sync(FULL, pte, ll);
      o.last_idx = widx0 + ll + 1;
      const unsigned rm = __ballot_sync(FULL, ov && sets_rem);
      if (rm) last_rem = __shfl_sync(FULL, lrem, 31 - __clz(rm));
      // identity terms over overlapping ops (clipped lengths for =/X/M)
      int32_t m_ = 0, x_ = 0, i_ = 0, d_ = 0;
      if (ov) {
        int32_t ol = (td == 0 || qd == 0) ? 0 : (pte - pts);
        if (op == IMPGX_OP_EQ || op == IMPGX_OP_M) m_ = (td == 0) ? 0 : ol;
        else if (op == IMPGX_OP_X) x_ = (td == 0) ? 0 : ol;
        else if (op == IMPGX_OP_I) i_ = 1;
        else if (op == IMPGX_OP_D) d_ = 1;
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        m_ += __shfl_xor_sync(FULL, m_, d);
        x_ += __shfl_xor_sync(FULL, x_, d);
        i_ += __shfl_xor_sync(FULL, i_, d);
        d_ += __shfl_xor_sync(FULL, d_, d);
      }
      o.matches += m_; o.mismatches += x_; o.n_ins += i_; o.n_del += d_;
    }
    if (brk) break;
    target_pos += __shfl_sync(FULL, ts_incl, 31);
    query_pos += __shfl_sync(FULL, qs_incl, 31) * dir;
    if (backward) {
      if (hi <= 32) break;
      hi -= 32;
    } else {
      lo += 32;
      if (lo >= n) break;
    }
  }
  o.first_off = first_off;
  o.last_rem = last_rem;
  o.ok = found && o.q_start != o.q_end && o.t_start != o.t_end;
  return o;
}

// Parameters of a liftover launch.
struct LiftParams {
  int clip;                 // 1: BFS/DFS clip request to the entry interval (src/impg.rs:2398-2403)
  int32_t min_output_len;   // < 0 none (applied by the caller in the reference; kept out of `ok`)
  int use_identity;         // identity filter on the clipped slice (src/impg.rs:1283-1287)
  double min_identity;
  const uint8_t *subset;    // per seq id keep mask or nullptr (src/impg.rs:2430-2438)
  const uint32_t *row_target;  // per batch row: the row's original target id (subset exemption)
};

// Per-hit CIGAR slice descriptor, written only when the caller wants CIGARs.
struct __align__(16) CigarSlice {
  uint32_t first_idx, n_ops;
  int32_t first_off, last_rem;
};

__global__ void __launch_bounds__(256) k_liftover(DevIndexView ix, const Frontier *__restrict__ fr,
                                                  const LiftTask *__restrict__ tasks, uint64_t n_tasks,
                                                  LiftParams lp, Hit *__restrict__ hits,
                                                  CigarSlice *__restrict__ slices,
                                                  unsigned long long *__restrict__ counters) {
  const unsigned lane = lane_id();
  uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long runs_acc = 0, ok_acc = 0, rov_acc = 0;
  for (; w < n_tasks; w += nw) {
    const LiftTask t = tasks[w];
    const Frontier f = fr[t.range];
    const EntryRec rec = ix.e_rec[t.entry];
    int32_t rs = f.start, re = f.end;
    if (lp.clip) {
      rs = max(rs, rec.t_start);
      re = min(re, rec.t_end);
    }
    LiftOut o = lift_one(rec, ix.stream, rs, re);
    bool ok = o.ok;
    if (ok && lp.use_identity) {
      int32_t total = o.matches + o.mismatches + o.n_ins + o.n_del;
      double idy = total == 0 ? 0.0 : (double)o.matches / (double)total;
      ok = !(idy < lp.min_identity);
    }
    if (ok && lp.subset) {
      ok = rec.query_id == lp.row_target[f.row] || lp.subset[rec.query_id] != 0;
    }
    runs_acc += o.runs_read;
    rov_acc += o.last_idx > o.first_idx ? o.last_idx - o.first_idx : 1;
    ok_acc += ok ? 1 : 0;
    if (lane == 0) {
      Hit h;
      h.row = ok ? f.row : INVALID_ID;
      h.q_id = rec.query_id;
      h.q_first = o.q_start;
      h.q_last = o.q_end;
      h.t_id = f.seq;
      h.t_first = o.t_start;
      h.t_last = o.t_end;
      h.vrank = rec.vrank;
      hits[w] = h;
      if (slices) slices[w] = CigarSlice{o.first_idx, ok ? o.last_idx - o.first_idx : 0u, o.first_off, o.last_rem};
    }
  }
  if (lane == 0 && counters) {
    atomicAdd(&counters[0], runs_acc);
    atomicAdd(&counters[1], ok_acc);
  }
}

// ------------------------------------------------------------------ K2e
// Endpoint liftover, ONE THREAD PER HIT. When neither the clipped CIGAR nor the
// identity is wanted, the result of project_target_range_through_alignment
// depends only on the FIRST and the LAST overlapping op (src/impg.rs:2806-2868:
// projected_*_start are written once, projected_*_end by every overlapping
// op), so only the run blocks at the two ends of the request are read; their
// start positions come from the checkpoints (one per 8 runs). A thread reads a
// 32-byte block as two 128-bit loads and evaluates the runs branch-free; there are no
// shuffles and no intra-warp dependencies, so the kernel is bound by the
// gather traffic (entry 32 B + checkpoint probes + 2 blocks + hit 32 B).
// (v1 = warp per hit, 1233 warp-instructions per hit; v2 = 8 lanes per hit,
// 391; both were instruction-issue bound — see profiles/.)

// first index b in [lo, hi) with !(P(b) < x) (strict) / !(P(b) <= x); P is
// nondecreasing. Galloping search from `guess`: alignments have a near-uniform
// run density, so an interpolated guess is within a few checkpoints of the
// answer and the probes stay inside one or two 32-byte sectors (a plain binary
// search touched ~7 different sectors per lookup and dominated the DRAM
// traffic of the kernel).
// The 32-byte checkpoint sector (4 checkpoints) around a clamped guess; loading it is split from the search so
// that the sectors of both lookups of a hit can be in flight together.
struct CkSector {
  uint4 a, b;
};
__device__ __forceinline__ uint32_t ck_clamp(uint32_t guess, uint32_t lo, uint32_t hi) {
  return guess < lo ? lo : (guess >= hi ? hi - 1 : guess);
}
__device__ __forceinline__ CkSector ck_load_sector(const Checkpoint *__restrict__ ck, uint32_t g) {
  const uint4 *sp = reinterpret_cast<const uint4 *>(ck + (g & ~3u));
  CkSector s;
  s.a = ld_pipe_u4(sp);
  s.b = ld_pipe_u4(sp + 1);
  return s;
}
// `g` = ck_clamp(guess, lo, hi) with lo < hi, `sec` = ck_load_sector(ck, g)
__device__ __forceinline__ uint32_t ck_partition_s(const Checkpoint *__restrict__ ck, bool swap_id, uint32_t lo,
                                                   uint32_t hi, int64_t x, bool inclusive, uint32_t g, const CkSector &sec) {
  auto pred = [&](uint32_t i) {
    const Checkpoint v = ck[i];
    const int64_t p = swap_id ? v.q_off : v.t_off;
    return inclusive ? p <= x : p < x;
  };
  {
    // fast path: the whole 32-byte sector of the guess (4 checkpoints; the array is padded to whole
    // sectors with the totals, which keeps it nondecreasing) in one go. If the partition point lies
    // inside it — the common case — no further probe is needed.
    const uint32_t s4 = g & ~3u;
    const uint4 a = sec.a, b = sec.b;
    const int64_t p0 = swap_id ? a.y : a.x, p1 = swap_id ? a.w : a.z, p2 = swap_id ? b.y : b.x, p3 = swap_id ? b.w : b.z;
    const bool t0 = inclusive ? p0 <= x : p0 < x, t1 = inclusive ? p1 <= x : p1 < x;
    const bool t2 = inclusive ? p2 <= x : p2 < x, t3 = inclusive ? p3 <= x : p3 < x;
    if ((t0 || s4 <= lo) && (!t3 || s4 + 4 >= hi)) {
      uint32_t r = s4 + (uint32_t)t0 + (uint32_t)t1 + (uint32_t)t2 + (uint32_t)t3;  // true entries are a prefix
      r = r < lo ? lo : r;
      return r > hi ? hi : r;
    }
    if (!t0) {  // partition point at or below the sector start
      hi = s4;
      g = s4 - 1;
    } else {    // beyond the sector
      lo = s4 + 4;
      g = lo;
    }
    if (lo >= hi) return lo;
  }
  if (pred(g)) {  // answer in (g, hi]
    lo = g + 1;
    uint32_t step = 1;
    while (lo < hi) {
      const uint32_t m = (hi - lo > step) ? lo + step - 1 : hi - 1;
      if (pred(m)) {
        lo = m + 1;
        step <<= 1;
      } else {
        hi = m;
        break;
      }
    }
  } else {  // answer in [lo, g]
    hi = g;
    uint32_t step = 1;
    while (hi > lo) {
      const uint32_t m = (hi - lo > step) ? hi - step : lo;
      if (!pred(m)) {
        hi = m;
        step <<= 1;
      } else {
        lo = m + 1;
        break;
      }
    }
  }
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const bool t = pred(mid);
    lo = t ? mid + 1 : lo;
    hi = t ? hi : mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t ck_partition(const Checkpoint *__restrict__ ck, bool swap_id, uint32_t lo,
                                                 uint32_t hi, int64_t x, bool inclusive, uint32_t guess) {
  if (lo >= hi) return lo;
  const uint32_t g = ck_clamp(guess, lo, hi);
  return ck_partition_s(ck, swap_id, lo, hi, x, inclusive, g, ck_load_sector(ck, g));
}

struct EndsAcc {
  // the first / last overlapping op of the walk so far: packed run + walk position before it
  uint32_t f_v, l_v;
  int32_t f_tp, f_qp, l_tp, l_qp;
  uint32_t n_ov;  // overlapping ops seen (r_ov accounting)
  bool found, broke;
};

// evaluates the (up to) 8 runs of physical block pb in walk order; only the per-op overlap test runs
// for every op — the projection arithmetic is done once, by ends_first / ends_last, on the selected ops
// the two 128-bit halves of physical block pb
struct RunBlock {
  uint4 lo, hi;
};
__device__ __forceinline__ RunBlock load_block(const uint32_t *__restrict__ blk, uint32_t pb) {
  const uint4 *src = reinterpret_cast<const uint4 *>(blk + pb * RUNS_PER_BLOCK);
  RunBlock b;
  b.lo = ld_pipe_u4(src);
  b.hi = ld_pipe_u4(src + 1);
  return b;
}
__device__ __forceinline__ void thread_eval_block(const RunBlock &rb, uint32_t n, uint32_t pb, uint32_t op_t0,
                                                  uint32_t op_q0, bool backward, int32_t dir, int32_t tp, int32_t qp,
                                                  int32_t rs, int32_t re, int32_t last_target_pos, EndsAcc &acc) {
  const uint32_t base = pb * RUNS_PER_BLOCK;
  const uint32_t cnt = min((uint32_t)RUNS_PER_BLOCK, n - base);
#pragma unroll
  for (int c = 0; c < RUNS_PER_BLOCK / 4; c++) {
    const int ci = backward ? RUNS_PER_BLOCK / 4 - 1 - c : c;
    uint4 v4 = ci ? rb.hi : rb.lo;
    if (backward) {  // walk order within the chunk is reversed too
      uint32_t t0 = v4.x, t1 = v4.y;
      v4.x = v4.w; v4.y = v4.z; v4.z = t1; v4.w = t0;
    }
    const uint32_t v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t off = (uint32_t)(ci * 4 + (backward ? 3 - k : k));  // physical offset in the block
      const uint32_t op = v[k] >> 29;
      const int32_t len = (int32_t)(v[k] & 0x1fffffffu);
      const int32_t t = (op == op_t0) ? 0 : len;  // op_t0: the op without target length in this orientation
      const int32_t q = (op == op_q0) ? 0 : len;
      const bool valid = off < cnt;
      const bool processed = valid && tp <= last_target_pos;
      acc.broke |= valid && tp > last_target_pos;
      const bool is_ins = t == 0, is_del = !is_ins && q == 0;
      const int32_t os = max(tp, rs);
      const int32_t oe = min(tp + t, is_del ? last_target_pos : re);
      const bool ov = processed && (is_ins ? tp >= rs : os < oe);
      const bool first = ov && !acc.found;
      acc.f_v = first ? v[k] : acc.f_v;
      acc.f_tp = first ? tp : acc.f_tp;
      acc.f_qp = first ? qp : acc.f_qp;
      acc.found |= ov;
      acc.l_v = ov ? v[k] : acc.l_v;
      acc.l_tp = ov ? tp : acc.l_tp;
      acc.l_qp = ov ? qp : acc.l_qp;
      acc.n_ov += ov ? 1u : 0u;
      tp += t;
      qp += dir > 0 ? q : -q;
    }
  }
}
// start of the projection = start of the first overlapping op (src/impg.rs:2806-2868)
__device__ __forceinline__ void ends_first(const EndsAcc &acc, uint32_t op_t0, uint32_t op_q0, int32_t dir, int32_t rs,
                                           int32_t &f_q, int32_t &f_t) {
  const uint32_t op = acc.f_v >> 29;
  const int32_t len = (int32_t)(acc.f_v & 0x1fffffffu);
  const int32_t t = (op == op_t0) ? 0 : len, q = (op == op_q0) ? 0 : len;
  const bool is_ins = t == 0, is_del = !is_ins && q == 0;
  const int32_t os = max(acc.f_tp, rs);
  f_t = is_ins ? acc.f_tp : os;
  f_q = (is_ins || is_del) ? acc.f_qp : acc.f_qp + (os - acc.f_tp) * dir;
}
// end of the projection = end of the last overlapping op
__device__ __forceinline__ void ends_last(const EndsAcc &acc, uint32_t op_t0, uint32_t op_q0, int32_t dir, int32_t rs,
                                          int32_t re, int32_t last_target_pos, int32_t &l_q, int32_t &l_t) {
  const uint32_t op = acc.l_v >> 29;
  const int32_t len = (int32_t)(acc.l_v & 0x1fffffffu);
  const int32_t t = (op == op_t0) ? 0 : len, q = (op == op_q0) ? 0 : len;
  const bool is_ins = t == 0, is_del = !is_ins && q == 0;
  const int32_t tp = acc.l_tp, qp = acc.l_qp;
  const int32_t os = max(tp, rs);
  const int32_t oe = min(tp + t, is_del ? last_target_pos : re);
  const int32_t pqs = (is_ins || is_del) ? qp : qp + (os - tp) * dir;
  l_t = is_ins ? tp : oe;
  l_q = is_ins ? qp + q * dir : (is_del ? qp : pqs + (oe - os) * dir);
}

// Where the accepted hits of the last hop of the direct BED path go: straight into their
// (row, query sequence) bucket as 32-byte BoxRecs (slot = atomic cursor of the bucket), so the
// segment merge reads contiguous memory and no global sort is needed.
struct BucketOut {
  uint32_t *cursor;      // rows x n_seqs, initialised with the bucket begins
  BoxRec *boxes;
  const uint32_t *orig;  // processing index -> frontier index of the hop, or nullptr
  const uint32_t *gmap;  // sharded index: local frontier index -> index in the GLOBAL frontier, or nullptr
  uint32_t level;        // ord level of the hop
};

// One endpoint liftover: project_target_range_through_alignment (src/impg.rs:2760-2898) of the range `f` through tree
// entry `entry`, by one thread; only the first and the last overlapping op are located (checkpoint search from an
// interpolated guess + one or two run blocks per end).
struct EndsHit {
  bool ok;
  uint32_t query_id, vrank;
  int32_t f_q, l_q, f_t, l_t;
  uint32_t nread, nck, r_ov;  // runs read, checkpoints probed, runs inside the window (roofline accounting)
};

// OVL: the dependent gathers of a hit are issued in two waves instead of four steps — both checkpoint sectors
// (the end-side guess is absolute instead of relative to the start-side result), then the checkpoints and run
// blocks of both ends — so a hit waits for three memory round trips (entry, sectors, blocks) instead of five.
template <bool OVL>
__device__ __forceinline__ EndsHit lift_ends_hit(const DevIndexView &ix, const Frontier &f, uint32_t entry,
                                                 const LiftParams &lp) {
  const uint4 *rp = reinterpret_cast<const uint4 *>(ix.e_rec + entry);
  const uint4 r0 = rp[0], r1 = rp[1];
  const int32_t t_start = (int32_t)r0.x, t_end = (int32_t)r0.y, q_start = (int32_t)r0.z, q_end = (int32_t)r0.w;
  const uint32_t query_id = r1.x, nruns_flags = r1.y, aln_off = r1.z, vrank = r1.w;
  int32_t rs = f.start, re = f.end;
  if (lp.clip) {
    rs = max(rs, t_start);
    re = min(re, t_end);
  }
  const uint32_t n = nruns_flags >> 2;
  const bool rev_strand = nruns_flags & FLAG_STRAND;
  const bool swap_id = nruns_flags & FLAG_REVERSED;
  const bool backward = swap_id && rev_strand;
  const int32_t dir = rev_strand ? -1 : 1;
  const uint32_t nblk = aln_nblk(n);
  const Checkpoint *ck = aln_ck(ix.stream, aln_off);
  const uint32_t *blk = aln_runs(ix.stream, aln_off, nblk);
  const int32_t last_target_pos = min(t_end, re);
  const int64_t rel = (int64_t)rs - t_start, rel_l = (int64_t)last_target_pos - t_start;

  // in this orientation: the op that consumes no target (insertion) / no query (deletion)
  const uint32_t op_t0 = swap_id ? IMPGX_OP_D : IMPGX_OP_I, op_q0 = swap_id ? IMPGX_OP_I : IMPGX_OP_D;
  EndsAcc acc;
  acc.f_v = acc.l_v = 0;
  acc.f_tp = acc.f_qp = acc.l_tp = acc.l_qp = 0;
  acc.n_ov = 0;
  acc.found = false;
  acc.broke = false;
  int32_t f_q = -1, f_t = -1, l_q = -1, l_t = -1;
  uint32_t nread = 0, nck = 0, r_ov = 1;
  if (nblk > 0 && rel_l >= 0) {
    // totals: needed exactly only when walking backwards; otherwise the record's
    // own span is a good enough denominator for the interpolation guess
    int64_t w_tot = (int64_t)t_end - t_start, wq_tot = 0;
    if (backward) {
      const Checkpoint tot = ck[nblk];
      w_tot = swap_id ? tot.q_off : tot.t_off;
      wq_tot = swap_id ? tot.t_off : tot.q_off;
    }
    const float scale = __fdividef((float)nblk, (float)(w_tot > 0 ? w_tot : 1));  // only seeds the search
    // walk blocks j in [0, nblk): js = last block starting before rs, je = last block starting at or before L
    uint32_t js, je;
    if (OVL) {
      // both lookups search the same range, so their sectors are requested together. Forward: the point of
      // (P <= rel_l) over [0, nblk) is the point over [js, nblk) because rel <= rel_l puts it at or above js.
      const uint32_t lo = backward ? 1u : 0u, hi = backward ? nblk + 1 : nblk;
      const int64_t x0 = backward ? w_tot - rel : rel, x1 = backward ? w_tot - rel_l : rel_l;
      const uint32_t g0 = ck_clamp(x0 > 0 ? (uint32_t)((float)x0 * scale) : 0u, lo, hi);
      const uint32_t g1 = ck_clamp((x1 > 0 ? (uint32_t)((float)x1 * scale) : 0u) + (backward ? 0u : 1u), lo, hi);
      const CkSector s0 = ck_load_sector(ck, g0), s1 = ck_load_sector(ck, g1);
      const uint32_t p0 = ck_partition_s(ck, swap_id, lo, hi, x0, backward, g0, s0);
      const uint32_t p1 = ck_partition_s(ck, swap_id, lo, hi, x1, !backward, g1, s1);
      const uint32_t a = backward ? nblk - (p0 - 1) : p0, b = backward ? nblk - (p1 - 1) : p1;
      js = a ? a - 1 : 0;
      je = b ? b - 1 : 0;
    } else if (!backward) {
      const uint32_t g0 = rel > 0 ? (uint32_t)((float)rel * scale) : 0u;
      const uint32_t a = ck_partition(ck, swap_id, 0, nblk, rel, false, g0);
      js = a ? a - 1 : 0;
      const uint32_t g1 = js + (uint32_t)((float)(rel_l - (rel > 0 ? rel : 0)) * scale) + 1;
      const uint32_t b = ck_partition(ck, swap_id, js, nblk, rel_l, true, g1);
      je = b ? b - 1 : 0;
    } else {
      // walk block j starts at W - P(nblk - j):  #{j : W - P(nblk-j) < rel} = nblk - #{b in [1,nblk] : P(b) <= W - rel}
      const int64_t x0 = w_tot - rel, x1 = w_tot - rel_l;
      const uint32_t g0 = x0 > 0 ? (uint32_t)((float)x0 * scale) : 0u;
      const uint32_t a = nblk - (ck_partition(ck, swap_id, 1, nblk + 1, x0, true, g0) - 1);
      const uint32_t g1 = x1 > 0 ? (uint32_t)((float)x1 * scale) : 0u;
      const uint32_t b = nblk - (ck_partition(ck, swap_id, 1, nblk + 1, x1, false, g1) - 1);
      js = a ? a - 1 : 0;
      je = b ? b - 1 : 0;
    }
    nck = 8;  // ~2 sectors of checkpoints per lookup (galloping from an interpolated guess)
    // ---- start side: walk forward from block js until the first overlap (or the loop break)
    uint32_t j = js;
    // OVL: the checkpoint and the run block of the end side are requested together with those of the start side
    const uint32_t jl0 = je > js ? je : js;
    uint2 kE = make_uint2(0, 0);
    RunBlock rbE;
    rbE.lo = rbE.hi = make_uint4(0, 0, 0, 0);
    uint2 kS = make_uint2(0, 0);
    RunBlock rbS;
    rbS.lo = rbS.hi = make_uint4(0, 0, 0, 0);
    if (OVL) {
      const uint32_t pbs = backward ? nblk - 1 - js : js;
      kS = ld_pipe_u2(backward ? ck + pbs + 1 : ck + pbs);
      rbS = load_block(blk, pbs);
      if (jl0 != js) {
        const uint32_t pbe = backward ? nblk - 1 - jl0 : jl0;
        kE = ld_pipe_u2(backward ? ck + pbe + 1 : ck + pbe);
        rbE = load_block(blk, pbe);
      }
    }
    for (;;) {
      const uint32_t pb = backward ? nblk - 1 - j : j;
      const bool pre = OVL && j == js;
      const uint2 kk = pre ? kS : ld_pipe_u2(backward ? ck + pb + 1 : ck + pb);
      const int64_t pt = swap_id ? kk.y : kk.x, pq = swap_id ? kk.x : kk.y;
      const int64_t tcons = backward ? w_tot - pt : pt, qcons = backward ? wq_tot - pq : pq;
      const int32_t tp0 = (int32_t)(t_start + tcons);
      const int32_t qp0 = rev_strand ? (int32_t)(q_end - qcons) : (int32_t)(q_start + qcons);
      const RunBlock rb = pre ? rbS : load_block(blk, pb);
      thread_eval_block(rb, n, pb, op_t0, op_q0, backward, dir, tp0, qp0, rs, re, last_target_pos, acc);
      nread += RUNS_PER_BLOCK;
      if (acc.found || acc.broke || j + 1 >= nblk) break;
      j++;
    }
    // ---- end side: the last overlapping op lies in blocks [j, max(je, j)]; walk backward from the top
    if (acc.found) {
      uint32_t jl = je > j ? je : j;
      r_ov = acc.n_ov;
      ends_first(acc, op_t0, op_q0, dir, rs, f_q, f_t);
      ends_last(acc, op_t0, op_q0, dir, rs, re, last_target_pos, l_q, l_t);
      while (jl > j) {
        const uint32_t pb = backward ? nblk - 1 - jl : jl;
        const bool pre = OVL && jl == jl0 && jl0 != js;
        const uint2 kk = pre ? kE : ld_pipe_u2(backward ? ck + pb + 1 : ck + pb);
        const int64_t pt = swap_id ? kk.y : kk.x, pq = swap_id ? kk.x : kk.y;
        const int64_t tcons = backward ? w_tot - pt : pt, qcons = backward ? wq_tot - pq : pq;
        const int32_t tp0 = (int32_t)(t_start + tcons);
        const int32_t qp0 = rev_strand ? (int32_t)(q_end - qcons) : (int32_t)(q_start + qcons);
        const RunBlock rb = pre ? rbE : load_block(blk, pb);
        EndsAcc a2;
        a2.f_v = a2.l_v = 0;
        a2.f_tp = a2.f_qp = a2.l_tp = a2.l_qp = 0;
        a2.n_ov = 0;
        a2.found = false;
        a2.broke = false;
        thread_eval_block(rb, n, pb, op_t0, op_q0, backward, dir, tp0, qp0, rs, re, last_target_pos, a2);
        nread += RUNS_PER_BLOCK;
        if (a2.found) {
          ends_last(a2, op_t0, op_q0, dir, rs, re, last_target_pos, l_q, l_t);
          // ops of block j from its first overlap on, whole blocks in between, ops of block jl up to its last overlap
          r_ov = acc.n_ov + (jl - j - 1) * RUNS_PER_BLOCK + a2.n_ov;
          break;
        }
        jl--;  // nothing overlapped up there: try the block below (block j already holds its own last)
      }
    }
  }
  bool ok = acc.found && f_q != l_q && f_t != l_t;
  if (ok && lp.subset) ok = query_id == lp.row_target[f.row] || lp.subset[query_id] != 0;
  EndsHit out;
  out.ok = ok;
  out.query_id = query_id;
  out.vrank = vrank;
  out.f_q = f_q; out.l_q = l_q; out.f_t = f_t; out.l_t = l_t;
  out.nread = nread; out.nck = nck; out.r_ov = r_ov;
  return out;
}

template <bool BUCKET, bool OVL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_liftover_ends(DevIndexView ix, const Frontier *__restrict__ fr,
                                                       const LiftTask *__restrict__ tasks, uint64_t n_tasks,
                                                       LiftParams lp, Hit *__restrict__ hits,
                                                       unsigned long long *__restrict__ counters, BucketOut bo) {
  unsigned long long runs_acc = 0, ok_acc = 0, ck_acc = 0, rov_acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_tasks; w += stride) {
    const LiftTask t = tasks[w];
    const Frontier f = fr[t.range];
    const EndsHit eh = lift_ends_hit<OVL>(ix, f, t.entry, lp);
    bool ok = eh.ok;
    const uint32_t query_id = eh.query_id, vrank = eh.vrank;
    const int32_t f_q = eh.f_q, l_q = eh.l_q, f_t = eh.f_t, l_t = eh.l_t;
    const uint32_t nread = eh.nread, nck = eh.nck, r_ov = eh.r_ov;
    runs_acc += nread;
    ck_acc += nck;
    rov_acc += r_ov;
    ok_acc += ok ? 1 : 0;
    if (BUCKET) {
      if (ok && lp.min_output_len >= 0) {
        const long long dq = (long long)l_q - (long long)f_q;
        ok = (dq < 0 ? -dq : dq) >= lp.min_output_len;
      }
      if (ok) {
        uint32_t r = bo.orig ? bo.orig[t.range] : t.range;
        r = bo.gmap ? bo.gmap[r] : r;
        const uint32_t slot = atomicAdd(bo.cursor + (uint64_t)f.row * ix.n_seqs + query_id, 1u);
        const uint64_t ord = ((uint64_t)bo.level << 58) | ((uint64_t)r << 32) | vrank;
        uint4 *dst = reinterpret_cast<uint4 *>(bo.boxes + slot);
        dst[0] = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)f_q, (uint32_t)l_q);
        dst[1] = make_uint4(f.seq, (uint32_t)f_t, (uint32_t)l_t, 0u);
      }
    } else {
      Hit h;
      h.row = ok ? f.row : INVALID_ID;
      h.q_id = query_id;
      h.q_first = f_q;
      h.q_last = l_q;
      h.t_id = f.seq;
      h.t_first = f_t;
      h.t_last = l_t;
      h.vrank = vrank;
      hits[w] = h;
    }
  }
  // one atomic per warp
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    runs_acc += __shfl_xor_sync(FULL, runs_acc, d);
    ok_acc += __shfl_xor_sync(FULL, ok_acc, d);
    ck_acc += __shfl_xor_sync(FULL, ck_acc, d);
    rov_acc += __shfl_xor_sync(FULL, rov_acc, d);
  }
  if (lane_id() == 0 && counters) {
    atomicAdd(&counters[0], runs_acc);
    atomicAdd(&counters[1], ok_acc);
    atomicAdd(&counters[2], ck_acc);
    atomicAdd(&counters[3], rov_acc);
  }
}

// Writes the clipped CIGAR slices (projected_cigar_ops, src/impg.rs:2878-2886)
// of accepted hits in walk order with the reference's op inversion applied.
__global__ void __launch_bounds__(256) k_emit_cigar(DevIndexView ix, const LiftTask *__restrict__ tasks,
                                                    const CigarSlice *__restrict__ slices,
                                                    const uint64_t *__restrict__ out_off, uint64_t n_tasks,
                                                    uint32_t *__restrict__ out) {
  const unsigned lane = lane_id();
  uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n_tasks; w += nw) {
    const CigarSlice s = slices[w];
    if (s.n_ops == 0) continue;
    const EntryRec rec = ix.e_rec[tasks[w].entry];
    const uint32_t n = rec.nruns_flags >> 2;
    const bool swap_id = rec.nruns_flags & FLAG_REVERSED;
    const bool backward = swap_id && (rec.nruns_flags & FLAG_STRAND);
    const uint32_t *blk = aln_runs(ix.stream, rec.aln_off, aln_nblk(n));
    uint32_t *dst = out + out_off[w];
    for (uint32_t k = lane; k < s.n_ops; k += 32) {
      uint32_t wi = s.first_idx + k;
      uint32_t v = blk[backward ? (n - 1 - wi) : wi];
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
