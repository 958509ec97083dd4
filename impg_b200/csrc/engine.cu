// engine.cu — index upload, the batched query pipeline (stab -> liftover ->
// fold -> frontier, per BFS level) and result assembly.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

#include <functional>

#include "bfs_kernels.cuh"
#include "comm.cuh"
#include "engine.cuh"
#include "bucket_kernels.cuh"
#include "merge_kernels.cuh"
#include "shard_kernels.cuh"
#include "small_bfs.cuh"

namespace impgx {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
const char *last_error() { return g_last_error.c_str(); }

static uint64_t env_u64(const char *name, uint64_t dflt);
static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

// grid sized as a multiple of the SM count (148 on B200): `per_sm` resident
// CTAs per SM, capped by the work available.
static unsigned grid_for(uint64_t work_items, unsigned threads, unsigned items_per_thread_group, unsigned per_sm) {
  uint64_t groups = (work_items + items_per_thread_group - 1) / items_per_thread_group;
  uint64_t blocks = (groups * 1 + (threads - 1)) / threads;
  uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
// one thread per item
static unsigned grid_threads(uint64_t n, unsigned threads = 256, unsigned per_sm = 8) {
  uint64_t blocks = (n + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}
// one warp per item
static unsigned grid_warps(uint64_t n, unsigned threads = 256, unsigned per_sm = 8) {
  uint64_t blocks = (n * 32 + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// ----------------------------------------------------------------- pinned pool
namespace {
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<void *, size_t>> free_list;
  ~PinnedPool() {
    for (auto &b : free_list) cudaFreeHost(b.first);
  }
};
PinnedPool &pinned_pool() {
  static PinnedPool p;
  return p;
}
}  // namespace

void *pinned_acquire(size_t bytes, size_t *cap_out) {
  size_t cap = 4096;
  while (cap < bytes) cap <<= 1;
  {
    PinnedPool &pool = pinned_pool();
    std::lock_guard<std::mutex> lock(pool.mu);
    for (size_t i = 0; i < pool.free_list.size(); i++) {
      if (pool.free_list[i].second == cap) {
        void *p = pool.free_list[i].first;
        pool.free_list.erase(pool.free_list.begin() + i);
        *cap_out = cap;
        return p;
      }
    }
  }
  void *p = nullptr;
  if (cudaMallocHost(&p, cap) != cudaSuccess) {
    cudaGetLastError();
    throw Error(IMPGX_E_NOMEM, "pinned host allocation of " + std::to_string(cap) + " bytes failed");
  }
  *cap_out = cap;
  return p;
}

void pinned_release(void *p, size_t cap) {
  PinnedPool &pool = pinned_pool();
  std::lock_guard<std::mutex> lock(pool.mu);
  if (pool.free_list.size() >= 64) {
    cudaFreeHost(p);
    return;
  }
  pool.free_list.emplace_back(p, cap);
}

// ----------------------------------------------------------------- CUB glue
// CUB temp storage: allocated per call from the arena (stack discipline frees it at once)
struct Scratch {
  Arena *a = nullptr;
};

// Small inputs (single-row calls: partition's windows, refine's candidates) are bound by the
// number of launches, not by bandwidth: one CTA does the whole scan / sort in shared memory
// instead of CUB's multi-kernel pipelines.
constexpr uint32_t SMALL_SCAN_MAX = 4096;
constexpr uint32_t SMALL_SORT_MAX = 2048;
constexpr int SMALL_SORT_IDX_BITS = 11;

__global__ void __launch_bounds__(1024) k_small_excl_scan_u64(uint64_t *__restrict__ a, uint32_t n) {
  __shared__ uint64_t warp_tot[32];
  uint64_t v[4];
  uint64_t sum = 0;
  const uint32_t base = threadIdx.x * 4;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    v[u] = base + u < n ? a[base + u] : 0;
    sum += v[u];
  }
  uint64_t incl = sum;
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (unsigned)d) incl += o;
  }
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint64_t t = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
    uint64_t ti = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint64_t o = __shfl_up_sync(0xffffffffu, ti, d);
      if (lane >= (unsigned)d) ti += o;
    }
    warp_tot[lane] = ti - t;
  }
  __syncthreads();
  uint64_t run = warp_tot[w] + incl - sum;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    if (base + u < n) a[base + u] = run;
    run += v[u];
  }
}

// stable sort of (key, value) pairs on the key bits [begin_bit, end_bit): bitonic network over
// (masked key, input position), so equal keys keep their input order like the LSD radix sort
template <class K>
__global__ void __launch_bounds__(1024) k_small_sort_pairs(K *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t n,
                                                           uint32_t P, int begin_bit, int end_bit) {
  __shared__ uint64_t sk[SMALL_SORT_MAX];
  __shared__ K k0[SMALL_SORT_MAX];
  __shared__ uint32_t v0[SMALL_SORT_MAX];
  const int nb = end_bit - begin_bit;
  const uint64_t mask = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < n) {
      const K k = keys[i];
      k0[i] = k;
      v0[i] = vals[i];
      sk[i] = ((((uint64_t)k >> begin_bit) & mask) << SMALL_SORT_IDX_BITS) | i;
    } else {
      sk[i] = ~0ull;
    }
  }
  __syncthreads();
  const uint32_t half = P >> 1;
  for (uint32_t k = 2; k <= P; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < half; i += blockDim.x) {
        const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
        const uint64_t x = sk[lo], y = sk[hi];
        if ((y < x) == ((lo & k) == 0)) {
          sk[lo] = y;
          sk[hi] = x;
        }
      }
      __syncthreads();
    }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t src = (uint32_t)(sk[i] & ((1u << SMALL_SORT_IDX_BITS) - 1));
    keys[i] = k0[src];
    vals[i] = v0[src];
  }
}

// callers count two launches per scan (CUB: init + scan); the single-CTA scan is one
static thread_local uint64_t tl_small_scans = 0;

static bool small_paths_enabled() {
  static const bool on = !getenv("IMPGX_NO_SMALL_PATHS");
  return on;
}

// in-place exclusive sum over n+1 elements: a[n] becomes the total
static void exclusive_scan_u64(uint64_t *a, uint64_t n_plus_1, Scratch &sc, cudaStream_t s) {
  if (n_plus_1 <= SMALL_SCAN_MAX && small_paths_enabled()) {
    const unsigned threads = (unsigned)std::min<uint64_t>(1024, ((n_plus_1 + 3) / 4 + 31) / 32 * 32);
    k_small_excl_scan_u64<<<1, threads, 0, s>>>(a, (uint32_t)n_plus_1);
    CUDA_CHECK(cudaGetLastError());
    tl_small_scans++;
    return;
  }
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, a, a, n_plus_1, s));
  DBuf<uint8_t> tmp(bytes, *sc.a);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, a, a, n_plus_1, s));
}

// exclusive sum over n u32 values into `out` (may alias `in`); the caller appends a zero to get the total
static void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, Scratch &sc, cudaStream_t s) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  DBuf<uint8_t> tmp(bytes, *sc.a);
  CUDA_CHECK(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, in, out, n, s));
}

static int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b)) b++;
  return b;
}

template <class K>
static void sort_pairs(DBuf<K> &keys, DBuf<uint32_t> &vals, uint64_t n, int begin_bit, int end_bit, Scratch &sc,
                       cudaStream_t s, Ctx &ctx) {
  if (n == 0) return;
  if (n <= SMALL_SORT_MAX && end_bit - begin_bit + SMALL_SORT_IDX_BITS <= 64 && small_paths_enabled()) {
    uint32_t P = 2;
    while (P < n) P <<= 1;
    const unsigned threads = std::min(1024u, std::max(32u, P >> 1));
    k_small_sort_pairs<K><<<1, threads, 0, s>>>(keys.get(), vals.get(), (uint32_t)n, P, begin_bit, end_bit);
    CUDA_CHECK(cudaGetLastError());
    ctx.launches += 1;
    return;
  }
  DBuf<K> k2(n, *sc.a);
  DBuf<uint32_t> v2(n, *sc.a);
  cub::DoubleBuffer<K> dk(keys.get(), k2.get());
  cub::DoubleBuffer<uint32_t> dv(vals.get(), v2.get());
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, begin_bit, end_bit, s));
  {
    DBuf<uint8_t> tmp(bytes, *sc.a);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, dk, dv, n, begin_bit, end_bit, s));
  }
  ctx.launches += 2 + (uint64_t)((end_bit - begin_bit + 7) / 8);
  // sorted data may sit in the alternate buffers: copy back so the caller's
  // (older, lower) blocks stay and the temporaries above them can be popped
  if (dk.Current() != keys.get()) CUDA_CHECK(cudaMemcpyAsync(keys.get(), k2.get(), n * sizeof(K), cudaMemcpyDeviceToDevice, s));
  if (dv.Current() != vals.get()) CUDA_CHECK(cudaMemcpyAsync(vals.get(), v2.get(), n * 4, cudaMemcpyDeviceToDevice, s));
}

// Size readbacks between the count / scan / fill stages land in a small pinned buffer (a copy into
// pageable memory is staged by the driver and costs several microseconds more per call).
static uint64_t *readback_slot() {
  static thread_local uint64_t *slot = nullptr;
  if (!slot && cudaMallocHost((void **)&slot, 64) != cudaSuccess) {
    cudaGetLastError();
    throw Error(IMPGX_E_NOMEM, "pinned host allocation failed");
  }
  return slot;
}

static uint64_t read_u64(const uint64_t *d, cudaStream_t s, Ctx &ctx) {
  uint64_t *h = readback_slot();
  CUDA_CHECK(cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx.d2h_bytes += 8;
  return h[0];
}

// two sizes, one synchronisation
static void read_u64x2(const uint64_t *d0, const uint64_t *d1, uint64_t &v0, uint64_t &v1, cudaStream_t s, Ctx &ctx) {
  uint64_t *h = readback_slot();
  CUDA_CHECK(cudaMemcpyAsync(h, d0, 8, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaMemcpyAsync(h + 1, d1, 8, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx.d2h_bytes += 16;
  v0 = h[0];
  v1 = h[1];
}

struct WallTimer {
  double &acc;
  std::chrono::steady_clock::time_point t0;
  explicit WallTimer(double &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~WallTimer() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

#define LAUNCH(kernel, grid, block, stream, ...)          \
  do {                                                    \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); \
    CUDA_CHECK(cudaGetLastError());                       \
    ctx.launches++;                                       \
  } while (0)

// ----------------------------------------------------------------- index
}  // namespace impgx

impgx_index::~impgx_index() {
  cudaSetDevice(device);
  cudaFree(d_start); cudaFree(d_end); cudaFree(d_pmax); cudaFree(d_seq_len);
  cudaFree(d_stream); cudaFree(d_rec); cudaFree(d_tgt_off); cudaFree(d_owner); cudaFree(d_qid); cudaFree(d_qorder);
  for (cudaStream_t st : stream_pool) cudaStreamDestroy(st);
}

namespace impgx {

template <class T>
static T *upload(const std::vector<T> &h, uint64_t &bytes_acc) {
  T *d = nullptr;
  // 16 bytes of slack: the stab kernels' bulk copies round the last tile up to 16 bytes
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T) + 16;
  CUDA_CHECK(cudaMalloc((void **)&d, bytes));
  if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  bytes_acc += bytes;
  return d;
}

void check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw Error(IMPGX_E_NO_DEVICE, std::string("no CUDA device available (libimpgx has no CPU fallback): ") +
                                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  REQUIRE(device >= 0 && device < n, IMPGX_E_INVALID, "device ordinal out of range");
  CUDA_CHECK(cudaSetDevice(device));
  // keep freed stream-ordered allocations cached in the pool between calls
  cudaMemPool_t pool;
  CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
}

impgx_index *index_build(const impgx_record *recs, size_t n, const uint32_t *runs, const uint64_t *run_offsets,
                         const uint64_t *seq_lens, uint32_t n_seqs, bool bidirectional, int device,
                         const uint32_t *owner, uint32_t rank, uint32_t n_ranks) {
  check_device(device);
  REQUIRE(recs || n == 0, IMPGX_E_INVALID, "records is NULL");
  REQUIRE(run_offsets && seq_lens, IMPGX_E_INVALID, "run_offsets / seq_lens is NULL");
  for (uint32_t s = 0; s < n_seqs; s++)
    REQUIRE(seq_lens[s] <= (uint64_t)INT32_MAX, IMPGX_E_INVALID, "sequence longer than 2^31-1 (coordinates are i32)");
  for (size_t i = 0; i < n; i++) REQUIRE(run_offsets[i] <= run_offsets[i + 1], IMPGX_E_INVALID, "run_offsets not monotone");
  if (owner) {
    REQUIRE(n_ranks >= 1 && rank < n_ranks, IMPGX_E_INVALID, "bad shard rank / n_ranks");
    for (uint32_t s = 0; s < n_seqs; s++) REQUIRE(owner[s] < n_ranks, IMPGX_E_INVALID, "owner id out of range");
  }

  std::unique_ptr<impgx_index> idx(new impgx_index());
  idx->device = device;
  idx->n_seqs = n_seqs;
  idx->n_records = n;
  idx->seq_lens.assign(seq_lens, seq_lens + n_seqs);
  idx->names.resize(n_seqs);

  HostColumns hc;
  build_host_columns(recs, n, run_offsets, n_seqs, bidirectional, hc, owner, rank);
  idx->n_entries = hc.e_start.size();
  idx->n_blocks = hc.aln_off[n];

  uint64_t bytes = 0;
  idx->d_start = upload(hc.e_start, bytes);
  idx->d_end = upload(hc.e_end, bytes);
  idx->d_pmax = upload(hc.e_pmax, bytes);
  idx->d_rec = upload(hc.e_rec, bytes);
  idx->d_qid = upload(hc.e_qid, bytes);
  idx->d_tgt_off = upload(hc.tgt_off, bytes);
  std::vector<int32_t> sl(n_seqs);
  for (uint32_t s = 0; s < n_seqs; s++) sl[s] = (int32_t)seq_lens[s];
  idx->d_seq_len = upload(sl, bytes);
  if (owner) {
    idx->owner.assign(owner, owner + n_seqs);
    idx->shard_rank = rank;
    idx->shard_size = n_ranks;
    idx->d_owner = upload(idx->owner, bytes);
    idx->q_order.resize(n_seqs);
    idx->q_first.assign((size_t)n_ranks + 1, 0);
    for (uint32_t s2 = 0; s2 < n_seqs; s2++) idx->q_first[owner[s2] + 1]++;
    for (uint32_t r = 0; r < n_ranks; r++) idx->q_first[r + 1] += idx->q_first[r];
    {
      std::vector<uint32_t> cur(idx->q_first.begin(), idx->q_first.end() - 1);
      for (uint32_t s2 = 0; s2 < n_seqs; s2++) idx->q_order[cur[owner[s2]]++] = s2;
    }
    idx->d_qorder = upload(idx->q_order, bytes);
  }

  // stream: per alignment checkpoints + 8-run blocks, built on the device from
  // the raw runs, uploaded in bounded chunks of alignments. A shard uploads only
  // the alignments its entries walk (gathered through a host staging buffer).
  size_t stream_bytes = std::max<uint64_t>(idx->n_blocks, 1) * 32;
  CUDA_CHECK(cudaMalloc((void **)&idx->d_stream, stream_bytes));
  bytes += stream_bytes;
  {
    const uint64_t chunk_runs = 256ull << 20;  // 1 GiB of raw runs per chunk
    uint32_t *d_raw = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t *d_blk = nullptr;
    int *d_bad = nullptr;
    CUDA_CHECK(cudaMalloc((void **)&d_bad, 4));
    CUDA_CHECK(cudaMemset(d_bad, 0, 4));
    size_t cap_raw = 0, cap_aln = 0;
    std::vector<uint64_t> rel;
    std::vector<uint32_t> blk, stage;
    Ctx ctx;
    size_t a = 0;
    while (a < n) {
      // chunk [a, b) of the records; `rel`/`blk` list the alignments of the chunk that are uploaded
      rel.assign(1, 0);
      blk.clear();
      uint64_t nr = 0;
      size_t b = a;
      while (b < n) {
        const bool needed = hc.aln_off[b + 1] != hc.aln_off[b];
        const uint64_t r = needed ? run_offsets[b + 1] - run_offsets[b] : 0;
        if (nr + r > chunk_runs && !blk.empty()) break;
        if (needed) {
          nr += r;
          rel.push_back(nr);
          blk.push_back(hc.aln_off[b]);
        }
        b++;
      }
      const size_t na = blk.size();
      if (na) {
        if (nr > cap_raw) {
          cudaFree(d_raw);
          CUDA_CHECK(cudaMalloc((void **)&d_raw, std::max<uint64_t>(nr, 1) * 4));
          cap_raw = nr;
        }
        if (na > cap_aln) {
          cudaFree(d_off);
          cudaFree(d_blk);
          CUDA_CHECK(cudaMalloc((void **)&d_off, (na + 1) * 8));
          CUDA_CHECK(cudaMalloc((void **)&d_blk, (na + 1) * 4));
          cap_aln = na;
        }
        if (nr) {
          if (na == b - a) {  // every alignment of the chunk: one contiguous copy
            CUDA_CHECK(cudaMemcpy(d_raw, runs + run_offsets[a], nr * 4, cudaMemcpyHostToDevice));
          } else {
            stage.resize(nr);
            size_t k = 0;
            for (size_t i = a; i < b; i++) {
              if (hc.aln_off[i + 1] == hc.aln_off[i]) continue;
              const uint64_t r = run_offsets[i + 1] - run_offsets[i];
              memcpy(stage.data() + rel[k], runs + run_offsets[i], r * 4);
              k++;
            }
            CUDA_CHECK(cudaMemcpy(d_raw, stage.data(), nr * 4, cudaMemcpyHostToDevice));
          }
        }
        CUDA_CHECK(cudaMemcpy(d_off, rel.data(), (na + 1) * 8, cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(d_blk, blk.data(), na * 4, cudaMemcpyHostToDevice));
        LAUNCH(k_build_blocks, grid_warps(na), 256, 0, d_raw, d_off, d_blk, (uint64_t)na, idx->d_stream, d_bad);
        CUDA_CHECK(cudaDeviceSynchronize());
      }
      a = b;
    }
    int bad = 0;
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    cudaFree(d_raw);
    cudaFree(d_off);
    cudaFree(d_blk);
    cudaFree(d_bad);
    REQUIRE(bad == 0, IMPGX_E_INVALID, "a CIGAR run carries an op code beyond IMPGX_OP_M (the packing is op << 29 | length, ops = X I D M)");
  }
  idx->device_bytes = bytes;
  return idx.release();
}

// ----------------------------------------------------------------- results
}  // namespace impgx

impgx_results::~impgx_results() {
  if (on_device) {
    cudaSetDevice(device);
    cudaStream_t s = (cudaStream_t)stream;
    cudaFreeAsync(d_row_off, s); cudaFreeAsync(d_qid, s); cudaFreeAsync(d_qf, s); cudaFreeAsync(d_ql, s);
    cudaFreeAsync(d_tid, s); cudaFreeAsync(d_tf, s); cudaFreeAsync(d_tl, s);
  }
}

namespace impgx {

// Visited set of all rows: sorted by (row << 32 | seq, start)
struct Visited {
  DBuf<uint64_t> keys;
  DBuf<int32_t> start, end;
  uint64_t n = 0;
};

struct LevelHits {
  DBuf<Hit> hits;          // ok hits in the reference's order (frontier index, visit rank)
  DBuf<uint32_t> entry;    // entry index per hit (CIGAR emission), only with store_cigar
  DBuf<CigarSlice> slices;
  uint64_t n = 0;
  bool seeds = false;      // the self-interval pieces of a masked query: never filtered by min_output_length
};

// Collector of the direct BED path: boxes of every result of the batch.
struct BedSink {
  // the raw hits of the last hop stay as the liftover wrote them (BoxSrc, merge_kernels.cuh)
  DBuf<Hit> raw_hits;
  DBuf<LiftTask> raw_tasks;
  DBuf<uint32_t> raw_orig;
  uint64_t n_raw = 0;
  bool raw_has_orig = false;
  const uint32_t *raw_gmap = nullptr;  // sharded index: local -> global frontier index of the raw hop
  DBuf<BoxD> boxes;
  DBuf<unsigned long long> counters;  // [0] valid boxes, [1] stage-A roots
  uint64_t prefix = 0;                // slots reserved for the seeds and the earlier (sorted) levels
  uint64_t n = 0;                     // slots in use
  uint32_t level = 0;                 // ord level of the raw (last) hop
  bool filled = false;
};

// (row, query sequence) buckets of the direct BED path (bucket_kernels.cuh): dense tables over
// rows x n_seqs, the boxes of bucket b at boxes[beg[b] .. cur[b]).
struct Buckets {
  uint64_t NB = 0;      // rows x n_seqs
  DBuf<uint32_t> cnt;   // NB + 1 (the last stays 0): boxes counted per bucket = its capacity
  DBuf<uint32_t> beg;   // NB + 1: exclusive scan of cnt, beg[NB] = slots in all
  DBuf<uint32_t> cur;   // NB: fill cursor, starts at beg
  DBuf<BoxRec> boxes;
  uint64_t total = 0;
  uint64_t expect = 0;  // host-side upper bound of the boxes counted so far (the tables are u32)
  uint32_t level = 0;   // ord level of the hop that fills the buckets directly
  const uint32_t *gmap = nullptr;  // sharded index: local -> global frontier index of that hop
  bool laid_out = false;
};

struct BatchOut {
  // device columns of the assembled results
  DBuf<uint64_t> row_off;
  DBuf<uint32_t> q_id, t_id;
  DBuf<int32_t> q_first, q_last, t_first, t_last;
  DBuf<uint64_t> cig_off;
  DBuf<uint32_t> cig;
  uint64_t n_results = 0, n_cig = 0;
};

// Raw output of stab + liftover of one frontier, in processing order.
struct Lifted {
  DBuf<Frontier> fr_loc;   // frontier in processing (target position) order, if permuted
  DBuf<uint32_t> orig;     // processing index -> frontier index
  DBuf<Window> win;
  DBuf<uint32_t> counts;
  DBuf<uint64_t> offs;
  DBuf<LiftTask> tasks;
  DBuf<Hit> hits;
  DBuf<CigarSlice> slices;
  const Frontier *fr = nullptr;      // the frontier the tasks index
  const uint32_t *d_orig = nullptr;  // orig.get() or nullptr (identity)
  uint64_t H = 0, n_ok = 0;
};

class Runner {
 public:
  Runner(impgx_index *idx, const impgx_params &p, cudaStream_t s, Arena &arena, Comm *comm = nullptr)
      : idx_(idx), p_(p), s_(s), ix_(idx->view()), ar_(arena), comm_(comm) {
    sc_.a = &ar_;
  }

  Ctx ctx;

  // One batch of rows whose ranges are on the device. Produces raw results in
  // reference order (bed == false) or BED-merged rows (bed == true).
  void run(const impgx_range *d_ranges, uint32_t n_rows, bool bed, BatchOut &out);
  // Same batch on a target-sharded index (collective over comm): BED rows of
  // the (row, sequence) groups whose sequence this rank owns.
  void run_sharded(const impgx_range *d_ranges, uint32_t n_rows, BatchOut &out);

 private:
  impgx_index *idx_;
  impgx_params p_;
  cudaStream_t s_;
  DevIndexView ix_;
  Arena &ar_;
  Scratch sc_;
  DBuf<uint8_t> d_subset_;
  DBuf<uint32_t> d_row_target_;
  DBuf<unsigned long long> d_counters_;
  cudaEvent_t ev_[2] = {nullptr, nullptr};
  Comm *comm_ = nullptr;  // sharded index only
  // MultiImpg modes: 0 = Impg order (frontier index, coitrees visit rank); 1 / 2 = the 5-key order of
  // src/multi_impg.rs:582-592 with the transitive / query drop rule (k_multi_drop)
  int multi_order_ = 0;
  // masked_regions (CSR over all sequences) of the transitive modes, resident for the batch
  DBuf<uint64_t> d_mask_off_;
  DBuf<int2> d_mask_rng_;
  bool masked_ = false;

  void prepare(const impgx_range *d_ranges, uint32_t n_rows);
  void lift_core(const DBuf<Frontier> &fr, uint64_t nF, bool closed, bool clip,
                 const std::function<void(uint64_t)> &alloc_outputs, Lifted &L, Buckets *bk = nullptr,
                 const std::function<void()> &bk_after_layout = nullptr);
  bool bucket_mode(uint32_t n_rows) const;
  void bk_begin(Buckets &bk, uint32_t n_rows);
  void bk_layout(Buckets &bk);
  void bk_add_boxd(Buckets &bk, const BoxD *boxes, uint64_t n, bool scatter);
  void merge_buckets(Buckets &bk, uint32_t n_rows, BatchOut &out);
  float run_bucket_kernels(Buckets &bk, uint32_t *out_cnt, bool reduce);
  void reduce_and_route(Buckets &bk, uint32_t n_rows, DBuf<BoxD> &recv, uint64_t &n_recv);
  void merge_oversized(Buckets &bk, const uint32_t *list, uint32_t n_over, uint32_t *out_cnt);
  void prefix_boxes(BedSink &sink, const impgx_range *d_ranges, uint32_t n_rows, std::vector<LevelHits> &levels,
                    bool query_mode);
  int bits_a_ = 0, bits_b_ = 0;  // widths of the packed merge keys of this batch
  OutCols alloc_out_cols(BatchOut &out, uint64_t n);
  void stage_a(const BoxD *boxes, uint64_t nB, uint64_t nv, DBuf<BoxD> &acc, DBuf<uint64_t> &is_root);
  void stage_b(const BoxD *acc, const uint64_t *is_root, uint64_t n, unsigned long long *d_root_counter,
               uint32_t n_rows, BatchOut &out, uint32_t *row_cnt, Buckets *to_buckets = nullptr,
               uint32_t *bucket_out_cnt = nullptr);
  bool merge_fused(const BoxSrc &src, uint64_t nB, BatchOut &out, uint32_t *row_cnt);
  void materialize_boxes(BedSink &sink);
  void route_boxes(const BoxSrc &src, uint64_t nB, DBuf<BoxD> &recv, uint64_t &n_recv);
  void route_hits(const Lifted &L, const uint32_t *gmap, LevelHits &lvl);
  void global_frontier(DBuf<Frontier> &fr, uint64_t &nF, DBuf<uint32_t> &gmap, uint64_t total,
                       const std::vector<uint64_t> &cnt);
  void stab_and_lift(const DBuf<Frontier> &fr, uint64_t nF, bool closed, bool clip, BedSink *sink, LevelHits &lvl,
                     Buckets *bk = nullptr, const std::function<void()> &bk_after_layout = nullptr);
  void bed_merge_direct(BedSink &sink, uint32_t n_rows, BatchOut &out);
  void fold(const LevelHits &lvl, uint32_t n_rows, Visited &V, DBuf<Frontier> &next, uint64_t &n_next,
            bool raw_pieces = false);
  void run_dfs(const impgx_range *d_ranges, uint32_t n_rows, DBuf<Frontier> &fr, uint64_t nF, Visited &V,
               std::vector<LevelHits> &levels);
  void dfs_pop(const DBuf<DfsEntry> &stack, uint64_t n_stack, uint32_t n_rows, DBuf<uint64_t> &popped,
               DBuf<uint32_t> &cur_depth, DBuf<Frontier> &fr, uint64_t &nF);
  bool dfs_restack(DBuf<DfsEntry> &stack, uint64_t &n_stack, const DBuf<uint64_t> &popped, const Frontier *pieces,
                   uint64_t n_pieces, const uint32_t *cur_depth, uint32_t n_rows);
  void sharded_dfs(DBuf<Frontier> &fr0, uint64_t nF0, uint32_t n_rows, Visited &V, uint32_t ord_level,
                   std::vector<DBuf<BoxD>> &level_boxes, std::vector<uint64_t> &level_n, uint64_t &prior,
                   unsigned long long *box_counter);
  void assemble(const impgx_range *d_ranges, uint32_t n_rows, std::vector<LevelHits> &levels, bool query_mode,
                BatchOut &out);
  void bed_merge(BatchOut &raw, uint32_t n_rows, BatchOut &out);
  float timed_begin();
};

// stab (count, scan, fill) + liftover of one frontier: raw hits in processing
// order. `alloc_outputs(H)` runs as soon as the hit count is known, before the
// temporaries, so that what the caller keeps sits below them in the arena.
void Runner::lift_core(const DBuf<Frontier> &fr_ref, uint64_t nF, bool closed, bool clip,
                       const std::function<void(uint64_t)> &alloc_outputs, Lifted &L, Buckets *bk,
                       const std::function<void()> &bk_after_layout) {
  L.H = L.n_ok = 0;
  if (nF == 0) return;
  std::unique_ptr<WallTimer> wt(new WallTimer(ctx.w_stab));
  cudaEvent_t e0, e1, e2;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventCreate(&e2));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  // Process the frontier in (target sequence, start) order rather than row
  // order: neighbouring threads then stab the same entry windows and lift
  // through the same alignments, so entry records, checkpoints and run blocks
  // are reused from L1/L2 instead of being re-fetched from HBM. `orig` maps
  // back to the reference's frontier index, which the order keys are built on.
  const bool locality = nF >= 4096 && !getenv("IMPGX_NO_LOCALITY");
  if (locality) {
    DBuf<uint64_t> lk(nF, ar_);
    L.orig.alloc(nF, ar_);
    LAUNCH(k_locality_keys, grid_threads(nF), 256, s_, fr_ref.get(), nF, lk.get(), L.orig.get());
    sort_pairs(lk, L.orig, nF, 0, 32 + bits_for(ix_.n_seqs), sc_, s_, ctx);
    L.fr_loc.alloc(nF, ar_);
    LAUNCH(k_gather<Frontier>, grid_threads(nF), 256, s_, fr_ref.get(), L.orig.get(), nF, L.fr_loc.get());
  }
  L.fr = locality ? L.fr_loc.get() : fr_ref.get();
  L.d_orig = locality ? L.orig.get() : nullptr;
  L.win.alloc(nF, ar_);
  L.counts.alloc(nF, ar_);
  L.offs.alloc(nF + 1, ar_);
  uint32_t *bcnt = bk ? bk->cnt.get() : nullptr;
  {
    auto kern = closed ? (bk ? k_stab_count<true, true> : k_stab_count<true, false>)
                       : (bk ? k_stab_count<false, true> : k_stab_count<false, false>);
    LAUNCH(kern, grid_warps(nF), 256, s_, ix_, L.fr, nF, L.win.get(), L.counts.get(), bcnt);
  }
  CUDA_CHECK(cudaMemsetAsync(L.offs.get() + nF, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(nF), 256, s_, L.counts.get(), nF, L.offs.get());
  exclusive_scan_u64(L.offs.get(), nF + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t H = read_u64(L.offs.get() + nF, s_, ctx);
  REQUIRE(H < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 hits in one hop of one batch; lower IMPGX_ROWS_PER_BATCH");
  ctx.stab_ranges += nF;
  ctx.liftovers += H;
  L.H = H;
  if (bk) {
    // every hit is counted: bucket offsets, the box array, and the boxes that already exist
    bk->expect += H;
    bk_layout(*bk);
    if (bk_after_layout) bk_after_layout();
  }
  if (H == 0) {
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return;
  }
  // outputs that survive the hop go below the temporaries (stack-discipline arena)
  if (alloc_outputs) alloc_outputs(H);
  L.tasks.alloc(H, ar_);
  if (closed) LAUNCH(k_stab_fill<true>, grid_warps(nF), 256, s_, ix_, L.fr, nF, L.win.get(), L.offs.get(), L.tasks.get());
  else LAUNCH(k_stab_fill<false>, grid_warps(nF), 256, s_, ix_, L.fr, nF, L.win.get(), L.offs.get(), L.tasks.get());
  CUDA_CHECK(cudaEventRecord(e1, s_));
  wt.reset(new WallTimer(ctx.w_lift));

  if (!bk) L.hits.alloc(H, ar_);
  if (p_.store_cigar) L.slices.alloc(H, ar_);
  CUDA_CHECK(cudaMemsetAsync(d_counters_.get(), 0, 32, s_));
  LiftParams lp;
  lp.clip = clip ? 1 : 0;
  lp.min_output_len = p_.min_output_length;
  lp.use_identity = std::isnan(p_.min_identity) ? 0 : 1;
  lp.min_identity = p_.min_identity;
  // Impg::query applies the subset filter afterwards with the same rule
  // (src/subset_filter.rs:84-100), so it is folded into the kernel in every mode
  lp.subset = p_.subset_mask ? d_subset_.get() : nullptr;
  lp.row_target = d_row_target_.get();
  // endpoint kernel when neither the clipped CIGAR nor the identity is needed
  const bool ends = !p_.store_cigar && !lp.use_identity && !getenv("IMPGX_FULL_SCAN");
  REQUIRE(ends || !bk, IMPGX_E_CUDA, "internal: the bucket path needs the endpoint liftover");
  // A/B switches: the variant with overlapped gathers; 3 or 4 resident CTAs per SM (80 or 64 registers per thread)
  static const bool pipe = env_u64("IMPGX_LIFT_OVL", 0) != 0;
  static const bool minb4 = env_u64("IMPGX_LIFT_MINB", 3) == 4;
  if (ends) {
    auto kern = bk ? (pipe ? (minb4 ? k_liftover_ends<true, true, 4> : k_liftover_ends<true, true, 3>)
                           : (minb4 ? k_liftover_ends<true, false, 4> : k_liftover_ends<true, false, 3>))
                   : (pipe ? (minb4 ? k_liftover_ends<false, true, 4> : k_liftover_ends<false, true, 3>)
                           : (minb4 ? k_liftover_ends<false, false, 4> : k_liftover_ends<false, false, 3>));
    const BucketOut bo = bk ? BucketOut{bk->cur.get(), bk->boxes.get(), L.d_orig, bk->gmap, bk->level} : BucketOut{};
    LAUNCH(kern, grid_threads(H, 256, 8), 256, s_, ix_, L.fr, L.tasks.get(), H, lp, bk ? (Hit *)nullptr : L.hits.get(),
           d_counters_.get(), bo);
  } else
    LAUNCH(k_liftover, grid_warps(H, 256, 8), 256, s_, ix_, L.fr, L.tasks.get(), H, lp, L.hits.get(), L.slices.get(),
           d_counters_.get());
  CUDA_CHECK(cudaEventRecord(e2, s_));
  unsigned long long cnt[4];
  {
    uint64_t *h = readback_slot();
    CUDA_CHECK(cudaMemcpyAsync(h, d_counters_.get(), 32, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    memcpy(cnt, h, 32);
  }
  ctx.d2h_bytes += 32;
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx.stab_ms += ms;
  cudaEventElapsedTime(&ms, e1, e2);
  ctx.lift_ms += ms;
  ctx.lift_launches++;
  {
    static const bool trace_launches = env_u64("IMPGX_TRACE", 0) >= 2;  // one line per liftover launch (ncu captures are matched by it)
    if (trace_launches)
      fprintf(stderr, "[impgx] liftover launch: ranges %llu hits %llu ok %llu %.3f ms%s\n", (unsigned long long)nF,
              (unsigned long long)H, cnt[1], ms, bk ? " (into buckets)" : "");
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  ctx.lift_runs += cnt[0];
  // bytes the kernel's own algorithm touches: task 8 + range 16 + entry 32 + hit 32 (+ slice 16),
  // the checkpoints probed and the run blocks read
  if (ends) ctx.lift_touched += H * (8 + 16 + 32 + 32) + cnt[2] * 8 + cnt[0] * 4;
  else ctx.lift_touched += H * (8 + 16 + 32 + 16 + 32 + (p_.store_cigar ? 16 : 0)) + cnt[0] * 4;
  // algorithmic bytes per SURVEY.md §8(d): 32 (entry) + 16 (two checkpoints) + 4 r_ov (runs
  // intersecting the request) + 24 (hit out) [+ 4 r_ov with store_cigar]
  ctx.lift_bytes += H * (32 + 16 + 24) + cnt[3] * 4 * (p_.store_cigar ? 2 : 1);
  ctx.lift_rov += cnt[3];
  L.n_ok = cnt[1];
}

// One hop on a whole (unsharded) index: leaves the accepted hits ordered by
// (frontier index, visit rank) in lvl, or — last hop of the direct BED path —
// writes their boxes straight into the sink.
void Runner::stab_and_lift(const DBuf<Frontier> &fr_ref, uint64_t nF, bool closed, bool clip, BedSink *sink,
                           LevelHits &lvl, Buckets *bk, const std::function<void()> &bk_after_layout) {
  lvl.n = 0;
  if (nF == 0) return;
  Lifted L;
  lift_core(fr_ref, nF, closed, clip,
            [&](uint64_t H) {
              if (!sink) lvl.hits.alloc(H, ar_);
              if (p_.store_cigar) {
                lvl.entry.alloc(H, ar_);
                lvl.slices.alloc(H, ar_);
              }
            },
            L, bk, bk_after_layout);
  if (bk) {  // the accepted hits of this hop sit in their buckets
    if (sink) sink->filled = true;
    return;
  }
  const uint64_t H = L.H;
  uint64_t n_ok = L.n_ok;
  if (H == 0) return;
  if (multi_order_) {
    CUDA_CHECK(cudaMemsetAsync(d_counters_.get(), 0, 8, s_));
    LAUNCH(k_multi_drop, grid_threads(H), 256, s_, L.hits.get(), L.tasks.get(), L.fr, H, multi_order_, d_counters_.get());
    unsigned long long c = 0;
    CUDA_CHECK(cudaMemcpyAsync(&c, d_counters_.get(), 8, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    ctx.d2h_bytes += 8;
    n_ok = c;
  }
  lvl.n = n_ok;
  if (n_ok == 0 && !sink) return;
  WallTimer wt(ctx.w_order);

  if (sink) {
    // direct BED path, last hop: no ordering sort and no copy — the hits stay where the liftover
    // wrote them; the reference order travels as (range, visit rank)
    sink->raw_hits = std::move(L.hits);
    sink->raw_tasks = std::move(L.tasks);
    sink->raw_has_orig = L.d_orig != nullptr;
    if (sink->raw_has_orig) sink->raw_orig = std::move(L.orig);
    sink->n_raw = H;
    sink->filled = true;
    lvl.n = 0;
    return;
  }
  DBuf<uint64_t> keys(H, ar_);
  DBuf<uint32_t> perm(H, ar_);
  if (multi_order_) {
    // LSD over the five keys: four stable 32-bit passes, then (range, query id)
    const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
    DBuf<uint32_t> k32(H, ar_);
    LAUNCH(k_iota_u32, grid_threads(H), 256, s_, perm.get(), H);
    for (int field = 0; field < 4; field++) {
      LAUNCH(k_multi_field_keys, grid_threads(H), 256, s_, L.hits.get(), perm.get(), H, field, k32.get());
      sort_pairs(k32, perm, H, 0, 32, sc_, s_, ctx);
    }
    LAUNCH(k_multi_major_keys, grid_threads(H), 256, s_, L.hits.get(), L.tasks.get(), L.d_orig, perm.get(), H,
           (uint32_t)nF, seq_bits, keys.get());
    sort_pairs(keys, perm, H, 0, seq_bits + bits_for(nF), sc_, s_, ctx);
  } else {
    LAUNCH(k_hit_order_keys, grid_threads(H), 256, s_, L.hits.get(), L.tasks.get(), L.d_orig, H, (uint32_t)nF,
           keys.get(), perm.get());
    sort_pairs(keys, perm, H, 0, 32 + bits_for(nF), sc_, s_, ctx);
  }
  LAUNCH(k_gather<Hit>, grid_threads(n_ok), 256, s_, L.hits.get(), perm.get(), n_ok, lvl.hits.get());
  if (p_.store_cigar) {
    LAUNCH(k_gather_entry, grid_threads(n_ok), 256, s_, L.tasks.get(), perm.get(), n_ok, lvl.entry.get());
    LAUNCH(k_gather<CigarSlice>, grid_threads(n_ok), 256, s_, L.slices.get(), perm.get(), n_ok, lvl.slices.get());
  }
}

// The sequential fold + next frontier of one level (src/impg.rs:2467-2593).
void Runner::fold(const LevelHits &lvl, uint32_t n_rows, Visited &V, DBuf<Frontier> &next, uint64_t &n_next,
                  bool raw_pieces) {
  n_next = 0;
  const uint64_t n = lvl.n;
  if (n == 0) return;
  WallTimer wt(ctx.w_fold);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  // group by (row, q_id), stable w.r.t. the reference's hit order
  DBuf<uint64_t> keys(n, ar_);
  DBuf<uint32_t> perm(n, ar_);
  LAUNCH(k_fold_keys, grid_threads(n), 256, s_, lvl.hits.get(), n, n_rows, keys.get(), perm.get());
  sort_pairs(keys, perm, n, 0, 32 + bits_for(n_rows), sc_, s_, ctx);
  DBuf<Hit> sorted(n, ar_);
  LAUNCH(k_gather<Hit>, grid_threads(n), 256, s_, lvl.hits.get(), perm.get(), n, sorted.get());

  DBuf<uint64_t> head(n + 1, ar_), head_scan(n + 1, ar_), n_inc(1, ar_);
  CUDA_CHECK(cudaMemcpyAsync(n_inc.get(), &n, 8, cudaMemcpyHostToDevice, s_));
  CUDA_CHECK(cudaMemsetAsync(head.get() + n, 0, 8, s_));
  const uint64_t limit = (uint64_t)n_rows << 32;
  LAUNCH(k_group_heads, grid_threads(n), 256, s_, keys.get(), n, limit, head.get(), n_inc.get());
  CUDA_CHECK(cudaMemcpyAsync(head_scan.get(), head.get(), (n + 1) * 8, cudaMemcpyDeviceToDevice, s_));
  exclusive_scan_u64(head_scan.get(), n + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t G = read_u64(head_scan.get() + n, s_, ctx);
  if (G == 0) {
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return;
  }
  DBuf<FoldGroup> groups(G, ar_);
  DBuf<uint64_t> list_off(G + 1, ar_), piece_off(G + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(list_off.get() + G, 0, 8, s_));
  CUDA_CHECK(cudaMemsetAsync(piece_off.get() + G, 0, 8, s_));
  LAUNCH(k_make_groups, grid_threads(n), 256, s_, keys.get(), head.get(), head_scan.get(), n, n_inc.get(),
         V.keys.get(), V.n, groups.get(), list_off.get(), piece_off.get(),
         masked_ ? (const uint64_t *)d_mask_off_.get() : nullptr);
  exclusive_scan_u64(list_off.get(), G + 1, sc_, s_);
  exclusive_scan_u64(piece_off.get(), G + 1, sc_, s_);
  ctx.launches += 4;
  uint64_t list_total = 0, piece_total = 0;
  read_u64x2(list_off.get() + G, piece_off.get() + G, list_total, piece_total, s_, ctx);
  LAUNCH(k_set_group_offsets, grid_threads(G), 256, s_, groups.get(), G, list_off.get(), piece_off.get());

  DBuf<int2> lists(list_total, ar_);
  DBuf<Frontier> pieces(piece_total, ar_);
  DBuf<uint32_t> list_len(G, ar_), piece_cnt(G, ar_);
  LAUNCH(k_fold, grid_threads(G, 128, 16), 128, s_, groups.get(), G, sorted.get(), V.start.get(), V.end.get(),
         ix_.seq_len, p_.min_distance_between_ranges, p_.min_transitive_len, lists.get(), list_len.get(), pieces.get(),
         piece_cnt.get(), masked_ ? (const uint64_t *)d_mask_off_.get() : nullptr,
         masked_ ? (const int2 *)d_mask_rng_.get() : nullptr);

  // ---- new visited set = untouched old entries + the groups' lists, re-sorted
  DBuf<uint64_t> lo(G + 1, ar_), po(G + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(lo.get() + G, 0, 8, s_));
  CUDA_CHECK(cudaMemsetAsync(po.get() + G, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(G), 256, s_, list_len.get(), G, lo.get());
  LAUNCH(k_u32_to_u64, grid_threads(G), 256, s_, piece_cnt.get(), G, po.get());
  exclusive_scan_u64(lo.get(), G + 1, sc_, s_);
  exclusive_scan_u64(po.get(), G + 1, sc_, s_);
  ctx.launches += 4;
  DBuf<uint64_t> keep(V.n + 1, ar_), keep_scan(V.n + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(keep.get() + V.n, 0, 8, s_));
  if (V.n) LAUNCH(k_visited_keep_flags, grid_threads(V.n), 256, s_, V.keys.get(), V.n, groups.get(), G, keep.get());
  CUDA_CHECK(cudaMemcpyAsync(keep_scan.get(), keep.get(), (V.n + 1) * 8, cudaMemcpyDeviceToDevice, s_));
  exclusive_scan_u64(keep_scan.get(), V.n + 1, sc_, s_);
  ctx.launches += 2;
  // the three sizes of this stage in one synchronisation
  uint64_t new_lists = 0, n_pieces = 0, kept = 0;
  {
    uint64_t *h = readback_slot();
    CUDA_CHECK(cudaMemcpyAsync(h, lo.get() + G, 8, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaMemcpyAsync(h + 1, po.get() + G, 8, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaMemcpyAsync(h + 2, keep_scan.get() + V.n, 8, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    ctx.d2h_bytes += 24;
    new_lists = h[0];
    n_pieces = h[1];
    kept = h[2];
  }
  const uint64_t vn = kept + new_lists;
  Visited nv;
  nv.keys.alloc(vn, ar_);
  nv.start.alloc(vn, ar_);
  nv.end.alloc(vn, ar_);
  nv.n = vn;
  if (V.n)
    LAUNCH(k_visited_copy_kept, grid_threads(V.n), 256, s_, V.keys.get(), V.start.get(), V.end.get(), V.n, keep.get(),
           keep_scan.get(), nv.keys.get(), nv.start.get(), nv.end.get());
  LAUNCH(k_compact_lists, grid_threads(G), 256, s_, groups.get(), G, lists.get(), list_len.get(), lo.get(), kept,
         nv.keys.get(), nv.start.get(), nv.end.get());
  if (vn) {
    // sort by (key, start): LSD = start first, then key (stable)
    DBuf<uint32_t> sk(vn, ar_), sp(vn, ar_);
    LAUNCH(k_iota_u32, grid_threads(vn), 256, s_, sp.get(), vn);
    LAUNCH(k_start_keys_i32, grid_threads(vn), 256, s_, nv.start.get(), vn, sk.get());
    sort_pairs(sk, sp, vn, 0, 32, sc_, s_, ctx);
    DBuf<uint64_t> k2(vn, ar_);
    LAUNCH(k_gather<uint64_t>, grid_threads(vn), 256, s_, nv.keys.get(), sp.get(), vn, k2.get());
    sort_pairs(k2, sp, vn, 0, 32 + bits_for(n_rows), sc_, s_, ctx);
    DBuf<int32_t> s2(vn, ar_), e2(vn, ar_);
    LAUNCH(k_gather<int32_t>, grid_threads(vn), 256, s_, nv.start.get(), sp.get(), vn, s2.get());
    LAUNCH(k_gather<int32_t>, grid_threads(vn), 256, s_, nv.end.get(), sp.get(), vn, e2.get());
    nv.keys = std::move(k2);
    nv.start = std::move(s2);
    nv.end = std::move(e2);
  }
  V = std::move(nv);

  // ---- next frontier: pieces sorted by (row, id, start), touching ones merged
  if (n_pieces && raw_pieces) {
    // DFS: the pieces are pushed on the per-row stacks as they are
    next.alloc(n_pieces, ar_);
    n_next = n_pieces;
    LAUNCH(k_compact_pieces, grid_threads(G), 256, s_, groups.get(), G, pieces.get(), piece_cnt.get(), po.get(),
           next.get());
  } else if (n_pieces) {
    DBuf<Frontier> pc(n_pieces, ar_);
    LAUNCH(k_compact_pieces, grid_threads(G), 256, s_, groups.get(), G, pieces.get(), piece_cnt.get(), po.get(),
           pc.get());
    DBuf<uint32_t> sk(n_pieces, ar_), sp(n_pieces, ar_);
    LAUNCH(k_piece_start_keys, grid_threads(n_pieces), 256, s_, pc.get(), n_pieces, sk.get(), sp.get());
    sort_pairs(sk, sp, n_pieces, 0, 32, sc_, s_, ctx);
    DBuf<uint64_t> k2(n_pieces, ar_);
    LAUNCH(k_piece_seq_keys, grid_threads(n_pieces), 256, s_, pc.get(), sp.get(), n_pieces, k2.get());
    sort_pairs(k2, sp, n_pieces, 0, 32 + bits_for(n_rows), sc_, s_, ctx);
    DBuf<Frontier> ps(n_pieces, ar_);
    LAUNCH(k_gather<Frontier>, grid_threads(n_pieces), 256, s_, pc.get(), sp.get(), n_pieces, ps.get());
    DBuf<uint64_t> fh(n_pieces + 1, ar_), fs(n_pieces + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(fh.get() + n_pieces, 0, 8, s_));
    LAUNCH(k_frontier_heads, grid_threads(n_pieces), 256, s_, ps.get(), n_pieces, fh.get());
    CUDA_CHECK(cudaMemcpyAsync(fs.get(), fh.get(), (n_pieces + 1) * 8, cudaMemcpyDeviceToDevice, s_));
    exclusive_scan_u64(fs.get(), n_pieces + 1, sc_, s_);
    ctx.launches += 2;
    n_next = read_u64(fs.get() + n_pieces, s_, ctx);
    next.alloc(n_next, ar_);
    LAUNCH(k_frontier_merge, grid_threads(n_pieces), 256, s_, ps.get(), n_pieces, fh.get(), fs.get(), next.get());
  }
  CUDA_CHECK(cudaEventRecord(e1, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx.fold_ms += ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

// Per-row concatenation: self interval, then each level's hits in order.
void Runner::assemble(const impgx_range *d_ranges, uint32_t n_rows, std::vector<LevelHits> &levels, bool query_mode,
                      BatchOut &out) {
  WallTimer wt(ctx.w_assemble);
  const int32_t min_out = p_.min_output_length;
  DBuf<uint32_t> seed_cnt(n_rows, ar_), total_cnt(n_rows, ar_);
  // Impg::query's caller filters every result incl. the self interval by
  // min_output_length (src/main.rs:11686-11691); BFS never filters the seed.
  LAUNCH(k_seed_counts, grid_threads(n_rows), 256, s_, d_ranges, n_rows, min_out, query_mode ? 1 : 0, seed_cnt.get());
  // masked query: the self intervals are the unmasked pieces, carried as a level of their own
  if (masked_) CUDA_CHECK(cudaMemsetAsync(seed_cnt.get(), 0, (size_t)n_rows * 4, s_));
  CUDA_CHECK(cudaMemcpyAsync(total_cnt.get(), seed_cnt.get(), n_rows * 4, cudaMemcpyDeviceToDevice, s_));

  struct PerLevel {
    DBuf<uint64_t> pass, pass_scan, row_start;
    DBuf<uint32_t> row_cnt;
  };
  std::vector<PerLevel> pl(levels.size());
  for (size_t l = 0; l < levels.size(); l++) {
    const uint64_t n = levels[l].n;
    pl[l].row_cnt.alloc(n_rows, ar_);
    CUDA_CHECK(cudaMemsetAsync(pl[l].row_cnt.get(), 0, n_rows * 4, s_));
    if (n == 0) continue;
    pl[l].pass.alloc(n + 1, ar_);
    pl[l].pass_scan.alloc(n + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(pl[l].pass.get() + n, 0, 8, s_));
    LAUNCH(k_level_pass, grid_threads(n), 256, s_, levels[l].hits.get(), n, levels[l].seeds ? -1 : min_out,
           pl[l].pass.get(), pl[l].row_cnt.get());
    CUDA_CHECK(cudaMemcpyAsync(pl[l].pass_scan.get(), pl[l].pass.get(), (n + 1) * 8, cudaMemcpyDeviceToDevice, s_));
    exclusive_scan_u64(pl[l].pass_scan.get(), n + 1, sc_, s_);
    pl[l].row_start.alloc(n_rows + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(pl[l].row_start.get() + n_rows, 0, 8, s_));
    LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, pl[l].row_cnt.get(), n_rows, pl[l].row_start.get());
    exclusive_scan_u64(pl[l].row_start.get(), n_rows + 1, sc_, s_);
    LAUNCH(k_add_u32, grid_threads(n_rows), 256, s_, total_cnt.get(), pl[l].row_cnt.get(), n_rows);
    ctx.launches += 4;
  }
  out.row_off.alloc(n_rows + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(out.row_off.get() + n_rows, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, total_cnt.get(), n_rows, out.row_off.get());
  exclusive_scan_u64(out.row_off.get(), n_rows + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t R = read_u64(out.row_off.get() + n_rows, s_, ctx);
  out.n_results = R;
  out.q_id.alloc(R, ar_); out.t_id.alloc(R, ar_);
  out.q_first.alloc(R, ar_); out.q_last.alloc(R, ar_);
  out.t_first.alloc(R, ar_); out.t_last.alloc(R, ar_);
  OutCols oc;
  oc.q_id = out.q_id.get(); oc.q_first = out.q_first.get(); oc.q_last = out.q_last.get();
  oc.t_id = out.t_id.get(); oc.t_first = out.t_first.get(); oc.t_last = out.t_last.get();
  DBuf<uint32_t> src_entry;
  DBuf<CigarSlice> src_slice;
  oc.cig_len = nullptr; oc.src_entry = nullptr; oc.src_slice = nullptr;
  if (p_.store_cigar) {
    out.cig_off.alloc(R + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(out.cig_off.get(), 0, (R + 1) * 8, s_));
    src_entry.alloc(R, ar_);
    src_slice.alloc(R, ar_);
    oc.cig_len = out.cig_off.get();
    oc.src_entry = src_entry.get();
    oc.src_slice = src_slice.get();
  }
  LAUNCH(k_scatter_seed, grid_threads(n_rows), 256, s_, d_ranges, n_rows, seed_cnt.get(), out.row_off.get(), oc);
  DBuf<uint32_t> base(n_rows, ar_);
  CUDA_CHECK(cudaMemcpyAsync(base.get(), seed_cnt.get(), n_rows * 4, cudaMemcpyDeviceToDevice, s_));
  for (size_t l = 0; l < levels.size(); l++) {
    const uint64_t n = levels[l].n;
    if (n) {
      LAUNCH(k_scatter_level, grid_threads(n), 256, s_, levels[l].hits.get(), levels[l].entry.get(),
             levels[l].slices.get(), n, pl[l].pass.get(), pl[l].pass_scan.get(), pl[l].row_start.get(),
             out.row_off.get(), base.get(), oc);
      LAUNCH(k_add_u32, grid_threads(n_rows), 256, s_, base.get(), pl[l].row_cnt.get(), n_rows);
    }
  }
  if (p_.store_cigar) {
    exclusive_scan_u64(out.cig_off.get(), R + 1, sc_, s_);
    ctx.launches += 2;
    out.n_cig = read_u64(out.cig_off.get() + R, s_, ctx);
    out.cig.alloc(std::max<uint64_t>(out.n_cig, 1), ar_);
    if (R) LAUNCH(k_emit_cigar_results, grid_warps(R), 256, s_, ix_, oc, out.cig_off.get(), R, out.cig.get());
    CUDA_CHECK(cudaStreamSynchronize(s_));  // src_* are released on scope exit
  }
}

// seeds, then the levels that were ordered for the fold (their index is their ordinal), as BoxD records
void Runner::prefix_boxes(BedSink &sink, const impgx_range *d_ranges, uint32_t n_rows, std::vector<LevelHits> &levels,
                          bool query_mode) {
  uint64_t held = 0;  // hits held by the ordered levels (BFS hops or DFS rounds)
  for (auto &l : levels) held += l.n;
  sink.prefix = (uint64_t)n_rows + held;
  sink.boxes.alloc(sink.prefix, ar_);
  if (masked_)  // the self intervals are the seed level below; the per-row seed slots stay invalid
    CUDA_CHECK(cudaMemsetAsync(sink.boxes.get(), 0, (size_t)n_rows * sizeof(BoxD), s_));
  else
    LAUNCH(k_boxes_from_seeds, grid_threads(n_rows), 256, s_, d_ranges, n_rows, p_.min_output_length,
           query_mode ? 1 : 0, sink.boxes.get(), sink.counters.get(), (const uint32_t *)nullptr, 0u);
  uint64_t off = n_rows;
  for (size_t l = 0; l < levels.size(); l++) {
    if (!levels[l].n) continue;
    LAUNCH(k_boxes_from_sorted_level, grid_threads(levels[l].n), 256, s_, levels[l].hits.get(), levels[l].n,
           (uint32_t)l + 1, levels[l].seeds ? -1 : p_.min_output_length, sink.boxes.get() + off, sink.counters.get());
    off += levels[l].n;
  }
}

// ---- (row, query sequence) buckets of the direct BED path (bucket_kernels.cuh)
bool Runner::bucket_mode(uint32_t n_rows) const {
  if (getenv("IMPGX_MERGE_SORTED") || getenv("IMPGX_MERGE_GLOBAL") || getenv("IMPGX_BED_GENERIC") || getenv("IMPGX_FULL_SCAN"))
    return false;  // test / diagnostic switches of the sort-based paths
  if (p_.store_cigar || !std::isnan(p_.min_identity)) return false;  // the last hop must be the endpoint liftover
  if (p_.merge_distance < 0 && !p_.merge_strands) return false;      // unsorted output: reference order per row
  return (uint64_t)n_rows * ix_.n_seqs <= (1ull << 27);
}

void Runner::bk_begin(Buckets &bk, uint32_t n_rows) {
  bk.NB = (uint64_t)n_rows * ix_.n_seqs;
  bk.cnt.alloc(bk.NB + 1, ar_);
  bk.beg.alloc(bk.NB + 1, ar_);
  bk.cur.alloc(bk.NB, ar_);
  CUDA_CHECK(cudaMemsetAsync(bk.cnt.get(), 0, (bk.NB + 1) * 4, s_));
}

// counts -> bucket offsets, cursors and the box array
void Runner::bk_layout(Buckets &bk) {
  REQUIRE(bk.expect < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 results in one batch; lower IMPGX_ROWS_PER_BATCH");
  exclusive_scan_u32(bk.cnt.get(), bk.beg.get(), bk.NB + 1, sc_, s_);
  ctx.launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(bk.cur.get(), bk.beg.get(), bk.NB * 4, cudaMemcpyDeviceToDevice, s_));
  {
    uint64_t *h = readback_slot();
    h[0] = 0;
    CUDA_CHECK(cudaMemcpyAsync(h, bk.beg.get() + bk.NB, 4, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    ctx.d2h_bytes += 4;
    bk.total = (uint32_t)h[0];
  }
  bk.boxes.alloc(bk.total, ar_);
  bk.laid_out = true;
}

// boxes that exist as BoxD records: counted into the buckets (before the layout) / written into them (after it)
void Runner::bk_add_boxd(Buckets &bk, const BoxD *boxes, uint64_t n, bool scatter) {
  if (n == 0) return;
  if (!scatter) bk.expect += n;
  if (scatter) LAUNCH(k_bucket_scatter_boxd, grid_threads(n), 256, s_, boxes, n, ix_.n_seqs, bk.cur.get(), bk.boxes.get());
  else LAUNCH(k_bucket_count_boxd, grid_threads(n), 256, s_, boxes, n, ix_.n_seqs, bk.cnt.get());
}

// The bucket kernels over every non-empty bucket: out_cnt[b] = rows (merge) or surviving boxes (reduce) staged
// over the first slots of bucket b. Returns the device time of the launches.
float Runner::run_bucket_kernels(Buckets &bk, uint32_t *out_cnt, bool reduce) {
  const uint64_t NB = bk.NB;
  const uint64_t cap = std::min<uint64_t>(NB, bk.total);
  DBuf<uint32_t> lists((uint64_t)(SEG_CLASSES + 2) * cap, ar_);
  DBuf<unsigned int> cls(SEG_CLASSES + 2, ar_);
  CUDA_CHECK(cudaMemsetAsync(cls.get(), 0, (SEG_CLASSES + 2) * 4, s_));
  int min_class = (int)env_u64("IMPGX_SEG_MIN_CLASS", 0);  // test hook: run the larger-bucket kernels on small data
  if (min_class < 0 || min_class >= SEG_CLASSES) min_class = 0;
  LAUNCH(k_bucket_classify, grid_threads(NB), 256, s_, bk.beg.get(), bk.cur.get(), NB, lists.get(), cap, cls.get(), min_class);
  unsigned int hc[SEG_CLASSES + 2];
  {
    static_assert(sizeof(hc) <= 32, "class counters must fit the readback slot");
    uint64_t *h = readback_slot();
    CUDA_CHECK(cudaMemcpyAsync(h, cls.get(), sizeof(hc), cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    memcpy(hc, h, sizeof(hc));
  }
  ctx.d2h_bytes += sizeof(hc);
  const int64_t d = p_.merge_distance;
  const int ms = p_.merge_strands ? 1 : 0;
  cudaEvent_t k0, k1;
  CUDA_CHECK(cudaEventCreate(&k0));
  CUDA_CHECK(cudaEventCreate(&k1));
  CUDA_CHECK(cudaEventRecord(k0, s_));
  if (hc[TINY_CLASS])  // one thread per bucket of <= 4 boxes
    LAUNCH(k_merge_tiny, grid_threads(hc[TINY_CLASS], 128, 16), 128, s_, bk.boxes.get(), bk.beg.get(), bk.cur.get(),
           lists.get() + (uint64_t)TINY_CLASS * cap, hc[TINY_CLASS], d, ms, reduce ? 1 : 0, out_cnt,
           (const unsigned int *)nullptr);
  auto launch = [&](auto kern, unsigned threads, size_t smem, unsigned per_sm, unsigned seg_per_cta, int c) {
    if (!hc[c]) return;
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid =
        (unsigned)std::min<uint64_t>(((uint64_t)hc[c] + seg_per_cta - 1) / seg_per_cta, (uint64_t)sm_count() * per_sm);
    kern<<<grid, threads, smem, s_>>>(bk.boxes.get(), bk.beg.get(), bk.cur.get(), lists.get() + (uint64_t)c * cap, hc[c], d,
                                      ms, reduce ? 1 : 0, out_cnt, (const unsigned int *)nullptr);
    CUDA_CHECK(cudaGetLastError());
    ctx.launches++;
  };
  launch(k_merge_buckets<32, seg_cap(0)>, 256, (size_t)8 * seg_cap(0) * BK_BYTES, 4, 8, 0);
  launch(k_merge_buckets<32, seg_cap(1)>, 256, (size_t)8 * seg_cap(1) * BK_BYTES, 2, 8, 1);
  {
    static const uint64_t t2 = env_u64("IMPGX_BK_T2", 128);  // threads per CTA of the <= 512 class (A/B switch)
    if (t2 == 64) launch(k_merge_buckets<64, seg_cap(2)>, 64, (size_t)seg_cap(2) * BK_BYTES, 9, 1, 2);
    else if (t2 == 256) launch(k_merge_buckets<256, seg_cap(2)>, 256, (size_t)seg_cap(2) * BK_BYTES, 8, 1, 2);
    else launch(k_merge_buckets<128, seg_cap(2)>, 128, (size_t)seg_cap(2) * BK_BYTES, 8, 1, 2);
  }
  launch(k_merge_buckets<128, seg_cap(3)>, 128, (size_t)seg_cap(3) * BK_BYTES, 4, 1, 3);
  launch(k_merge_buckets<512, seg_cap(4)>, 512, (size_t)seg_cap(4) * BK_BYTES, 1, 1, 4);
  CUDA_CHECK(cudaEventRecord(k1, s_));
  if (hc[SEG_CLASSES]) {
    const uint32_t *over = lists.get() + (uint64_t)SEG_CLASSES * cap;
    if (reduce)  // every box of such a bucket travels as it is (raw: stage A runs where it arrives)
      LAUNCH(k_oversized_keep, grid_threads(hc[SEG_CLASSES]), 256, s_, over, hc[SEG_CLASSES], bk.beg.get(), bk.cur.get(), out_cnt);
    else
      merge_oversized(bk, over, hc[SEG_CLASSES], out_cnt);
  }
  CUDA_CHECK(cudaStreamSynchronize(s_));  // the lists are released on return
  float kms = 0;
  cudaEventElapsedTime(&kms, k0, k1);
  cudaEventDestroy(k0);
  cudaEventDestroy(k1);
  return kms;
}

// Both BED merges, one (row, q) bucket per warp / CTA; the rows come out in (row, q, start) order.
void Runner::merge_buckets(Buckets &bk, uint32_t n_rows, BatchOut &out) {
  WallTimer wt(ctx.w_merge);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  out.row_off.alloc((uint64_t)n_rows + 1, ar_);
  out.n_results = 0;
  ctx.merge_boxes += bk.total;
  if (bk.total == 0) {
    CUDA_CHECK(cudaMemsetAsync(out.row_off.get(), 0, ((size_t)n_rows + 1) * 8, s_));
  } else {
    const uint64_t NB = bk.NB;
    DBuf<uint32_t> out_cnt(NB + 1, ar_), out_off(NB + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(out_cnt.get(), 0, (NB + 1) * 4, s_));
    ctx.merge_kernel_ms += run_bucket_kernels(bk, out_cnt.get(), /*reduce=*/false);
    exclusive_scan_u32(out_cnt.get(), out_off.get(), NB + 1, sc_, s_);
    ctx.launches += 2;
    uint64_t M = 0;
    {
      uint64_t *h = readback_slot();
      h[0] = 0;
      CUDA_CHECK(cudaMemcpyAsync(h, out_off.get() + NB, 4, cudaMemcpyDeviceToHost, s_));
      CUDA_CHECK(cudaStreamSynchronize(s_));
      ctx.d2h_bytes += 4;
      M = (uint32_t)h[0];
    }
    OutCols oc = alloc_out_cols(out, M);
    LAUNCH(k_bucket_compact, grid_threads(NB), 256, s_, bk.boxes.get(), bk.beg.get(), out_cnt.get(), out_off.get(), NB,
           ix_.n_seqs, oc);
    LAUNCH(k_bucket_row_offsets, grid_threads((uint64_t)n_rows + 1), 256, s_, out_off.get(), n_rows, ix_.n_seqs,
           out.row_off.get());
  }
  CUDA_CHECK(cudaEventRecord(e1, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));  // the tables are released on return
  float msf = 0;
  cudaEventElapsedTime(&msf, e0, e1);
  ctx.merge_ms += msf;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

// Buckets beyond SEG_MAX boxes — and only those — through the global two-sort merge; their rows land in
// the same staging slots as the rows of the on-chip merge.
void Runner::merge_oversized(Buckets &bk, const uint32_t *list, uint32_t n_over, uint32_t *out_cnt) {
  DBuf<uint64_t> offs((uint64_t)n_over + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(offs.get() + n_over, 0, 8, s_));
  LAUNCH(k_oversized_sizes, grid_threads(n_over), 256, s_, list, n_over, bk.beg.get(), bk.cur.get(), offs.get());
  exclusive_scan_u64(offs.get(), (uint64_t)n_over + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t n = read_u64(offs.get() + n_over, s_, ctx);
  DBuf<BoxD> bx(n, ar_);
  LAUNCH(k_oversized_to_boxd, grid_warps(n_over), 256, s_, list, n_over, offs.get(), bk.boxes.get(), bk.beg.get(), ix_.n_seqs,
         bx.get());
  const uint32_t n_rows = (uint32_t)(bk.NB / ix_.n_seqs);
  const int row_bits = bits_for(n_rows > 1 ? n_rows - 1 : 1);
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  bits_a_ = row_bits + 2 * seq_bits + 1;
  bits_b_ = row_bits + seq_bits + 33;
  REQUIRE(bits_a_ <= 63 && bits_b_ <= 63, IMPGX_E_INVALID, "batch too large for the packed merge keys; lower IMPGX_ROWS_PER_BATCH");
  DBuf<BoxD> acc;
  DBuf<uint64_t> is_root;
  DBuf<unsigned long long> rc(2, ar_);
  stage_a(bx.get(), n, n, acc, is_root);
  BatchOut none;
  stage_b(acc.get(), is_root.get(), n, rc.get(), n_rows, none, nullptr, &bk, out_cnt);
}

// masked_regions as the reference holds them: one SortedRanges per sequence. Returns the number of ranges.
static uint64_t validate_masks(const impgx_params &p, uint32_t n_seqs) {
  REQUIRE(p.mask_ranges || p.mask_offsets[n_seqs] == 0, IMPGX_E_INVALID, "mask_ranges is NULL");
  for (uint32_t q = 0; q < n_seqs; q++) {
    REQUIRE(p.mask_offsets[q] <= p.mask_offsets[q + 1], IMPGX_E_INVALID, "mask_offsets not monotone");
    for (uint64_t k = p.mask_offsets[q]; k < p.mask_offsets[q + 1]; k++) {
      const int32_t a = p.mask_ranges[2 * k], b = p.mask_ranges[2 * k + 1];
      REQUIRE(a <= b && (k == p.mask_offsets[q] || p.mask_ranges[2 * k - 1] < a), IMPGX_E_INVALID,
              "masked regions of a sequence must be sorted and disjoint (a SortedRanges)");
    }
  }
  return p.mask_offsets[n_seqs];
}

void Runner::prepare(const impgx_range *d_ranges, uint32_t n_rows) {
  d_counters_.alloc(4, ar_);
  // validation (perform_query bounds checks)
  {
    DBuf<int> bad(1, ar_);
    CUDA_CHECK(cudaMemsetAsync(bad.get(), 0, 4, s_));
    LAUNCH(k_validate, grid_threads(n_rows), 256, s_, d_ranges, n_rows, ix_.seq_len, ix_.n_seqs, bad.get());
    int hb = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hb, bad.get(), 4, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    REQUIRE(hb == 0, IMPGX_E_INVALID,
            "range " + std::to_string(hb - 1) +
                ": unknown target id, start >= end, negative start or end beyond the sequence length");
  }
  if (p_.subset_mask) {
    d_subset_.alloc(ix_.n_seqs, ar_);
    CUDA_CHECK(cudaMemcpyAsync(d_subset_.get(), p_.subset_mask, ix_.n_seqs, cudaMemcpyHostToDevice, s_));
    ctx.h2d_bytes += ix_.n_seqs;
    d_row_target_.alloc(n_rows, ar_);
    LAUNCH(k_row_targets, grid_threads(n_rows), 256, s_, d_ranges, n_rows, d_row_target_.get());
  }
  const bool transitive = p_.mode != IMPGX_MODE_QUERY && p_.mode != IMPGX_MODE_MULTI_QUERY;
  masked_ = transitive && p_.mask_offsets != nullptr;
  if (masked_) {
    const uint64_t nm = validate_masks(p_, ix_.n_seqs);
    d_mask_off_.alloc((uint64_t)ix_.n_seqs + 1, ar_);
    d_mask_rng_.alloc(std::max<uint64_t>(nm, 1), ar_);
    CUDA_CHECK(cudaMemcpyAsync(d_mask_off_.get(), p_.mask_offsets, ((size_t)ix_.n_seqs + 1) * 8, cudaMemcpyHostToDevice, s_));
    if (nm) CUDA_CHECK(cudaMemcpyAsync(d_mask_rng_.get(), p_.mask_ranges, nm * 8, cudaMemcpyHostToDevice, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    ctx.h2d_bytes += ((size_t)ix_.n_seqs + 1) * 8 + nm * 8;
  }
}

void Runner::run(const impgx_range *d_ranges, uint32_t n_rows, bool bed, BatchOut &out) {
  REQUIRE(p_.mode <= IMPGX_MODE_MULTI_DFS, IMPGX_E_INVALID, "unknown mode");
  const bool multi = p_.mode >= IMPGX_MODE_MULTI_QUERY;
  const bool query_mode = p_.mode == IMPGX_MODE_QUERY || p_.mode == IMPGX_MODE_MULTI_QUERY;
  const bool dfs_like = p_.mode == IMPGX_MODE_DFS || p_.mode == IMPGX_MODE_MULTI_BFS || p_.mode == IMPGX_MODE_MULTI_DFS;
  multi_order_ = multi ? (query_mode ? 2 : 1) : 0;
  REQUIRE(!(bed && p_.store_cigar), IMPGX_E_INVALID, "BED output carries no CIGAR (src/main.rs:7447)");
  REQUIRE(idx_->owner.empty(), IMPGX_E_INVALID,
          "this index is one shard of a target-sharded index: use impgx_query_batch_bed_sharded");
  prepare(d_ranges, n_rows);

  DBuf<Frontier> fr(n_rows, ar_);
  LAUNCH(k_init_frontier, grid_threads(n_rows), 256, s_, d_ranges, n_rows, fr.get());
  uint64_t nF = n_rows;
  std::vector<LevelHits> levels;

  // direct BED path: boxes straight from the hits (no raw result assembly)
  const bool direct = bed && !getenv("IMPGX_BED_GENERIC");
  BedSink sink;
  uint64_t prior = 0;  // hits held by the sorted levels so far
  if (direct) {
    sink.counters.alloc(2, ar_);
    CUDA_CHECK(cudaMemsetAsync(sink.counters.get(), 0, 16, s_));
  }
  auto sink_for = [&](uint32_t level) -> BedSink * {
    // MultiImpg order is not (range, visit rank): those levels are always sorted
    if (!direct || multi || nF >= (1ull << 26) || level >= 58) return nullptr;
    sink.prefix = (uint64_t)n_rows + prior;
    sink.level = (uint32_t)levels.size() + 1;  // ord level behind the seeds and every ordered level so far
    return &sink;
  };
  // Bucket path of the direct BED merge (bucket_kernels.cuh): the boxes that already exist (seeds, ordered
  // levels) are counted first, the last hop counts its hits while it stabs and writes them into the buckets.
  const bool buckets = direct && bucket_mode(n_rows);
  Buckets bk;
  auto last_hop = [&](bool closed, bool clip, uint32_t level) {
    BedSink *sk = sink_for(level);
    if (sk && buckets && nF > 0) {
      prefix_boxes(sink, d_ranges, n_rows, levels, query_mode);
      bk_begin(bk, n_rows);
      bk.level = sink.level;
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/false);
      levels.emplace_back();
      stab_and_lift(fr, nF, closed, clip, sk, levels.back(), &bk,
                    [&]() { bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/true); });
    } else {
      levels.emplace_back();
      stab_and_lift(fr, nF, closed, clip, sk, levels.back());
    }
  };

  if (query_mode) {
    last_hop(/*closed=*/true, /*clip=*/false, 0);
  } else {
    // seed: visited[target].insert(range) on an empty set returns the range
    // itself (bounds were validated), which is output and, if long enough,
    // becomes the level-0 frontier (src/impg.rs:2337-2373)
    Visited V;
    if (masked_) {
      // visited[target] = mask[target] + range; the unmasked pieces are the self intervals (a level of
      // their own, first) and the level-0 frontier
      DBuf<uint64_t> off((uint64_t)n_rows + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(off.get() + n_rows, 0, 8, s_));
      LAUNCH(k_seed_mask_caps, grid_threads(n_rows), 256, s_, d_ranges, n_rows, d_mask_off_.get(), off.get());
      exclusive_scan_u64(off.get(), (uint64_t)n_rows + 1, sc_, s_);
      ctx.launches += 2;
      const uint64_t cap = read_u64(off.get() + n_rows, s_, ctx);
      DBuf<int2> lists(cap, ar_), pieces(cap, ar_);
      DBuf<uint32_t> list_len(n_rows, ar_), piece_cnt(n_rows, ar_);
      LAUNCH(k_seed_masked, grid_threads(n_rows), 256, s_, d_ranges, n_rows, d_mask_off_.get(), d_mask_rng_.get(),
             ix_.seq_len, off.get(), lists.get(), pieces.get(), list_len.get(), piece_cnt.get());
      DBuf<uint64_t> ls((uint64_t)n_rows + 1, ar_), ps((uint64_t)n_rows + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(ls.get() + n_rows, 0, 8, s_));
      CUDA_CHECK(cudaMemsetAsync(ps.get() + n_rows, 0, 8, s_));
      LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, list_len.get(), (uint64_t)n_rows, ls.get());
      LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, piece_cnt.get(), (uint64_t)n_rows, ps.get());
      exclusive_scan_u64(ls.get(), (uint64_t)n_rows + 1, sc_, s_);
      exclusive_scan_u64(ps.get(), (uint64_t)n_rows + 1, sc_, s_);
      ctx.launches += 4;
      uint64_t nl = 0, np = 0;
      read_u64x2(ls.get() + n_rows, ps.get() + n_rows, nl, np, s_, ctx);
      V.keys.alloc(nl, ar_);
      V.start.alloc(nl, ar_);
      V.end.alloc(nl, ar_);
      V.n = nl;
      levels.emplace_back();
      LevelHits &sl = levels.back();
      sl.seeds = true;
      sl.n = np;
      sl.hits.alloc(np, ar_);
      if (p_.store_cigar) {
        sl.entry.alloc(np, ar_);
        sl.slices.alloc(np, ar_);
        if (np) LAUNCH(k_fill_seed_slices, grid_threads(np), 256, s_, sl.entry.get(), sl.slices.get(), np);
      }
      DBuf<Frontier> f2(np, ar_);
      LAUNCH(k_seed_masked_compact, grid_threads(n_rows), 256, s_, d_ranges, n_rows, off.get(), lists.get(), pieces.get(),
             list_len.get(), piece_cnt.get(), ls.get(), ps.get(), V.keys.get(), V.start.get(), V.end.get(),
             sl.hits.get(), f2.get());
      CUDA_CHECK(cudaStreamSynchronize(s_));
      fr = std::move(f2);
      nF = np;
      prior += np;
    } else {
      V.keys.alloc(n_rows, ar_);
      V.start.alloc(n_rows, ar_);
      V.end.alloc(n_rows, ar_);
      V.n = n_rows;
      LAUNCH(k_seed_visited, grid_threads(n_rows), 256, s_, d_ranges, n_rows, V.keys.get(), V.start.get(), V.end.get());
    }
    if (p_.min_transitive_len > 0 && nF > 0) {
      // drop rows shorter than min_transitive_len from the frontier
      DBuf<uint64_t> flag(nF + 1, ar_), scan(nF + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(flag.get() + nF, 0, 8, s_));
      LAUNCH(k_frontier_len_flags, grid_threads(nF), 256, s_, fr.get(), nF, p_.min_transitive_len, flag.get());
      CUDA_CHECK(cudaMemcpyAsync(scan.get(), flag.get(), (nF + 1) * 8, cudaMemcpyDeviceToDevice, s_));
      exclusive_scan_u64(scan.get(), nF + 1, sc_, s_);
      ctx.launches += 2;
      uint64_t keep = read_u64(scan.get() + nF, s_, ctx);
      if (keep != nF) {
        DBuf<Frontier> f2(keep, ar_);
        LAUNCH(k_frontier_compact, grid_threads(nF), 256, s_, fr.get(), nF, flag.get(), scan.get(), f2.get());
        fr = std::move(f2);
        nF = keep;
      }
    }
    uint32_t depth = 0;
    if (dfs_like) {
      run_dfs(d_ranges, n_rows, fr, nF, V, levels);
      nF = 0;
    }
    while (nF > 0 && (p_.max_depth == 0 || depth < p_.max_depth)) {
      const bool last = p_.max_depth != 0 && depth + 1 >= p_.max_depth;
      if (last) {
        last_hop(/*closed=*/false, /*clip=*/true, depth);
      } else {
        levels.emplace_back();
        stab_and_lift(fr, nF, /*closed=*/false, /*clip=*/true, nullptr, levels.back());
      }
      prior += levels.back().n;
      depth++;
      DBuf<Frontier> next;
      uint64_t n_next = 0;
      if (!last) fold(levels.back(), n_rows, V, next, n_next);
      fr = std::move(next);
      nF = n_next;
    }
  }
  if (bk.laid_out) {
    merge_buckets(bk, n_rows, out);
    return;
  }
  if (direct && levels.size() < 60) {
    {
      WallTimer wt(ctx.w_assemble);
      prefix_boxes(sink, d_ranges, n_rows, levels, query_mode);
      sink.n = sink.prefix + sink.n_raw;
    }
    if (buckets && sink.n_raw == 0) {  // no hop wrote into buckets (the walk ended early, or a DFS-like walk)
      bk_begin(bk, n_rows);
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/false);
      bk_layout(bk);
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/true);
      merge_buckets(bk, n_rows, out);
      return;
    }
    bed_merge_direct(sink, n_rows, out);
    return;
  }
  BatchOut raw;
  assemble(d_ranges, n_rows, levels, query_mode, bed ? raw : out);
  if (bed) bed_merge(raw, n_rows, out);
}

}  // namespace impgx

// ============================================================ BED merge
namespace impgx {

struct Groups {
  DBuf<uint32_t> begins;  // G + 1
  uint64_t G = 0;
};

static void build_groups(const uint64_t *keys, uint64_t n, Groups &g, Scratch &sc, cudaStream_t s, Ctx &ctx) {
  DBuf<uint64_t> head(n + 1, *sc.a), scan(n + 1, *sc.a);
  CUDA_CHECK(cudaMemsetAsync(head.get() + n, 0, 8, s));
  LAUNCH(k_heads_u64, grid_threads(n), 256, s, keys, n, head.get());
  CUDA_CHECK(cudaMemcpyAsync(scan.get(), head.get(), (n + 1) * 8, cudaMemcpyDeviceToDevice, s));
  exclusive_scan_u64(scan.get(), n + 1, sc, s);
  ctx.launches += 2;
  g.G = read_u64(scan.get() + n, s, ctx);
  g.begins.alloc(g.G + 1, *sc.a);
  LAUNCH(k_group_begins, grid_threads(n), 256, s, head.get(), scan.get(), n, g.begins.get());
  uint32_t n32 = (uint32_t)n;
  CUDA_CHECK(cudaMemcpyAsync(g.begins.get() + g.G, &n32, 4, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));  // n32 is a stack variable
}

void Runner::bed_merge(BatchOut &raw, uint32_t n_rows, BatchOut &out) {
  WallTimer wt(ctx.w_merge);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  const uint64_t R = raw.n_results;
  const int32_t d = p_.merge_distance;
  const bool ms = p_.merge_strands != 0;
  DBuf<uint32_t> row_cnt(n_rows, ar_);
  CUDA_CHECK(cudaMemsetAsync(row_cnt.get(), 0, (size_t)n_rows * 4, s_));
  out.row_off.alloc(n_rows + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(out.row_off.get(), 0, ((size_t)n_rows + 1) * 8, s_));
  out.n_results = 0;
  if (R > 0) {
    REQUIRE(R < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 results in one batch; lower IMPGX_ROWS_PER_BATCH");
    DBuf<uint32_t> row(R, ar_);
    LAUNCH(k_fill_rows, grid_threads(R), 256, s_, raw.row_off.get(), n_rows, R, row.get());
    ResCols rc{raw.q_id.get(), raw.q_first.get(), raw.q_last.get(), raw.t_id.get(), raw.t_first.get(), raw.t_last.get()};
    const int row_bits = bits_for(n_rows > 1 ? n_rows - 1 : 1);
    DBuf<Box> boxes;
    uint64_t nB = 0;
    if (d >= 0) {
      // ---- stage A
      const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
      REQUIRE(row_bits + 2 * seq_bits + 1 <= 64, IMPGX_E_INVALID, "batch too large for the packed merge key");
      DBuf<uint32_t> k1(R, ar_), perm(R, ar_);
      LAUNCH(k_m2d_key1, grid_threads(R), 256, s_, rc, R, k1.get(), perm.get());
      sort_pairs(k1, perm, R, 0, 32, sc_, s_, ctx);
      DBuf<uint64_t> k2(R, ar_);
      LAUNCH(k_m2d_key2, grid_threads(R), 256, s_, rc, row.get(), perm.get(), R, seq_bits, k2.get());
      sort_pairs(k2, perm, R, 0, row_bits + 2 * seq_bits + 1, sc_, s_, ctx);
      Groups g;
      build_groups(k2.get(), R, g, sc_, s_, ctx);
      DBuf<uint32_t> parent(R, ar_);
      DBuf<Box> box(R, ar_);
      DBuf<uint64_t> is_root(R + 1, ar_), root_scan(R + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(is_root.get() + R, 0, 8, s_));
      LAUNCH(k_merge2d, grid_threads(g.G, 128, 16), 128, s_, rc, row.get(), perm.get(), g.begins.get(), g.G, (int64_t)d,
             parent.get(), box.get(), is_root.get());
      CUDA_CHECK(cudaMemcpyAsync(root_scan.get(), is_root.get(), (R + 1) * 8, cudaMemcpyDeviceToDevice, s_));
      exclusive_scan_u64(root_scan.get(), R + 1, sc_, s_);
      ctx.launches += 2;
      nB = read_u64(root_scan.get() + R, s_, ctx);
      DBuf<Box> cb(nB, ar_);
      DBuf<uint32_t> ok(nB, ar_), ov(nB, ar_);
      LAUNCH(k_compact_boxes, grid_threads(R), 256, s_, box.get(), is_root.get(), root_scan.get(), R, cb.get(), ok.get(),
             ov.get());
      sort_pairs(ok, ov, nB, 0, bits_for(R), sc_, s_, ctx);
      boxes.alloc(nB, ar_);
      LAUNCH(k_gather<Box>, grid_threads(nB), 256, s_, cb.get(), ov.get(), nB, boxes.get());
    } else {
      nB = R;
      boxes.alloc(nB, ar_);
      LAUNCH(k_results_to_boxes, grid_threads(R), 256, s_, rc, row.get(), R, boxes.get());
    }
    OutCols oc;
    oc.cig_len = nullptr; oc.src_entry = nullptr; oc.src_slice = nullptr;
    auto alloc_out = [&](uint64_t n) {
      out.n_results = n;
      out.q_id.alloc(n, ar_); out.t_id.alloc(n, ar_);
      out.q_first.alloc(n, ar_); out.q_last.alloc(n, ar_);
      out.t_first.alloc(n, ar_); out.t_last.alloc(n, ar_);
      oc.q_id = out.q_id.get(); oc.q_first = out.q_first.get(); oc.q_last = out.q_last.get();
      oc.t_id = out.t_id.get(); oc.t_first = out.t_first.get(); oc.t_last = out.t_last.get();
    };
    if (d >= 0 || ms) {
      // ---- stage B
      DBuf<uint64_t> k1(nB, ar_);
      DBuf<uint32_t> perm(nB, ar_);
      LAUNCH(k_mq_key1, grid_threads(nB), 256, s_, boxes.get(), nB, k1.get(), perm.get());
      sort_pairs(k1, perm, nB, 0, 33, sc_, s_, ctx);
      DBuf<uint64_t> k2(nB, ar_);
      LAUNCH(k_mq_key2, grid_threads(nB), 256, s_, boxes.get(), perm.get(), nB, k2.get());
      sort_pairs(k2, perm, nB, 0, 32 + row_bits, sc_, s_, ctx);
      DBuf<Box> sorted(nB, ar_);
      LAUNCH(k_gather<Box>, grid_threads(nB), 256, s_, boxes.get(), perm.get(), nB, sorted.get());
      Groups g;
      build_groups(k2.get(), nB, g, sc_, s_, ctx);
      DBuf<Box> swept(nB, ar_);
      DBuf<uint32_t> cnt(g.G, ar_);
      LAUNCH(k_sweep, grid_threads(g.G, 128, 16), 128, s_, sorted.get(), g.begins.get(), g.G, d, ms ? 1 : 0, swept.get(),
             cnt.get());
      DBuf<uint64_t> scan(g.G + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(scan.get() + g.G, 0, 8, s_));
      LAUNCH(k_u32_to_u64, grid_threads(g.G), 256, s_, cnt.get(), g.G, scan.get());
      exclusive_scan_u64(scan.get(), g.G + 1, sc_, s_);
      ctx.launches += 2;
      const uint64_t M = read_u64(scan.get() + g.G, s_, ctx);
      alloc_out(M);
      LAUNCH(k_sweep_compact, grid_threads(g.G), 256, s_, swept.get(), g.begins.get(), cnt.get(), scan.get(), g.G, oc,
             row_cnt.get());
    } else {
      alloc_out(nB);
      LAUNCH(k_boxes_to_cols, grid_threads(nB), 256, s_, boxes.get(), nB, oc, row_cnt.get());
    }
    LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, row_cnt.get(), n_rows, out.row_off.get());
    exclusive_scan_u64(out.row_off.get(), (uint64_t)n_rows + 1, sc_, s_);
    ctx.launches += 2;
  }
  CUDA_CHECK(cudaEventRecord(e1, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));
  float msf = 0;
  cudaEventElapsedTime(&msf, e0, e1);
  ctx.merge_ms += msf;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}


// Stage A of output_results_bed (merge_adjusted_intervals_gap_2d, src/main.rs:12858-13011):
// group the valid boxes by (row, q, t, strand), pairwise union-find per group.
// acc[i] / is_root[i] for i < nv hold the merged boxes.
void Runner::stage_a(const BoxD *boxes, uint64_t nB, uint64_t nv, DBuf<BoxD> &acc, DBuf<uint64_t> &is_root) {
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  const int bits_a = bits_a_;
  acc.alloc(nv, ar_);
  is_root.alloc(nv, ar_);
  if (nv == 0) return;
  DBuf<uint64_t> ka(nB, ar_);
  DBuf<uint32_t> perm(nB, ar_);
  LAUNCH(k_bd_key_a, grid_threads(nB), 256, s_, boxes, nB, seq_bits, 1ull << bits_a, ka.get(), perm.get());
  sort_pairs(ka, perm, nB, 0, bits_a + 1, sc_, s_, ctx);
  Groups g;
  build_groups(ka.get(), nv, g, sc_, s_, ctx);
  DBuf<uint32_t> parent(nv, ar_);
  LAUNCH(k_merge2d_direct, grid_threads(g.G, 128, 16), 128, s_, boxes, perm.get(), g.begins.get(), g.G,
         (int64_t)p_.merge_distance, parent.get(), acc.get(), is_root.get());
}

// Stage B (merge_query_adjusted_intervals, src/main.rs:12474-12560): global sort of
// the roots by (row, q, start, strand), sweep per (row, q), compaction into `out`.
// is_root == nullptr: every box is a root.
void Runner::stage_b(const BoxD *acc, const uint64_t *is_root, uint64_t n, unsigned long long *d_root_counter,
                     uint32_t n_rows, BatchOut &out, uint32_t *row_cnt, Buckets *to_buckets, uint32_t *bucket_out_cnt) {
  if (n == 0) return;
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  const int bits_b = bits_b_;
  const int32_t d = p_.merge_distance;
  const bool ms = p_.merge_strands != 0;
  (void)n_rows;
  DBuf<uint64_t> kb(n, ar_);
  DBuf<uint32_t> permb(n, ar_);
  CUDA_CHECK(cudaMemsetAsync(d_root_counter, 0, 8, s_));
  LAUNCH(k_bd_key_b, grid_threads(n), 256, s_, acc, is_root, n, seq_bits, 1ull << bits_b, kb.get(), permb.get(),
         d_root_counter);
  sort_pairs(kb, permb, n, 0, bits_b + 1, sc_, s_, ctx);
  unsigned long long nr = 0;
  CUDA_CHECK(cudaMemcpyAsync(&nr, d_root_counter, 8, cudaMemcpyDeviceToHost, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));
  ctx.d2h_bytes += 8;
  if (nr == 0) return;
  DBuf<BoxD> sorted(nr, ar_);
  LAUNCH(k_gather<BoxD>, grid_threads(nr), 256, s_, acc, permb.get(), (uint64_t)nr, sorted.get());
  DBuf<uint64_t> seg(nr, ar_);
  LAUNCH(k_keys_shift, grid_threads(nr), 256, s_, kb.get(), (uint64_t)nr, 33, seg.get());
  Groups g;
  build_groups(seg.get(), nr, g, sc_, s_, ctx);
  DBuf<BoxD> swept(nr, ar_);
  DBuf<uint32_t> cnt(g.G, ar_);
  LAUNCH(k_sweep_direct, grid_threads(g.G, 128, 16), 128, s_, sorted.get(), kb.get(), g.begins.get(), g.G, d, ms ? 1 : 0,
         swept.get(), cnt.get());
  if (to_buckets) {  // the rows of these (row, q) segments join the staging slots of the bucket merge
    LAUNCH(k_groups_to_buckets, grid_threads(g.G), 256, s_, swept.get(), g.begins.get(), cnt.get(), g.G, ix_.n_seqs,
           to_buckets->beg.get(), to_buckets->boxes.get(), bucket_out_cnt);
    CUDA_CHECK(cudaStreamSynchronize(s_));
    return;
  }
  DBuf<uint64_t> scan(g.G + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(scan.get() + g.G, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(g.G), 256, s_, cnt.get(), g.G, scan.get());
  exclusive_scan_u64(scan.get(), g.G + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t M = read_u64(scan.get() + g.G, s_, ctx);
  OutCols oc = alloc_out_cols(out, M);
  LAUNCH(k_sweep_compact_direct, grid_threads(g.G), 256, s_, swept.get(), g.begins.get(), cnt.get(), scan.get(), g.G, oc,
         row_cnt);
}

// Both stages on chip, one (row, q) segment per warp / CTA (merge_kernels.cuh,
// k_merge_segments). Returns false (nothing written) when a segment exceeds
// SEG_MAX boxes: the caller then runs the global two-sort path.
bool Runner::merge_fused(const BoxSrc &src, uint64_t nB, BatchOut &out, uint32_t *row_cnt) {
  if (nB == 0) return true;
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  DBuf<uint64_t> ka0(nB, ar_);
  DBuf<uint32_t> perm0(nB, ar_);
  CUDA_CHECK(cudaMemsetAsync(d_counters_.get(), 0, 8, s_));
  LAUNCH(k_bd_key_a_src, grid_threads(nB), 256, s_, src, nB, seq_bits, 1ull << bits_a_, ka0.get(), perm0.get(),
         d_counters_.get());
  const uint64_t nv = read_u64((const uint64_t *)d_counters_.get(), s_, ctx);
  if (nv == 0) return true;
  DBuf<BoxD> swept(nv, ar_);
  DBuf<uint32_t> cnt;
  Groups g;
  {
    DBuf<uint64_t> &ka = ka0;
    DBuf<uint32_t> &perm = perm0;
    sort_pairs(ka, perm, nB, 0, bits_a_ + 1, sc_, s_, ctx);
    // segments = runs of equal (row, q)
    {
      DBuf<uint64_t> head(nv + 1, ar_), scan(nv + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(head.get() + nv, 0, 8, s_));
      LAUNCH(k_heads_u64_shift, grid_threads(nv), 256, s_, ka.get(), nv, seq_bits + 1, head.get());
      CUDA_CHECK(cudaMemcpyAsync(scan.get(), head.get(), (nv + 1) * 8, cudaMemcpyDeviceToDevice, s_));
      exclusive_scan_u64(scan.get(), nv + 1, sc_, s_);
      ctx.launches += 2;
      g.G = read_u64(scan.get() + nv, s_, ctx);
      g.begins.alloc(g.G + 1, ar_);
      LAUNCH(k_group_begins, grid_threads(nv), 256, s_, head.get(), scan.get(), nv, g.begins.get());
      // begins[G] = nv from a pinned word that no readback touches: the copy needs no synchronisation
      uint32_t *n32 = reinterpret_cast<uint32_t *>(readback_slot() + 7);
      *n32 = (uint32_t)nv;
      CUDA_CHECK(cudaMemcpyAsync(g.begins.get() + g.G, n32, 4, cudaMemcpyHostToDevice, s_));
    }
    DBuf<uint32_t> lists((uint64_t)SEG_CLASSES * g.G, ar_);
    DBuf<unsigned int> cls(SEG_CLASSES + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(cls.get(), 0, (SEG_CLASSES + 1) * 4, s_));
    int min_class = (int)env_u64("IMPGX_SEG_MIN_CLASS", 0);  // test hook: run the larger-segment kernels on small data
    if (min_class < 0 || min_class >= SEG_CLASSES) min_class = 0;
    LAUNCH(k_seg_classify, grid_threads(g.G), 256, s_, g.begins.get(), g.G, lists.get(), cls.get(), min_class);
    unsigned int hc[SEG_CLASSES + 1];
    {
      static_assert(sizeof(hc) <= 32, "class counters must fit the readback slot");
      uint64_t *h = readback_slot();
      CUDA_CHECK(cudaMemcpyAsync(h, cls.get(), sizeof(hc), cudaMemcpyDeviceToHost, s_));
      CUDA_CHECK(cudaStreamSynchronize(s_));
      memcpy(hc, h, sizeof(hc));
    }
    ctx.d2h_bytes += sizeof(hc);
    if (hc[SEG_CLASSES] != 0) return false;
    cnt.alloc(g.G, ar_);
    const int64_t d = p_.merge_distance;
    const int ms = p_.merge_strands ? 1 : 0;
    auto launch = [&](auto kern, unsigned threads, size_t smem, unsigned per_sm, unsigned seg_per_cta, int c) {
      if (!hc[c]) return;
      CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned grid =
          (unsigned)std::min<uint64_t>(((uint64_t)hc[c] + seg_per_cta - 1) / seg_per_cta, (uint64_t)sm_count() * per_sm);
      kern<<<grid, threads, smem, s_>>>(src, perm.get(), g.begins.get(), lists.get() + (uint64_t)c * g.G, hc[c], d, ms,
                                        swept.get(), cnt.get());
      CUDA_CHECK(cudaGetLastError());
      ctx.launches++;
    };
    launch(k_merge_segments<32, seg_cap(0)>, 256, (size_t)8 * seg_cap(0) * SEG_BYTES, 5, 8, 0);
    launch(k_merge_segments<32, seg_cap(1)>, 256, (size_t)8 * seg_cap(1) * SEG_BYTES, 2, 8, 1);
    launch(k_merge_segments<256, seg_cap(2)>, 256, (size_t)seg_cap(2) * SEG_BYTES, 8, 1, 2);
    launch(k_merge_segments<128, seg_cap(3)>, 128, (size_t)seg_cap(3) * SEG_BYTES, 5, 1, 3);
    launch(k_merge_segments<512, seg_cap(4)>, 512, (size_t)seg_cap(4) * SEG_BYTES, 1, 1, 4);
  }
  DBuf<uint64_t> scan(g.G + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(scan.get() + g.G, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(g.G), 256, s_, cnt.get(), g.G, scan.get());
  exclusive_scan_u64(scan.get(), g.G + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t M = read_u64(scan.get() + g.G, s_, ctx);
  OutCols oc = alloc_out_cols(out, M);
  LAUNCH(k_sweep_compact_direct, grid_threads(g.G), 256, s_, swept.get(), g.begins.get(), cnt.get(), scan.get(), g.G, oc,
         row_cnt);
  CUDA_CHECK(cudaStreamSynchronize(s_));  // `swept` and the lists are released on return
  return true;
}

OutCols Runner::alloc_out_cols(BatchOut &out, uint64_t n) {
  OutCols oc;
  oc.cig_len = nullptr; oc.src_entry = nullptr; oc.src_slice = nullptr;
  out.n_results = n;
  out.q_id.alloc(n, ar_); out.t_id.alloc(n, ar_);
  out.q_first.alloc(n, ar_); out.q_last.alloc(n, ar_);
  out.t_first.alloc(n, ar_); out.t_last.alloc(n, ar_);
  oc.q_id = out.q_id.get(); oc.q_first = out.q_first.get(); oc.q_last = out.q_last.get();
  oc.t_id = out.t_id.get(); oc.t_first = out.t_first.get(); oc.t_last = out.t_last.get();
  return oc;
}

// The raw hits of the last hop as BoxD records behind the seeds and the ordered levels (the
// paths that need one array: unsorted --no-merge output, the global two-sort merge, routing).
void Runner::materialize_boxes(BedSink &sink) {
  if (sink.n_raw == 0) return;
  const uint64_t H = sink.n_raw;
  DBuf<BoxD> full(sink.prefix + H, ar_);
  if (sink.prefix)
    CUDA_CHECK(cudaMemcpyAsync(full.get(), sink.boxes.get(), sink.prefix * sizeof(BoxD), cudaMemcpyDeviceToDevice, s_));
  LAUNCH(k_boxes_from_raw_level, grid_threads(H), 256, s_, sink.raw_hits.get(), sink.raw_tasks.get(),
         sink.raw_has_orig ? (const uint32_t *)sink.raw_orig.get() : nullptr, (const uint64_t *)nullptr,
         (const uint64_t *)nullptr, H, sink.level, p_.min_output_length, full.get() + sink.prefix, sink.counters.get(),
         (const uint32_t *)nullptr);
  CUDA_CHECK(cudaStreamSynchronize(s_));
  sink.boxes = std::move(full);
  sink.raw_hits.release();
  sink.raw_tasks.release();
  sink.raw_orig.release();
  sink.n_raw = 0;
  sink.n = sink.prefix + H;
}

// output_results_bed's two merges straight from the boxes of the batch. On a
// sharded index stage A runs where the hits were produced (owner of the target
// sequence: a (row, q, t, strand) group never spans ranks), the merged boxes
// travel to the owner of their query sequence, and stage B runs there.
void Runner::bed_merge_direct(BedSink &sink, uint32_t n_rows, BatchOut &out) {
  WallTimer wt(ctx.w_merge);
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  const int32_t d = p_.merge_distance;
  const bool ms = p_.merge_strands != 0;
  DBuf<uint32_t> row_cnt(n_rows, ar_);
  CUDA_CHECK(cudaMemsetAsync(row_cnt.get(), 0, (size_t)n_rows * 4, s_));
  out.row_off.alloc(n_rows + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(out.row_off.get(), 0, ((size_t)n_rows + 1) * 8, s_));
  out.n_results = 0;
  REQUIRE(sink.n < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 results in one batch; lower IMPGX_ROWS_PER_BATCH");
  const int row_bits = bits_for(n_rows > 1 ? n_rows - 1 : 1);
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  bits_a_ = row_bits + 2 * seq_bits + 1;
  bits_b_ = row_bits + seq_bits + 33;
  const bool unsorted_out = d < 0 && !ms;
  REQUIRE(!(unsorted_out && comm_), IMPGX_E_UNSUPPORTED,
          "--no-merge with --consider-strandness keeps the reference's unsorted result order, which a sharded "
          "index does not assemble; use an unsharded index");
  if (!unsorted_out)
    REQUIRE(bits_a_ <= 63 && bits_b_ <= 63, IMPGX_E_INVALID,
            "batch too large for the packed merge keys; lower IMPGX_ROWS_PER_BATCH");
  bool done = false;
  if (!unsorted_out && !comm_ && !getenv("IMPGX_MERGE_GLOBAL")) {
    // the common path: both merges on chip, straight from the raw hits of the last hop
    BoxSrc src{sink.boxes.get(), sink.prefix, sink.raw_hits.get(), sink.raw_tasks.get(),
               sink.raw_has_orig ? sink.raw_orig.get() : nullptr, sink.level, p_.min_output_length, nullptr};
    done = merge_fused(src, sink.n, out, row_cnt.get());
  }
  if (!done && comm_) {
    // sharded index: every valid box travels to the owner of its query sequence first, so that a
    // (row, q) segment is complete on one rank; what arrives is merged like an unsharded batch
    DBuf<BoxD> recv;
    uint64_t n_recv = 0;
    {
      BoxSrc src{sink.boxes.get(), sink.prefix, sink.raw_hits.get(), sink.raw_tasks.get(),
                 sink.raw_has_orig ? sink.raw_orig.get() : nullptr, sink.level, p_.min_output_length, sink.raw_gmap};
      route_boxes(src, sink.n, recv, n_recv);
    }
    sink.boxes.release();
    sink.raw_hits.release();
    sink.raw_tasks.release();
    sink.raw_orig.release();
    REQUIRE(n_recv < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 results in one batch; lower IMPGX_ROWS_PER_BATCH");
    bool fused = false;
    if (n_recv > 0 && !getenv("IMPGX_MERGE_GLOBAL")) {
      BoxSrc src{recv.get(), n_recv, nullptr, nullptr, nullptr, 0, -1, nullptr};
      fused = merge_fused(src, n_recv, out, row_cnt.get());
    }
    if (n_recv > 0 && !fused) {
      DBuf<BoxD> acc;
      DBuf<uint64_t> is_root;
      stage_a(recv.get(), n_recv, n_recv, acc, is_root);
      stage_b(acc.get(), is_root.get(), n_recv, sink.counters.get() + 1, n_rows, out, row_cnt.get());
    }
  } else if (!done) {
    // the remaining paths work on one BoxD array
    materialize_boxes(sink);
    const uint64_t nB = sink.n;
    unsigned long long nv = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nv, sink.counters.get(), 8, cudaMemcpyDeviceToHost, s_));
    CUDA_CHECK(cudaStreamSynchronize(s_));
    ctx.d2h_bytes += 8;
    if (unsorted_out) {
      if (nv > 0) {
        // nothing merges and nothing is sorted (src/main.rs:12479,12859): reference order per row
        DBuf<uint64_t> k1(nB, ar_);
        DBuf<uint32_t> perm(nB, ar_);
        LAUNCH(k_bd_key_ord, grid_threads(nB), 256, s_, sink.boxes.get(), nB, k1.get(), perm.get());
        sort_pairs(k1, perm, nB, 0, 64, sc_, s_, ctx);
        DBuf<uint64_t> k2(nB, ar_);
        LAUNCH(k_bd_key_row, grid_threads(nB), 256, s_, sink.boxes.get(), perm.get(), nB, n_rows, k2.get());
        sort_pairs(k2, perm, nB, 0, bits_for(n_rows), sc_, s_, ctx);
        OutCols oc = alloc_out_cols(out, nv);
        LAUNCH(k_boxd_to_cols, grid_threads(nv), 256, s_, sink.boxes.get(), perm.get(), (uint64_t)nv, oc, row_cnt.get());
      }
    } else if (nv > 0) {
      DBuf<BoxD> acc;
      DBuf<uint64_t> is_root;
      stage_a(sink.boxes.get(), nB, nv, acc, is_root);
      stage_b(acc.get(), is_root.get(), nv, sink.counters.get() + 1, n_rows, out, row_cnt.get());
    }
  }
  LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, row_cnt.get(), n_rows, out.row_off.get());
  exclusive_scan_u64(out.row_off.get(), (uint64_t)n_rows + 1, sc_, s_);
  ctx.launches += 2;
  CUDA_CHECK(cudaEventRecord(e1, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));
  float msf = 0;
  cudaEventElapsedTime(&msf, e0, e1);
  ctx.merge_ms += msf;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

// ============================================================ sharded index
// Partition `n` records by destination rank (dest[i] in [0, N], N = "nowhere")
// and exchange them: returns the received records, grouped by source rank.
// `gather(idx, count, send)` launches the kernel that writes record idx[i] to send[i].
template <class T, class Gather>
static void exchange_by_dest(Comm &cm, Gather gather, DBuf<uint32_t> &dest, DBuf<uint32_t> &idx,
                             const unsigned long long *d_dest_cnt, uint64_t n, DBuf<T> &recv, uint64_t &n_recv,
                             Arena &ar, Scratch &sc, cudaStream_t s, Ctx &ctx) {
  const int N = cm.size();
  std::vector<uint64_t> cnt((size_t)N + 1, 0);
  CUDA_CHECK(cudaMemcpyAsync(cnt.data(), d_dest_cnt, ((size_t)N + 1) * 8, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx.d2h_bytes += ((size_t)N + 1) * 8;
  std::vector<uint64_t> all((size_t)N * N);
  cm.allgather_u64(cnt.data(), (size_t)N, all.data(), s);
  std::vector<uint64_t> send_off((size_t)N, 0), recv_cnt((size_t)N, 0), recv_off((size_t)N, 0);
  uint64_t so = 0, ro = 0;
  for (int p = 0; p < N; p++) {
    send_off[p] = so;
    so += cnt[p];
    recv_cnt[p] = all[(size_t)p * N + cm.rank()];
    recv_off[p] = ro;
    ro += recv_cnt[p];
  }
  n_recv = ro;
  REQUIRE(n_recv < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 records received in one exchange; lower IMPGX_ROWS_PER_BATCH");
  recv.alloc(n_recv, ar);
  DBuf<T> send(so, ar);
  if (n) {
    sort_pairs(dest, idx, n, 0, bits_for((uint64_t)N), sc, s, ctx);  // stable: source order kept per destination
    if (so) {
      gather(idx.get(), so, send.get());
      CUDA_CHECK(cudaGetLastError());
      ctx.launches++;
    }
  }
  cudaEvent_t x0, x1;
  CUDA_CHECK(cudaEventCreate(&x0));
  CUDA_CHECK(cudaEventCreate(&x1));
  CUDA_CHECK(cudaEventRecord(x0, s));
  cm.alltoallv(send.get(), cnt.data(), send_off.data(), recv.get(), recv_cnt.data(), recv_off.data(), sizeof(T), s);
  CUDA_CHECK(cudaEventRecord(x1, s));
  CUDA_CHECK(cudaStreamSynchronize(s));  // `send` is released on return
  float xms = 0;
  cudaEventElapsedTime(&xms, x0, x1);
  ctx.exch_ms += xms;
  cudaEventDestroy(x0);
  cudaEventDestroy(x1);
}

// Non-last hop on a shard: the accepted hits go to the owner of the sequence
// they land on; what arrives is brought into the reference's hit order
// (global frontier index, visit rank) for the fold.
void Runner::route_hits(const Lifted &L, const uint32_t *gmap, LevelHits &lvl) {
  Comm &cm = *comm_;
  const uint32_t N = (uint32_t)cm.size();
  const uint64_t H = L.H;
  DBuf<unsigned long long> dcnt(N + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(dcnt.get(), 0, (N + 1) * 8, s_));
  DBuf<RoutedHit> routed(H, ar_), recv;
  DBuf<uint32_t> dest(H, ar_), idx(H, ar_);
  if (H)
    LAUNCH(k_route_hits, grid_threads(H), 256, s_, L.hits.get(), L.tasks.get(), L.d_orig, gmap, idx_->d_owner, H, N,
           routed.get(), dest.get(), idx.get(), dcnt.get());
  uint64_t n_recv = 0;
  const RoutedHit *rp = routed.get();
  cudaStream_t st = s_;
  exchange_by_dest(
      cm, [=](const uint32_t *ix, uint64_t cnt, RoutedHit *send) { k_gather<RoutedHit><<<grid_threads(cnt), 256, 0, st>>>(rp, ix, cnt, send); },
      dest, idx, dcnt.get(), H, recv, n_recv, ar_, sc_, s_, ctx);
  lvl.n = n_recv;
  if (n_recv == 0) return;
  lvl.hits.alloc(n_recv, ar_);
  DBuf<Hit> tmp(n_recv, ar_);
  DBuf<uint64_t> keys(n_recv, ar_);
  DBuf<uint32_t> perm(n_recv, ar_);
  LAUNCH(k_routed_to_hits, grid_threads(n_recv), 256, s_, recv.get(), n_recv, tmp.get(), keys.get(), perm.get());
  sort_pairs(keys, perm, n_recv, 0, 64, sc_, s_, ctx);
  LAUNCH(k_gather<Hit>, grid_threads(n_recv), 256, s_, tmp.get(), perm.get(), n_recv, lvl.hits.get());
  CUDA_CHECK(cudaStreamSynchronize(s_));
}

// All-gather-v of the next frontier (the exchange north_star names): every rank
// learns the global frontier, whose order (row, sequence, start) is the
// reference's frontier order; a rank keeps the ranges on the sequences it owns,
// tagged with their global index.
void Runner::global_frontier(DBuf<Frontier> &fr, uint64_t &nF, DBuf<uint32_t> &gmap, uint64_t total,
                             const std::vector<uint64_t> &cnt) {
  Comm &cm = *comm_;
  const int N = cm.size();
  std::vector<uint64_t> off((size_t)N, 0);
  for (int p = 1; p < N; p++) off[p] = off[p - 1] + cnt[p - 1];
  DBuf<Frontier> all(total, ar_);
  cm.allgatherv(fr.get(), nF, all.get(), cnt.data(), off.data(), sizeof(Frontier), s_);
  const int seq_bits = bits_for(ix_.n_seqs > 1 ? ix_.n_seqs - 1 : 1);
  DBuf<uint64_t> keys(total, ar_);
  DBuf<uint32_t> perm(total, ar_);
  LAUNCH(k_frontier_gkeys, grid_threads(total), 256, s_, all.get(), total, seq_bits, keys.get(), perm.get());
  sort_pairs(keys, perm, total, 0, 26 + seq_bits, sc_, s_, ctx);  // row < 2^26 (checked in run_sharded)
  DBuf<uint64_t> flag(total + 1, ar_), scan(total + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(flag.get() + total, 0, 8, s_));
  LAUNCH(k_frontier_own_flags, grid_threads(total), 256, s_, all.get(), perm.get(), total, idx_->d_owner,
         (uint32_t)cm.rank(), flag.get());
  CUDA_CHECK(cudaMemcpyAsync(scan.get(), flag.get(), (total + 1) * 8, cudaMemcpyDeviceToDevice, s_));
  exclusive_scan_u64(scan.get(), total + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t mine = read_u64(scan.get() + total, s_, ctx);
  REQUIRE(mine == nF, IMPGX_E_CUDA, "sharded frontier: a rank produced ranges on sequences it does not own");
  DBuf<Frontier> f2(mine, ar_);
  DBuf<uint32_t> g2(mine, ar_);
  LAUNCH(k_frontier_take_owned, grid_threads(total), 256, s_, all.get(), perm.get(), total, flag.get(), scan.get(),
         f2.get(), g2.get());
  CUDA_CHECK(cudaStreamSynchronize(s_));
  fr = std::move(f2);
  gmap = std::move(g2);
}

// Valid boxes travel to the owner of their query sequence (both BED merges group by (row, q));
// they are gathered straight from the box source (raw hits of the last hop included).
void Runner::route_boxes(const BoxSrc &src, uint64_t nB, DBuf<BoxD> &recv, uint64_t &n_recv) {
  Comm &cm = *comm_;
  const uint32_t N = (uint32_t)cm.size();
  DBuf<unsigned long long> dcnt(N + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(dcnt.get(), 0, (N + 1) * 8, s_));
  DBuf<uint32_t> dest(nB, ar_), idx(nB, ar_);
  if (nB) LAUNCH(k_box_dest, grid_threads(nB), 256, s_, src, nB, idx_->d_owner, N, dest.get(), idx.get(), dcnt.get());
  cudaStream_t st = s_;
  exchange_by_dest(
      cm, [=](const uint32_t *ix, uint64_t cnt, BoxD *send) { k_gather_src<<<grid_threads(cnt), 256, 0, st>>>(src, ix, cnt, send); },
      dest, idx, dcnt.get(), nB, recv, n_recv, ar_, sc_, s_, ctx);
}

// Sharded index, after the last hop: every bucket is reduced where its boxes were produced (stage A + the
// boxes no sweep can see dropped, bucket_kernels.cuh), what survives travels to the owner of the query
// sequence (all-to-all-v over the comm), grouped by destination through a destination-major scan.
void Runner::reduce_and_route(Buckets &bk, uint32_t n_rows, DBuf<BoxD> &recv, uint64_t &n_recv) {
  WallTimer wt(ctx.w_merge);
  Comm &cm = *comm_;
  const int N = cm.size();
  cudaEvent_t e0, e1, x0, x1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventCreate(&x0));
  CUDA_CHECK(cudaEventCreate(&x1));
  CUDA_CHECK(cudaEventRecord(e0, s_));
  const uint64_t NB = bk.NB;
  ctx.merge_boxes += bk.total;
  DBuf<uint32_t> out_cnt(NB + 1, ar_), cv(NB + 1, ar_), sv(NB + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(out_cnt.get(), 0, (NB + 1) * 4, s_));
  if (bk.total) ctx.merge_kernel_ms += run_bucket_kernels(bk, out_cnt.get(), /*reduce=*/true);
  CUDA_CHECK(cudaMemsetAsync(cv.get() + NB, 0, 4, s_));
  LAUNCH(k_send_counts, grid_threads(NB), 256, s_, out_cnt.get(), NB, n_rows, ix_.n_seqs, idx_->d_qorder, cv.get());
  exclusive_scan_u32(cv.get(), sv.get(), NB + 1, sc_, s_);
  ctx.launches += 2;
  std::vector<uint32_t> bnd((size_t)N + 1, 0);
  for (int p = 0; p <= N; p++)
    CUDA_CHECK(cudaMemcpyAsync(&bnd[p], sv.get() + (uint64_t)idx_->q_first[p] * n_rows, 4, cudaMemcpyDeviceToHost, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));
  ctx.d2h_bytes += ((size_t)N + 1) * 4;
  std::vector<uint64_t> cnt((size_t)N + 1, 0), send_off((size_t)N, 0), recv_cnt((size_t)N, 0), recv_off((size_t)N, 0);
  for (int p = 0; p < N; p++) {
    cnt[p] = bnd[p + 1] - bnd[p];
    send_off[p] = bnd[p];
  }
  const uint64_t n_send = bnd[N];
  std::vector<uint64_t> all((size_t)N * N);
  cm.allgather_u64(cnt.data(), (size_t)N, all.data(), s_);
  uint64_t ro = 0;
  for (int p = 0; p < N; p++) {
    recv_cnt[p] = all[(size_t)p * N + cm.rank()];
    recv_off[p] = ro;
    ro += recv_cnt[p];
  }
  n_recv = ro;
  REQUIRE(n_recv < (1ull << 32), IMPGX_E_INVALID, "more than 2^32 records received in one exchange; lower IMPGX_ROWS_PER_BATCH");
  recv.alloc(n_recv, ar_);
  DBuf<BoxD> send(n_send, ar_);
  if (n_send)
    LAUNCH(k_send_copy, grid_threads(NB), 256, s_, bk.boxes.get(), bk.beg.get(), out_cnt.get(), sv.get(), NB, n_rows,
           ix_.n_seqs, idx_->d_qorder, send.get());
  CUDA_CHECK(cudaEventRecord(x0, s_));
  cm.alltoallv(send.get(), cnt.data(), send_off.data(), recv.get(), recv_cnt.data(), recv_off.data(), sizeof(BoxD), s_);
  CUDA_CHECK(cudaEventRecord(x1, s_));
  CUDA_CHECK(cudaEventRecord(e1, s_));
  CUDA_CHECK(cudaStreamSynchronize(s_));  // `send` and the tables are released on return
  float ms = 0;
  cudaEventElapsedTime(&ms, x0, x1);
  ctx.exch_ms += ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx.merge_ms += ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(x0); cudaEventDestroy(x1);
}

void Runner::run_sharded(const impgx_range *d_ranges, uint32_t n_rows, BatchOut &out) {
  REQUIRE(comm_ && !idx_->owner.empty() && idx_->d_owner, IMPGX_E_INVALID, "run_sharded needs a shard index and a comm");
  REQUIRE((uint32_t)comm_->size() == idx_->shard_size && (uint32_t)comm_->rank() == idx_->shard_rank, IMPGX_E_INVALID,
          "comm rank/size do not match the shard index");
  REQUIRE(comm_->size() <= MAX_RANKS, IMPGX_E_INVALID, "too many ranks");
  REQUIRE(p_.mode == IMPGX_MODE_QUERY || p_.mode == IMPGX_MODE_BFS || p_.mode == IMPGX_MODE_DFS, IMPGX_E_UNSUPPORTED,
          "the sharded index runs Impg::query and the transitive BFS / DFS (MultiImpg walks: use an unsharded index)");
  REQUIRE(!p_.store_cigar, IMPGX_E_INVALID, "BED output carries no CIGAR (src/main.rs:7447)");
  REQUIRE(n_rows < (1u << 26), IMPGX_E_INVALID, "more than 2^26 rows in one batch; lower IMPGX_ROWS_PER_BATCH");
  Comm &cm = *comm_;
  const int N = cm.size();
  const uint32_t me = (uint32_t)cm.rank();
  const bool dfs = p_.mode == IMPGX_MODE_DFS;
  const bool bfs = p_.mode == IMPGX_MODE_BFS || dfs;  // a transitive walk: seeds, visited sets, clipping
  prepare(d_ranges, n_rows);

  // Level-0 frontier: the seed ranges on targets this rank owns. Every rank derives the GLOBAL seed list (one range
  // per row, or — under masked_regions — the unmasked pieces of every row, src/impg.rs:2337-2373) on its own: that
  // is rows-sized work and needs no exchange; the global frontier index of a seed is its index in that list.
  // Visited sets are seeded for every row; only the owner of a sequence ever touches its entries.
  DBuf<Frontier> fr;
  DBuf<uint32_t> gmap;
  uint64_t nF = 0;
  Visited V;
  DBuf<BoxD> seed_boxes;  // masked: the self-interval pieces are a level of their own (ord level 1)
  uint64_t n_seed_boxes = 0;
  const uint32_t level_base = masked_ ? 1u : 0u;  // ord level of hop d is d + 1 + level_base
  {
    DBuf<Frontier> all;
    uint64_t n_all = n_rows;
    if (masked_) {
      DBuf<uint64_t> off((uint64_t)n_rows + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(off.get() + n_rows, 0, 8, s_));
      LAUNCH(k_seed_mask_caps, grid_threads(n_rows), 256, s_, d_ranges, n_rows, d_mask_off_.get(), off.get());
      exclusive_scan_u64(off.get(), (uint64_t)n_rows + 1, sc_, s_);
      ctx.launches += 2;
      const uint64_t cap = read_u64(off.get() + n_rows, s_, ctx);
      DBuf<int2> lists(cap, ar_), pieces(cap, ar_);
      DBuf<uint32_t> list_len(n_rows, ar_), piece_cnt(n_rows, ar_);
      LAUNCH(k_seed_masked, grid_threads(n_rows), 256, s_, d_ranges, n_rows, d_mask_off_.get(), d_mask_rng_.get(),
             ix_.seq_len, off.get(), lists.get(), pieces.get(), list_len.get(), piece_cnt.get());
      DBuf<uint64_t> ls((uint64_t)n_rows + 1, ar_), ps((uint64_t)n_rows + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(ls.get() + n_rows, 0, 8, s_));
      CUDA_CHECK(cudaMemsetAsync(ps.get() + n_rows, 0, 8, s_));
      LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, list_len.get(), (uint64_t)n_rows, ls.get());
      LAUNCH(k_u32_to_u64, grid_threads(n_rows), 256, s_, piece_cnt.get(), (uint64_t)n_rows, ps.get());
      exclusive_scan_u64(ls.get(), (uint64_t)n_rows + 1, sc_, s_);
      exclusive_scan_u64(ps.get(), (uint64_t)n_rows + 1, sc_, s_);
      ctx.launches += 4;
      uint64_t nl = 0, np = 0;
      read_u64x2(ls.get() + n_rows, ps.get() + n_rows, nl, np, s_, ctx);
      V.keys.alloc(nl, ar_);
      V.start.alloc(nl, ar_);
      V.end.alloc(nl, ar_);
      V.n = nl;
      DBuf<Hit> seed_hits(np, ar_);
      all.alloc(np, ar_);
      LAUNCH(k_seed_masked_compact, grid_threads(n_rows), 256, s_, d_ranges, n_rows, off.get(), lists.get(), pieces.get(),
             list_len.get(), piece_cnt.get(), ls.get(), ps.get(), V.keys.get(), V.start.get(), V.end.get(),
             seed_hits.get(), all.get());
      n_all = np;
      // the seed pieces as boxes (never filtered by min_output_length), valid on the owner of the row's target only
      seed_boxes.alloc(np, ar_);
      n_seed_boxes = np;
      DBuf<unsigned long long> scratch_cnt(1, ar_);
      if (np) {
        LAUNCH(k_boxes_from_sorted_level, grid_threads(np), 256, s_, seed_hits.get(), np, 1u, -1, seed_boxes.get(),
               scratch_cnt.get());
        LAUNCH(k_boxes_keep_owned, grid_threads(np), 256, s_, seed_boxes.get(), np, idx_->d_owner, me);
      }
      CUDA_CHECK(cudaStreamSynchronize(s_));
    } else {
      all.alloc(n_rows, ar_);
      LAUNCH(k_init_frontier, grid_threads(n_rows), 256, s_, d_ranges, n_rows, all.get());
      if (bfs) {
        V.keys.alloc(n_rows, ar_);
        V.start.alloc(n_rows, ar_);
        V.end.alloc(n_rows, ar_);
        V.n = n_rows;
        LAUNCH(k_seed_visited, grid_threads(n_rows), 256, s_, d_ranges, n_rows, V.keys.get(), V.start.get(), V.end.get());
      }
    }
    DBuf<uint64_t> flag(n_all + 1, ar_), scan(n_all + 1, ar_);
    CUDA_CHECK(cudaMemsetAsync(flag.get() + n_all, 0, 8, s_));
    if (n_all && dfs)  // the stacks are replicated: every seed that is long enough, whoever owns its target
      LAUNCH(k_frontier_len_flags, grid_threads(n_all), 256, s_, all.get(), n_all, p_.min_transitive_len, flag.get());
    else if (n_all)
      LAUNCH(k_shard_seed_flags, grid_threads(n_all), 256, s_, all.get(), n_all, bfs ? p_.min_transitive_len : 0,
             idx_->d_owner, me, flag.get());
    CUDA_CHECK(cudaMemcpyAsync(scan.get(), flag.get(), (n_all + 1) * 8, cudaMemcpyDeviceToDevice, s_));
    exclusive_scan_u64(scan.get(), n_all + 1, sc_, s_);
    ctx.launches += 2;
    nF = read_u64(scan.get() + n_all, s_, ctx);
    DBuf<Frontier> f2(nF, ar_);
    DBuf<uint32_t> g2(nF, ar_);
    if (n_all) {
      LAUNCH(k_frontier_compact, grid_threads(n_all), 256, s_, all.get(), n_all, flag.get(), scan.get(), f2.get());
      LAUNCH(k_compact_indices, grid_threads(n_all), 256, s_, flag.get(), scan.get(), n_all, g2.get());
    }
    CUDA_CHECK(cudaStreamSynchronize(s_));
    fr = std::move(f2);
    gmap = std::move(g2);
  }

  BedSink sink;
  sink.counters.alloc(2, ar_);
  CUDA_CHECK(cudaMemsetAsync(sink.counters.get(), 0, 16, s_));
  std::vector<DBuf<BoxD>> level_boxes;  // boxes of the non-last hops
  std::vector<uint64_t> level_n;
  uint64_t prior = 0;
  uint32_t depth = 0;
  std::vector<uint64_t> cnt((size_t)N);
  // the seeds of the rows on owned targets and the boxes of the earlier hops as one BoxD array
  bool prefix_built = false;
  auto build_prefix = [&]() {
    WallTimer wt(ctx.w_assemble);
    sink.prefix = (uint64_t)n_rows + n_seed_boxes + prior;
    sink.boxes.alloc(sink.prefix, ar_);
    sink.n = sink.prefix + sink.n_raw;
    if (masked_) {  // the self intervals are the seed level; the per-row seed slots stay invalid
      CUDA_CHECK(cudaMemsetAsync(sink.boxes.get(), 0, (size_t)n_rows * sizeof(BoxD), s_));
      if (n_seed_boxes)
        CUDA_CHECK(cudaMemcpyAsync(sink.boxes.get() + n_rows, seed_boxes.get(), n_seed_boxes * sizeof(BoxD),
                                   cudaMemcpyDeviceToDevice, s_));
    } else {
      LAUNCH(k_boxes_from_seeds, grid_threads(n_rows), 256, s_, d_ranges, n_rows, p_.min_output_length, bfs ? 0 : 1,
             sink.boxes.get(), sink.counters.get(), (const uint32_t *)idx_->d_owner, me);
    }
    uint64_t off = (uint64_t)n_rows + n_seed_boxes;
    for (size_t l = 0; l < level_boxes.size(); l++) {
      if (level_n[l])
        CUDA_CHECK(cudaMemcpyAsync(sink.boxes.get() + off, level_boxes[l].get(), level_n[l] * sizeof(BoxD),
                                   cudaMemcpyDeviceToDevice, s_));
      off += level_n[l];
    }
    CUDA_CHECK(cudaStreamSynchronize(s_));
    level_boxes.clear();
    prefix_built = true;
  };
  // bucket path (bucket_kernels.cuh): local buckets, reduced where the boxes are produced, then routed
  const bool buckets = bucket_mode(n_rows);
  Buckets bk;
  bool bk_begun = false;
  if (dfs) {
    sharded_dfs(fr, nF, n_rows, V, 1 + level_base, level_boxes, level_n, prior, sink.counters.get());
    nF = 0;
  }
  for (;;) {
    if (dfs) break;
    // every rank learns every rank's frontier size: loop control must agree
    cm.allgather_u64(&nF, 1, cnt.data(), s_);
    uint64_t total = 0;
    for (int p = 0; p < N; p++) total += cnt[p];
    if (total == 0) break;
    if (bfs && p_.max_depth != 0 && depth >= p_.max_depth) break;
    if (!bfs && depth >= 1) break;
    REQUIRE(total < (1ull << 26) && depth < 58, IMPGX_E_INVALID,
            "global frontier exceeds 2^26 ranges in one batch; lower IMPGX_ROWS_PER_BATCH");
    if (depth > 0) global_frontier(fr, nF, gmap, total, cnt);
    const bool last = !bfs || (p_.max_depth != 0 && depth + 1 >= p_.max_depth);
    if (last && buckets) {
      // the last hop counts its hits per (row, q) while it stabs and writes them into the local buckets
      build_prefix();
      bk_begin(bk, n_rows);
      bk_begun = true;
      bk.level = depth + 1 + level_base;
      bk.gmap = gmap.get();
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/false);
      Lifted L;
      lift_core(fr, nF, /*closed=*/!bfs, /*clip=*/bfs, nullptr, L, &bk,
                [&]() { bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/true); });
      break;
    }
    Lifted L;
    DBuf<BoxD> boxes_h;
    BoxD *dst = nullptr;
    lift_core(fr, nF, /*closed=*/!bfs, /*clip=*/bfs,
              [&](uint64_t H) {
                if (!last) {
                  boxes_h.alloc(H, ar_);
                  dst = boxes_h.get();
                }
              },
              L);
    if (last) {
      // the raw hits stay where the liftover wrote them; they are gathered into the send buffers
      // of the box exchange with their global ordinals (BoxSrc)
      if (L.H) {
        sink.raw_hits = std::move(L.hits);
        sink.raw_tasks = std::move(L.tasks);
        sink.raw_has_orig = L.d_orig != nullptr;
        if (sink.raw_has_orig) sink.raw_orig = std::move(L.orig);
        sink.n_raw = L.H;
        sink.raw_gmap = gmap.get();
        sink.level = depth + 1 + level_base;
      }
      sink.filled = true;
      break;
    }
    if (L.H)
      LAUNCH(k_boxes_from_raw_level, grid_threads(L.H), 256, s_, L.hits.get(), L.tasks.get(), L.d_orig, L.offs.get(),
             (const uint64_t *)nullptr, L.H, depth + 1 + level_base, p_.min_output_length, dst, sink.counters.get(),
             gmap.get());
    depth++;
    prior += L.H;
    level_n.push_back(L.H);
    level_boxes.push_back(std::move(boxes_h));
    LevelHits lvl;
    route_hits(L, gmap.get(), lvl);
    DBuf<Frontier> next;
    uint64_t n_next = 0;
    fold(lvl, n_rows, V, next, n_next);
    fr = std::move(next);
    nF = n_next;
  }
  if (!prefix_built) build_prefix();
  if (buckets) {
    if (!bk_begun) {
      bk_begin(bk, n_rows);
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/false);
    }
    if (!bk.laid_out) {  // this rank had nothing to lift in the last hop (or the walk ended early)
      bk_layout(bk);
      bk_add_boxd(bk, sink.boxes.get(), sink.prefix, /*scatter=*/true);
    }
    DBuf<BoxD> recv;
    uint64_t n_recv = 0;
    reduce_and_route(bk, n_rows, recv, n_recv);
    // what arrives is merged like an unsharded batch (stage-A results carry their flag)
    Buckets fb;
    bk_begin(fb, n_rows);
    bk_add_boxd(fb, recv.get(), n_recv, /*scatter=*/false);
    bk_layout(fb);
    bk_add_boxd(fb, recv.get(), n_recv, /*scatter=*/true);
    merge_buckets(fb, n_rows, out);
    return;
  }
  bed_merge_direct(sink, n_rows, out);
}

// Transitive DFS (src/impg.rs:2057-2309). The walk of one row is inherently
// sequential (pop one range, expand, re-sort the stack), so the device runs it
// in lock step over the rows of the batch: every round each row with a
// non-empty stack pops its top range; the popped ranges are stabbed, lifted and
// folded together, the pieces are pushed with depth + 1 and every stack is
// re-sorted by (id, start) and merged exactly like the reference does after
// each pop (:2289-2304). The stacks of all rows live in one array sorted by
// (row, id, start); a row's top is the last entry of its segment.
// pop the top of every non-empty stack: the popped ranges that are walked on (depth below max_depth), one per row
void Runner::dfs_pop(const DBuf<DfsEntry> &stack, uint64_t n_stack, uint32_t n_rows, DBuf<uint64_t> &popped,
                     DBuf<uint32_t> &cur_depth, DBuf<Frontier> &fr, uint64_t &nF) {
  popped.alloc(n_stack + 1, ar_);
  DBuf<uint64_t> is_fr(n_rows + 1, ar_), fr_scan(n_rows + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(popped.get(), 0, (n_stack + 1) * 8, s_));
  CUDA_CHECK(cudaMemsetAsync(is_fr.get(), 0, ((size_t)n_rows + 1) * 8, s_));
  DBuf<Frontier> cand(n_rows, ar_);
  LAUNCH(k_dfs_pop, grid_threads(n_stack), 256, s_, stack.get(), n_stack, p_.max_depth,
         p_.mode == IMPGX_MODE_MULTI_BFS ? 1 : 0, popped.get(), cand.get(), is_fr.get(), cur_depth.get());
  CUDA_CHECK(cudaMemcpyAsync(fr_scan.get(), is_fr.get(), ((size_t)n_rows + 1) * 8, cudaMemcpyDeviceToDevice, s_));
  exclusive_scan_u64(fr_scan.get(), (uint64_t)n_rows + 1, sc_, s_);
  ctx.launches += 2;
  nF = read_u64(fr_scan.get() + n_rows, s_, ctx);
  fr.alloc(nF, ar_);
  if (nF) LAUNCH(k_frontier_compact, grid_threads(n_rows), 256, s_, cand.get(), (uint64_t)n_rows, is_fr.get(), fr_scan.get(), fr.get());
  CUDA_CHECK(cudaStreamSynchronize(s_));  // cand / is_fr / fr_scan are released on return
}

// new stack = kept entries + pushed pieces, sorted by (row, id, start), merged per row. False: every stack is empty.
bool Runner::dfs_restack(DBuf<DfsEntry> &stack, uint64_t &n_stack, const DBuf<uint64_t> &popped, const Frontier *pieces,
                         uint64_t n_pieces, const uint32_t *cur_depth, uint32_t n_rows) {
  DBuf<uint64_t> keep_scan(n_stack + 1, ar_);
  LAUNCH(k_dfs_keep_flags, grid_threads(n_stack + 1), 256, s_, popped.get(), n_stack, keep_scan.get());
  exclusive_scan_u64(keep_scan.get(), n_stack + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t kept = read_u64(keep_scan.get() + n_stack, s_, ctx);
  const uint64_t m = kept + n_pieces;
  if (m == 0) return false;
  DBuf<DfsEntry> tmp(m, ar_);
  LAUNCH(k_dfs_copy_kept, grid_threads(n_stack), 256, s_, stack.get(), n_stack, popped.get(), keep_scan.get(), tmp.get());
  if (n_pieces) LAUNCH(k_dfs_push, grid_threads(n_pieces), 256, s_, pieces, n_pieces, cur_depth, tmp.get() + kept);
  DBuf<uint32_t> sk(m, ar_), sp(m, ar_);
  LAUNCH(k_dfs_start_keys, grid_threads(m), 256, s_, tmp.get(), m, sk.get(), sp.get());
  sort_pairs(sk, sp, m, 0, 32, sc_, s_, ctx);
  DBuf<uint64_t> k2(m, ar_);
  LAUNCH(k_dfs_seq_keys, grid_threads(m), 256, s_, tmp.get(), sp.get(), m, k2.get());
  sort_pairs(k2, sp, m, 0, 32 + bits_for(n_rows), sc_, s_, ctx);
  DBuf<DfsEntry> sorted(m, ar_);
  LAUNCH(k_gather<DfsEntry>, grid_threads(m), 256, s_, tmp.get(), sp.get(), m, sorted.get());
  // per-row sequential merge (write/read sweep of :2291-2304)
  DBuf<uint64_t> rowkeys(m, ar_);
  LAUNCH(k_keys_shift, grid_threads(m), 256, s_, k2.get(), m, 32, rowkeys.get());
  Groups g;
  build_groups(rowkeys.get(), m, g, sc_, s_, ctx);
  DBuf<uint32_t> cnt(g.G, ar_);
  LAUNCH(k_dfs_merge, grid_threads(g.G, 128, 16), 128, s_, sorted.get(), g.begins.get(), g.G, cnt.get());
  DBuf<uint64_t> scan(g.G + 1, ar_);
  CUDA_CHECK(cudaMemsetAsync(scan.get() + g.G, 0, 8, s_));
  LAUNCH(k_u32_to_u64, grid_threads(g.G), 256, s_, cnt.get(), g.G, scan.get());
  exclusive_scan_u64(scan.get(), g.G + 1, sc_, s_);
  ctx.launches += 2;
  const uint64_t n_new = read_u64(scan.get() + g.G, s_, ctx);
  DBuf<DfsEntry> ns(n_new, ar_);
  LAUNCH(k_dfs_compact, grid_threads(g.G), 256, s_, sorted.get(), g.begins.get(), cnt.get(), scan.get(), g.G, ns.get());
  CUDA_CHECK(cudaStreamSynchronize(s_));  // the temporaries are released on return
  stack = std::move(ns);
  n_stack = n_new;
  return true;
}

void Runner::run_dfs(const impgx_range *d_ranges, uint32_t n_rows, DBuf<Frontier> &fr0, uint64_t nF0, Visited &V,
                     std::vector<LevelHits> &levels) {
  (void)d_ranges;
  DBuf<DfsEntry> stack(nF0, ar_);
  uint64_t n_stack = nF0;
  if (nF0) LAUNCH(k_dfs_init_stack, grid_threads(nF0), 256, s_, fr0.get(), nF0, stack.get());
  DBuf<uint32_t> cur_depth(n_rows, ar_);
  while (n_stack > 0) {
    DBuf<uint64_t> popped;
    DBuf<Frontier> fr;
    uint64_t nF = 0;
    dfs_pop(stack, n_stack, n_rows, popped, cur_depth, fr, nF);
    // ---- expand
    DBuf<Frontier> pieces;
    uint64_t n_pieces = 0;
    if (nF) {
      levels.emplace_back();
      // MultiImpg walks with Impg::query (closed visit, unclipped request: src/multi_impg.rs:875-883)
      const bool mq = p_.mode != IMPGX_MODE_DFS;
      stab_and_lift(fr, nF, /*closed=*/mq, /*clip=*/!mq, nullptr, levels.back());
      fold(levels.back(), n_rows, V, pieces, n_pieces, /*raw_pieces=*/true);
    }
    if (!dfs_restack(stack, n_stack, popped, pieces.get(), n_pieces, cur_depth.get(), n_rows)) break;
  }
}

// Transitive DFS on a target-sharded index. The stacks of all rows are REPLICATED: every rank pops the same ranges
// (no exchange needed to agree), lifts the popped ranges on the targets it owns, routes the hits to the owners of the
// sequences they land on (all-to-all-v, as in the BFS), folds them there, and the uncovered pieces of all ranks are
// all-gathered, so that every rank pushes the same pieces and re-sorts the same stacks. A row pops at most one range
// per round, so the reference's result order of a row is (round, visit rank): that is the ordinal of the boxes.
void Runner::sharded_dfs(DBuf<Frontier> &fr0, uint64_t nF0, uint32_t n_rows, Visited &V, uint32_t ord_level,
                         std::vector<DBuf<BoxD>> &level_boxes, std::vector<uint64_t> &level_n, uint64_t &prior,
                         unsigned long long *box_counter) {
  Comm &cm = *comm_;
  const int N = cm.size();
  const uint32_t me = (uint32_t)cm.rank();
  DBuf<DfsEntry> stack(nF0, ar_);
  uint64_t n_stack = nF0;
  if (nF0) LAUNCH(k_dfs_init_stack, grid_threads(nF0), 256, s_, fr0.get(), nF0, stack.get());
  DBuf<uint32_t> cur_depth(n_rows, ar_);
  std::vector<uint64_t> cnt((size_t)N), off((size_t)N);
  uint32_t round = 0;
  while (n_stack > 0) {
    REQUIRE(round < (1u << 26), IMPGX_E_INVALID, "transitive DFS on a sharded index: more than 2^26 rounds");
    DBuf<uint64_t> popped;
    DBuf<Frontier> fr;
    uint64_t nF = 0;
    dfs_pop(stack, n_stack, n_rows, popped, cur_depth, fr, nF);
    DBuf<Frontier> all_pieces;
    uint64_t n_all = 0;
    if (nF) {  // the same on every rank: the exchanges below are entered by all or none
      // the popped ranges on targets owned here, tagged with their index among the popped
      DBuf<uint64_t> flag(nF + 1, ar_), scan(nF + 1, ar_);
      CUDA_CHECK(cudaMemsetAsync(flag.get() + nF, 0, 8, s_));
      LAUNCH(k_shard_seed_flags, grid_threads(nF), 256, s_, fr.get(), nF, 0, idx_->d_owner, me, flag.get());
      CUDA_CHECK(cudaMemcpyAsync(scan.get(), flag.get(), (nF + 1) * 8, cudaMemcpyDeviceToDevice, s_));
      exclusive_scan_u64(scan.get(), nF + 1, sc_, s_);
      ctx.launches += 2;
      const uint64_t nM = read_u64(scan.get() + nF, s_, ctx);
      DBuf<Frontier> mine(nM, ar_);
      DBuf<uint32_t> gmap(nM, ar_);
      LAUNCH(k_frontier_compact, grid_threads(nF), 256, s_, fr.get(), nF, flag.get(), scan.get(), mine.get());
      LAUNCH(k_compact_indices, grid_threads(nF), 256, s_, flag.get(), scan.get(), nF, gmap.get());
      Lifted L;
      DBuf<BoxD> boxes_h;
      lift_core(mine, nM, /*closed=*/false, /*clip=*/true, [&](uint64_t H) { boxes_h.alloc(H, ar_); }, L);
      if (L.H)
        LAUNCH(k_boxes_dfs_round, grid_threads(L.H), 256, s_, L.hits.get(), L.H, ord_level, round, p_.min_output_length,
               boxes_h.get(), box_counter);
      prior += L.H;
      level_n.push_back(L.H);
      level_boxes.push_back(std::move(boxes_h));
      LevelHits lvl;
      route_hits(L, gmap.get(), lvl);
      DBuf<Frontier> pieces;
      uint64_t n_pieces = 0;
      fold(lvl, n_rows, V, pieces, n_pieces, /*raw_pieces=*/true);
      cm.allgather_u64(&n_pieces, 1, cnt.data(), s_);
      for (int p = 0; p < N; p++) {
        off[p] = n_all;
        n_all += cnt[p];
      }
      all_pieces.alloc(n_all, ar_);
      cm.allgatherv(pieces.get(), n_pieces, all_pieces.get(), cnt.data(), off.data(), sizeof(Frontier), s_);
    }
    if (!dfs_restack(stack, n_stack, popped, all_pieces.get(), n_all, cur_depth.get(), n_rows)) break;
    round++;
  }
}

// ============================================================ batch driver
// Scratch of one call: an arena and (host entry points) a stream of its own, leased from the index.
struct CallScratch {
  impgx_index *idx;
  std::unique_ptr<Arena> arena;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  CallScratch(impgx_index *i, bool want_stream) : idx(i), own_stream(want_stream) {
    {
      std::lock_guard<std::mutex> lock(idx->mu);
      if (!idx->arena_pool.empty()) {
        arena = std::move(idx->arena_pool.back());
        idx->arena_pool.pop_back();
      }
      if (own_stream && !idx->stream_pool.empty()) {
        stream = idx->stream_pool.back();
        idx->stream_pool.pop_back();
      }
    }
    if (!arena) arena.reset(new Arena());
    if (own_stream && !stream) CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  }
  ~CallScratch() {
    arena->reset();
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->arena_pool.push_back(std::move(arena));
    if (own_stream && stream) idx->stream_pool.push_back(stream);
  }
};

struct TooManyHits {};

static uint64_t env_u64(const char *name, uint64_t dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  return strtoull(v, nullptr, 10);
}


// ---- calls of a few rows: the whole walk in one launch (small_bfs.cuh). Returns false when the call is not
// eligible or a row did not fit the capacities; the caller then runs the batched path.
static bool small_eligible(const impgx_index *idx, size_t n, const impgx_params &p, bool bed, bool results_to_host,
                           const Comm *comm) {
  if (n == 0 || n > SB_MAX_ROWS || comm || !results_to_host || !idx->owner.empty()) return false;
  if (p.mode != IMPGX_MODE_BFS && p.mode != IMPGX_MODE_DFS && p.mode != IMPGX_MODE_QUERY) return false;
  if (p.store_cigar || !std::isnan(p.min_identity)) return false;  // the endpoint liftover carries no CIGAR
  if (bed && p.merge_distance < 0 && !p.merge_strands) return false;  // unsorted output: reference order per row
  // diagnostic / test switches of the batched path keep selecting it
  for (const char *v : {"IMPGX_NO_SMALL_BFS", "IMPGX_MERGE_SORTED", "IMPGX_MERGE_GLOBAL", "IMPGX_BED_GENERIC",
                        "IMPGX_FULL_SCAN", "IMPGX_SEG_MIN_CLASS", "IMPGX_ROWS_PER_BATCH", "IMPGX_NO_LOCALITY"})
    if (getenv(v)) return false;
  return true;
}

static bool try_small(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params &p, bool bed,
                      bool ranges_on_device, cudaStream_t s, Arena &arena, impgx_results *res, Ctx &ctx) {
  {
    std::lock_guard<std::mutex> lock(idx->mu);
    if (idx->small_skip > 0) {  // recent calls did not fit: do not pay for the attempt every time
      idx->small_skip--;
      return false;
    }
  }
  const uint32_t R = (uint32_t)n;
  const DevIndexView view = idx->view();
  const bool query_mode = p.mode == IMPGX_MODE_QUERY;
  const bool masked = !query_mode && p.mask_offsets != nullptr;
  SbParams sp{};
  sp.n_rows = R;
  sp.max_depth = p.max_depth;
  sp.min_transitive_len = p.min_transitive_len;
  sp.min_dist = p.min_distance_between_ranges;
  sp.min_out = p.min_output_length;
  sp.query_mode = query_mode ? 1 : 0;
  sp.dfs = p.mode == IMPGX_MODE_DFS ? 1 : 0;
  sp.bed = bed ? 1 : 0;

  const impgx_range *d_r = ranges;
  DBuf<impgx_range> d_stage;
  if (!ranges_on_device) {
    d_stage.alloc(R, arena);
    CUDA_CHECK(cudaMemcpyAsync(d_stage.get(), ranges, (size_t)R * sizeof(impgx_range), cudaMemcpyHostToDevice, s));
    ctx.h2d_bytes += (size_t)R * sizeof(impgx_range);
    d_r = d_stage.get();
  }
  DBuf<uint64_t> d_mask_off;
  DBuf<int2> d_mask_rng;
  if (masked) {
    const uint64_t nm = validate_masks(p, view.n_seqs);
    d_mask_off.alloc((uint64_t)view.n_seqs + 1, arena);
    d_mask_rng.alloc(std::max<uint64_t>(nm, 1), arena);
    CUDA_CHECK(cudaMemcpyAsync(d_mask_off.get(), p.mask_offsets, ((size_t)view.n_seqs + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nm) CUDA_CHECK(cudaMemcpyAsync(d_mask_rng.get(), p.mask_ranges, nm * 8, cudaMemcpyHostToDevice, s));
    ctx.h2d_bytes += ((size_t)view.n_seqs + 1) * 8 + nm * 8;
    sp.mask_off = d_mask_off.get();
    sp.mask_rng = d_mask_rng.get();
  }
  DBuf<uint8_t> d_subset;
  if (p.subset_mask) {
    d_subset.alloc(view.n_seqs, arena);
    CUDA_CHECK(cudaMemcpyAsync(d_subset.get(), p.subset_mask, view.n_seqs, cudaMemcpyHostToDevice, s));
    ctx.h2d_bytes += view.n_seqs;
    sp.subset = d_subset.get();
  }

  SbCall c{};
  c.row_bytes = sb_row_bytes();
  DBuf<char> rows((size_t)R * c.row_bytes, arena);
  c.rows = rows.get();
  DBuf<uint32_t> per_row((size_t)4 * R, arena);
  c.row_target = per_row.get();
  c.status = per_row.get() + R;
  c.n_res = per_row.get() + 2 * (size_t)R;
  c.n_bk = per_row.get() + 3 * (size_t)R;
  DBuf<SbCtl> ctl(R, arena);
  c.ctl = ctl.get();
  // zeroed counters: [0..1] stats, then the class list lengths
  DBuf<unsigned long long> zeroed(2 + SB_PHASES + (SEG_CLASSES + 2 + 1) / 2, arena);
  CUDA_CHECK(cudaMemsetAsync(zeroed.get(), 0, zeroed.bytes(), s));
  c.stats = zeroed.get();
  c.cls = reinterpret_cast<unsigned int *>(zeroed.get() + 2 + SB_PHASES);
  const uint64_t cap = (uint64_t)R * SB_CAP;
  DBuf<BoxRec> boxes;
  DBuf<uint32_t> bk, lists;
  if (bed) {
    boxes.alloc(cap, arena);
    bk.alloc(4 * cap, arena);
    lists.alloc((uint64_t)(SEG_CLASSES + 2) * cap, arena);
    c.boxes = boxes.get();
    c.bk_beg = bk.get();
    c.bk_cur = bk.get() + cap;
    c.bk_q = bk.get() + 2 * cap;
    c.out_cnt = bk.get() + 3 * cap;
    c.lists = lists.get();
  }
  const size_t hdr_bytes = (((size_t)R + 4) * 4 + 7) & ~(size_t)7;
  DBuf<char> outbuf(hdr_bytes + cap * sizeof(SbOut), arena);
  c.hdr = reinterpret_cast<uint32_t *>(outbuf.get());
  c.out = reinterpret_cast<SbOut *>(outbuf.get() + hdr_bytes);

  {
    static std::mutex attr_mu;
    static bool attr_done[64] = {false};
    std::lock_guard<std::mutex> lock(attr_mu);
    if (idx->device < 0 || idx->device >= 64 || !attr_done[idx->device]) {
      CUDA_CHECK(cudaFuncSetAttribute(k_small_bfs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM));
      CUDA_CHECK(cudaFuncSetAttribute(k_merge_buckets<32, seg_cap(0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * seg_cap(0) * BK_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(k_merge_buckets<32, seg_cap(1)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * seg_cap(1) * BK_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(k_merge_buckets<128, seg_cap(2)>, cudaFuncAttributeMaxDynamicSharedMemorySize, seg_cap(2) * BK_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(k_merge_buckets<128, seg_cap(3)>, cudaFuncAttributeMaxDynamicSharedMemorySize, seg_cap(3) * BK_BYTES));
      CUDA_CHECK(cudaFuncSetAttribute(k_merge_buckets<512, seg_cap(4)>, cudaFuncAttributeMaxDynamicSharedMemorySize, seg_cap(4) * BK_BYTES));
      if (idx->device >= 0 && idx->device < 64) attr_done[idx->device] = true;
    }
  }
  {
    // a cluster of CTAs per row while every cluster of the call is resident at once (one CTA per SM: registers),
    // single CTAs beyond that
    static const uint32_t forced = (uint32_t)env_u64("IMPGX_SMALL_CLUSTER", 0);
    uint32_t cl = 8;
    while (cl > 1 && (uint64_t)R * cl > (uint64_t)sm_count()) cl >>= 1;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) cl = forced;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(R * cl);
    cfg.blockDim = dim3(SB_THREADS);
    cfg.dynamicSmemBytes = SB_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, k_small_bfs, view, d_r, sp, c);
    if (le != cudaSuccess && cl > 1) {
      // a device partition that cannot co-schedule the cluster: the walk is the same with one CTA per row
      cudaGetLastError();
      cfg.gridDim = dim3(R);
      at[0].val.clusterDim.x = 1;
      le = cudaLaunchKernelEx(&cfg, k_small_bfs, view, d_r, sp, c);
    }
    CUDA_CHECK(le);
    ctx.launches++;
  }
  if (bed) {
    // the bucket merge of the batched path, on the lists the walk left on the device (fixed grids, strided loops)
    const int64_t d = p.merge_distance;
    const int ms = p.merge_strands ? 1 : 0;
    auto list_of = [&](int cl) { return c.lists + (uint64_t)cl * cap; };
    k_merge_tiny<<<4 * R, 128, 0, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(TINY_CLASS), 0u, d, ms, 0, c.out_cnt,
                                       c.cls + TINY_CLASS);
    k_merge_buckets<32, seg_cap(0)><<<4 * R, 256, 8 * seg_cap(0) * BK_BYTES, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(0), 0u,
                                                                                 d, ms, 0, c.out_cnt, c.cls + 0);
    k_merge_buckets<32, seg_cap(1)><<<2 * R, 256, 8 * seg_cap(1) * BK_BYTES, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(1), 0u,
                                                                                 d, ms, 0, c.out_cnt, c.cls + 1);
    k_merge_buckets<128, seg_cap(2)><<<8 * R, 128, seg_cap(2) * BK_BYTES, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(2), 0u, d,
                                                                              ms, 0, c.out_cnt, c.cls + 2);
    k_merge_buckets<128, seg_cap(3)><<<4 * R, 128, seg_cap(3) * BK_BYTES, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(3), 0u, d,
                                                                              ms, 0, c.out_cnt, c.cls + 3);
    k_merge_buckets<512, seg_cap(4)><<<2 * R, 512, seg_cap(4) * BK_BYTES, s>>>(c.boxes, c.bk_beg, c.bk_cur, list_of(4), 0u, d,
                                                                              ms, 0, c.out_cnt, c.cls + 4);
    CUDA_CHECK(cudaGetLastError());
    ctx.launches += 6;
  }
  k_small_finish<<<1, SB_THREADS, 0, s>>>(sp, c);
  CUDA_CHECK(cudaGetLastError());
  ctx.launches++;

  // header + the first rows in one copy; the rest only when there is more
  constexpr size_t SPEC = 2048;
  const size_t spec_bytes = hdr_bytes + std::min<size_t>(SPEC, cap) * sizeof(SbOut);
  size_t pcap = 0;
  char *h = (char *)pinned_acquire(spec_bytes, &pcap);
  struct Release {
    void *p;
    size_t cap;
    ~Release() { pinned_release(p, cap); }
  } rel{h, pcap};
  CUDA_CHECK(cudaMemcpyAsync(h, outbuf.get(), spec_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  ctx.d2h_bytes += spec_bytes;
  const uint32_t *hdr = reinterpret_cast<const uint32_t *>(h);
  const uint32_t worst = hdr[R + 1];
  if (worst != SB_OK) {
    if (worst == SB_OVERFLOW) {
      std::lock_guard<std::mutex> lock(idx->mu);
      idx->small_penalty = std::min<uint32_t>(std::max<uint32_t>(idx->small_penalty * 2, 1u), 256u);
      idx->small_skip = idx->small_penalty;
    }
    return false;  // SB_INVALID: the batched path words the error
  }
  {
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->small_penalty = 0;
  }
  const size_t total = hdr[R];
  std::vector<SbOut> rest;
  if (total > SPEC) {
    rest.resize(total - SPEC);
    CUDA_CHECK(cudaMemcpyAsync(rest.data(), c.out + SPEC, (total - SPEC) * sizeof(SbOut), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    ctx.d2h_bytes += (total - SPEC) * sizeof(SbOut);
  }
  for (uint32_t i = 0; i <= R; i++) res->row_off[i] = hdr[i];
  res->qid.resize(total); res->tid.resize(total);
  res->qf.resize(total); res->ql.resize(total);
  res->tf.resize(total); res->tl.resize(total);
  const SbOut *first = reinterpret_cast<const SbOut *>(h + hdr_bytes);
  for (size_t i = 0; i < total; i++) {
    const SbOut &o = i < SPEC ? first[i] : rest[i - SPEC];
    res->qid[i] = o.q_id; res->qf[i] = o.q_first; res->ql[i] = o.q_last;
    res->tid[i] = o.t_id; res->tf[i] = o.t_first; res->tl[i] = o.t_last;
  }
  ctx.stab_ranges += hdr[R + 2];
  ctx.liftovers += hdr[R + 3];
  ctx.lift_launches++;
  {
    static const bool trace = env_u64("IMPGX_TRACE", 0) >= 2;  // cycles per phase of the walk, summed over the rows
    if (trace) {
      unsigned long long cyc[SB_PHASES];
      CUDA_CHECK(cudaMemcpy(cyc, c.stats + 2, sizeof(cyc), cudaMemcpyDeviceToHost));
      static const char *names[SB_PHASES] = {"stab", "fill", "lift", "order", "results", "group", "fold", "visited", "frontier", "buckets"};
      fprintf(stderr, "[impgx] single-launch walk, kcycles:");
      for (int k = 0; k < SB_PHASES; k++) fprintf(stderr, " %s %.1f", names[k], cyc[k] / 1e3);
      fprintf(stderr, " | ranges %u hits %u rows out %zu\n", hdr[R + 2], hdr[R + 3], total);
    }
  }
  return true;
}

impgx_results *query_batch(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params &p, bool bed,
                           bool ranges_on_device, bool results_to_host, void *stream, Comm *comm) {
  REQUIRE(idx, IMPGX_E_INVALID, "index is NULL");
  REQUIRE(!comm || bed, IMPGX_E_UNSUPPORTED, "the sharded index returns BED rows only");
  REQUIRE(ranges || n == 0, IMPGX_E_INVALID, "ranges is NULL");
  REQUIRE(p.mode <= IMPGX_MODE_MULTI_DFS, IMPGX_E_INVALID, "unknown mode");
  REQUIRE(p.max_depth <= 65535, IMPGX_E_INVALID, "max_depth is a u16 in the reference");
  check_device(idx->device);
  // a call with host buffers on both sides and no caller stream runs on a stream of its own, so that concurrent
  // callers overlap; the device entry points stay on the stream they were given (0 = the legacy default stream)
  CallScratch scratch(idx, /*want_stream=*/!ranges_on_device && results_to_host && stream == nullptr);
  Arena &arena = *scratch.arena;
  cudaStream_t s = scratch.own_stream ? scratch.stream : (cudaStream_t)stream;
  auto t0 = std::chrono::steady_clock::now();

  std::unique_ptr<impgx_results> res(new impgx_results());
  res->device = idx->device;
  res->n_rows = n;
  res->has_cigar = p.store_cigar != 0;
  res->row_off.assign(n + 1, 0);
  if (res->has_cigar) res->cig_off.push_back(0);

  // device-resident mode: per-chunk copies of the output columns (stream-ordered
  // allocations, a few MB each) until they are concatenated at the end
  struct DevChunk {
    uint32_t *qid = nullptr, *tid = nullptr;
    int32_t *qf = nullptr, *ql = nullptr, *tf = nullptr, *tl = nullptr;
    uint64_t n = 0;
  };
  std::vector<DevChunk> dev_chunks;
  auto free_chunks = [&]() {
    for (auto &c : dev_chunks) {
      cudaFreeAsync(c.qid, s); cudaFreeAsync(c.tid, s); cudaFreeAsync(c.qf, s);
      cudaFreeAsync(c.ql, s); cudaFreeAsync(c.tf, s); cudaFreeAsync(c.tl, s);
    }
    dev_chunks.clear();
  };

  Ctx total;
  tl_small_scans = 0;
  // Row batches are sized so that one batch lifts about IMPGX_HITS_PER_BATCH hits
  // (scratch memory is proportional to the hits in flight); the hits-per-row
  // figure is learned from the first, small batch and kept on the index.
  const size_t fixed_chunk = (size_t)env_u64("IMPGX_ROWS_PER_BATCH", 0);
  // Scratch is ~45 bytes per liftover in flight. Larger batches amortise the per-batch launches and size readbacks
  // (C4: 2.09 s per step at 2*10^8 hits per batch, 2.02 s at 10^9), so a batch aims at as many hits as half of the
  // memory the index leaves free will hold, 10^9 at most; the ranks of a sharded index agree on the smallest figure.
  double target_hits = (double)env_u64("IMPGX_HITS_PER_BATCH", 0);
  auto size_batches = [&]() {  // on the first batch of the batched path only: calls of a few rows never get here
    if (target_hits > 0) return;
    size_t free_b = 0, total_b = 0;
    CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    const double left = (double)total_b - (double)idx->device_bytes - 8e9;
    uint64_t t = (uint64_t)std::min(std::max(left / 2 / 48, 5e7), 1e9);
    if (comm) {
      std::vector<uint64_t> all((size_t)comm->size());
      comm->allgather_u64(&t, 1, all.data(), s);
      t = *std::min_element(all.begin(), all.end());
    }
    target_hits = (double)t;
  };
  double hits_per_row;
  {
    std::lock_guard<std::mutex> lock(idx->mu);
    hits_per_row = idx->hits_per_row;
  }
  auto next_chunk = [&]() -> size_t {
    if (fixed_chunk) return fixed_chunk;
    size_batches();
    if (hits_per_row <= 0) return 512;
    double c = target_hits / hits_per_row;
    // the direct BED path keeps dense tables over rows x sequences (bucket_kernels.cuh)
    const double table_rows = (double)(1ull << 27) / (double)std::max<uint32_t>(idx->n_seqs, 1u);
    if (bed && c > table_rows) c = table_rows;
    if (c < 64) c = 64;
    if (c > 4194304) c = 4194304;
    return (size_t)c;
  };
  size_t done = 0;
  uint64_t res_base = 0, cig_base = 0;
  try {
    if (small_eligible(idx, n, p, bed, results_to_host, comm)) {
      REQUIRE(!(bed && p.store_cigar), IMPGX_E_INVALID, "BED output carries no CIGAR (src/main.rs:7447)");
      if (try_small(idx, ranges, n, p, bed, ranges_on_device, s, arena, res.get(), total)) {
        done = n;
        res_base = res->qid.size();
      }
      arena.reset();
    }
    while (done < n) {
      const size_t m = std::min(next_chunk(), n - done);
      uint64_t chunk_hits = 0;
      {
        const impgx_range *d_r;
        DBuf<impgx_range> d_stage;
        if (ranges_on_device) {
          d_r = ranges + done;
        } else {
          d_stage.alloc(m, arena);
          CUDA_CHECK(cudaMemcpyAsync(d_stage.get(), ranges + done, m * sizeof(impgx_range), cudaMemcpyHostToDevice, s));
          total.h2d_bytes += m * sizeof(impgx_range);
          d_r = d_stage.get();
        }
        BatchOut bo;
        Runner runner(idx, p, s, arena, comm);
        if (comm) runner.run_sharded(d_r, (uint32_t)m, bo);
        else runner.run(d_r, (uint32_t)m, bed, bo);
        const Ctx &c = runner.ctx;
        total.launches += c.launches; total.lift_launches += c.lift_launches;
        total.stab_ranges += c.stab_ranges; total.liftovers += c.liftovers;
        chunk_hits = c.liftovers;
        total.lift_runs += c.lift_runs; total.lift_bytes += c.lift_bytes;
        total.lift_touched += c.lift_touched; total.lift_rov += c.lift_rov;
        total.h2d_bytes += c.h2d_bytes; total.d2h_bytes += c.d2h_bytes;
        total.lift_ms += c.lift_ms; total.stab_ms += c.stab_ms; total.fold_ms += c.fold_ms; total.merge_ms += c.merge_ms;
        total.exch_ms += c.exch_ms;
        total.merge_boxes += c.merge_boxes;
        total.merge_kernel_ms += c.merge_kernel_ms;
        total.w_stab += c.w_stab; total.w_lift += c.w_lift; total.w_order += c.w_order; total.w_fold += c.w_fold;
        total.w_assemble += c.w_assemble; total.w_merge += c.w_merge;
        WallTimer wcopy(total.w_copy);

        const uint64_t R = bo.n_results;
        // chunk-relative row offsets straight into the (pinned) result column; rebased after the sync
        CUDA_CHECK(cudaMemcpyAsync(res->row_off.data() + done + 1, bo.row_off.get() + 1, m * 8, cudaMemcpyDeviceToHost, s));
        total.d2h_bytes += (m + 1) * 8;
        if (results_to_host) {
          const size_t old = res->qid.size();
          if (done + m < n) {
            // more batches follow: room for all of them at the rate seen so far, so that the columns are not
            // re-allocated and copied batch after batch
            const double per_row = (double)(old + R) / (double)(done + m);
            const size_t want = (size_t)(per_row * (double)n * 1.05) + 4096;
            res->qid.reserve(want); res->tid.reserve(want);
            res->qf.reserve(want); res->ql.reserve(want);
            res->tf.reserve(want); res->tl.reserve(want);
          }
          res->qid.resize(old + R); res->tid.resize(old + R);
          res->qf.resize(old + R); res->ql.resize(old + R);
          res->tf.resize(old + R); res->tl.resize(old + R);
          if (R) {
            CUDA_CHECK(cudaMemcpyAsync(res->qid.data() + old, bo.q_id.get(), R * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(res->qf.data() + old, bo.q_first.get(), R * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(res->ql.data() + old, bo.q_last.get(), R * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(res->tid.data() + old, bo.t_id.get(), R * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(res->tf.data() + old, bo.t_first.get(), R * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(res->tl.data() + old, bo.t_last.get(), R * 4, cudaMemcpyDeviceToHost, s));
          }
          total.d2h_bytes += R * 24;
          if (res->has_cigar) {
            std::vector<uint64_t> co(R + 1);
            CUDA_CHECK(cudaMemcpyAsync(co.data(), bo.cig_off.get(), (R + 1) * 8, cudaMemcpyDeviceToHost, s));
            const size_t oldc = res->cig.size();
            res->cig.resize(oldc + bo.n_cig);
            if (bo.n_cig)
              CUDA_CHECK(cudaMemcpyAsync(res->cig.data() + oldc, bo.cig.get(), bo.n_cig * 4, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            {
              const size_t o0 = res->cig_off.size();
              res->cig_off.resize(o0 + R);
              for (uint64_t i = 1; i <= R; i++) res->cig_off[o0 + i - 1] = cig_base + co[i];
            }
            total.d2h_bytes += (R + 1) * 8 + bo.n_cig * 4;
            cig_base += bo.n_cig;
          }
        } else {
          DevChunk dc;
          dc.n = R;
          const size_t rb = std::max<uint64_t>(R, 1) * 4;
          CUDA_CHECK(cudaMallocAsync((void **)&dc.qid, rb, s));
          CUDA_CHECK(cudaMallocAsync((void **)&dc.tid, rb, s));
          CUDA_CHECK(cudaMallocAsync((void **)&dc.qf, rb, s));
          CUDA_CHECK(cudaMallocAsync((void **)&dc.ql, rb, s));
          CUDA_CHECK(cudaMallocAsync((void **)&dc.tf, rb, s));
          CUDA_CHECK(cudaMallocAsync((void **)&dc.tl, rb, s));
          dev_chunks.push_back(dc);
          if (R) {
            CUDA_CHECK(cudaMemcpyAsync(dc.qid, bo.q_id.get(), R * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(dc.qf, bo.q_first.get(), R * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(dc.ql, bo.q_last.get(), R * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(dc.tid, bo.t_id.get(), R * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(dc.tf, bo.t_first.get(), R * 4, cudaMemcpyDeviceToDevice, s));
            CUDA_CHECK(cudaMemcpyAsync(dc.tl, bo.t_last.get(), R * 4, cudaMemcpyDeviceToDevice, s));
          }
        }
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (res_base)
          for (size_t i = 1; i <= m; i++) res->row_off[done + i] += res_base;
        res_base += R;
      }
      // every arena block of this batch is out of scope and the stream is idle
      arena.reset();
      if (comm) {
        // every rank must cut the same row batches: size them by the busiest rank
        std::vector<uint64_t> all((size_t)comm->size());
        comm->allgather_u64(&chunk_hits, 1, all.data(), s);
        chunk_hits = *std::max_element(all.begin(), all.end());
      }
      hits_per_row = std::max(hits_per_row * 0.9, (double)chunk_hits / (double)m + 1.0);
      {
        std::lock_guard<std::mutex> lock(idx->mu);
        idx->hits_per_row = hits_per_row;
      }
      done += m;
    }
  } catch (...) {
    if (comm) comm->abort();  // release peers blocked in an exchange
    cudaStreamSynchronize(s);
    free_chunks();
    arena.reset();
    throw;
  }
  res->n_results = res_base;
  res->n_cig = cig_base;

  if (!results_to_host) {
    // concatenate the chunk outputs into one set of device columns
    res->on_device = true;
    res->stream = s;
    const uint64_t R = res_base;
    const size_t rb = std::max<uint64_t>(R, 1) * 4;
    CUDA_CHECK(cudaMallocAsync((void **)&res->d_row_off, (n + 1) * 8, s));
    CUDA_CHECK(cudaMemcpyAsync(res->d_row_off, res->row_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    if (dev_chunks.size() == 1) {
      DevChunk &c = dev_chunks[0];
      res->d_qid = c.qid; res->d_tid = c.tid; res->d_qf = c.qf; res->d_ql = c.ql; res->d_tf = c.tf; res->d_tl = c.tl;
      dev_chunks.clear();
    } else {
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_qid, rb, s));
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_qf, rb, s));
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_ql, rb, s));
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_tid, rb, s));
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_tf, rb, s));
      CUDA_CHECK(cudaMallocAsync((void **)&res->d_tl, rb, s));
      uint64_t off = 0;
      for (auto &c : dev_chunks) {
        const uint64_t r = c.n;
        if (r) {
          CUDA_CHECK(cudaMemcpyAsync(res->d_qid + off, c.qid, r * 4, cudaMemcpyDeviceToDevice, s));
          CUDA_CHECK(cudaMemcpyAsync(res->d_qf + off, c.qf, r * 4, cudaMemcpyDeviceToDevice, s));
          CUDA_CHECK(cudaMemcpyAsync(res->d_ql + off, c.ql, r * 4, cudaMemcpyDeviceToDevice, s));
          CUDA_CHECK(cudaMemcpyAsync(res->d_tid + off, c.tid, r * 4, cudaMemcpyDeviceToDevice, s));
          CUDA_CHECK(cudaMemcpyAsync(res->d_tf + off, c.tf, r * 4, cudaMemcpyDeviceToDevice, s));
          CUDA_CHECK(cudaMemcpyAsync(res->d_tl + off, c.tl, r * 4, cudaMemcpyDeviceToDevice, s));
        }
        off += r;
      }
      free_chunks();
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
  }

  auto t1 = std::chrono::steady_clock::now();
  impgx_stats st = impgx_stats{};
  st.kernel_launches = total.launches - std::min(total.launches, tl_small_scans);
  st.stab_ranges = total.stab_ranges;
  st.liftovers = total.liftovers;
  st.lift_runs = total.lift_runs;
  st.lift_bytes = total.lift_bytes;
  st.lift_touched_bytes = total.lift_touched;
  st.lift_window_runs = total.lift_rov;
  st.lift_launches = total.lift_launches;
  st.results = res_base;
  st.merged = bed ? res_base : 0;
  st.h2d_bytes = total.h2d_bytes;
  st.d2h_bytes = total.d2h_bytes;
  st.lift_ms = total.lift_ms;
  st.stab_ms = total.stab_ms;
  st.fold_ms = total.fold_ms;
  st.merge_ms = total.merge_ms;
  st.exchange_ms = total.exch_ms;
  st.merge_kernel_ms = total.merge_kernel_ms;
  st.merge_boxes = total.merge_boxes;
  st.total_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();
  if (getenv("IMPGX_TRACE"))
    fprintf(stderr,
            "[impgx] rows=%zu total=%.2f ms | wall: stab %.2f lift %.2f order %.2f fold %.2f assemble %.2f merge %.2f "
            "copy %.2f | dev: stab %.2f lift %.2f fold %.2f merge %.2f | launches %llu | arena %.2f GB peak %.2f GB\n",
            n, st.total_ms, total.w_stab, total.w_lift, total.w_order, total.w_fold, total.w_assemble, total.w_merge,
            total.w_copy, st.stab_ms, st.lift_ms, st.fold_ms, st.merge_ms, (unsigned long long)st.kernel_launches,
            arena.capacity() / 1e9, arena.peak() / 1e9);
  {
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->last = st;
  }
  return res.release();
}

// Number of tree entries the closed stab of [start, end] on `target` visits (coitrees' closed-interval test,
// src/impg.rs:1940): what populate_cigar_cache would cache (refine.cu).
uint64_t stab_count_closed(impgx_index *idx, uint32_t target, int32_t start, int32_t end) {
  check_device(idx->device);
  CallScratch scratch(idx, /*want_stream=*/true);
  Arena &arena = *scratch.arena;
  Ctx ctx;
  cudaStream_t s = scratch.stream;
  DBuf<Frontier> fr(1, arena);
  DBuf<Window> win(1, arena);
  DBuf<uint32_t> cnt(1, arena);
  const Frontier f{0u, target, start, end};
  CUDA_CHECK(cudaMemcpyAsync(fr.get(), &f, sizeof(f), cudaMemcpyHostToDevice, s));
  auto kern = k_stab_count<true, false>;
  LAUNCH(kern, 1, 32, s, idx->view(), fr.get(), (uint64_t)1, win.get(), cnt.get(), (uint32_t *)nullptr);
  uint32_t c = 0;
  CUDA_CHECK(cudaMemcpyAsync(&c, cnt.get(), 4, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  fr.release(); win.release(); cnt.release();
  return c;
}

// ============================================================ test hooks for the device primitives
// (host arrays in and out; `force` = 0 automatic, 1 the single-CTA kernels, 2 the CUB pipelines)
void debug_sort_pairs(int device, uint64_t *keys, uint32_t *vals, uint64_t n, int begin_bit, int end_bit, int key_bytes) {
  check_device(device);
  Arena ar;
  Scratch sc{&ar};
  Ctx ctx;
  cudaStream_t s = nullptr;
  DBuf<uint32_t> dv(n, ar);
  CUDA_CHECK(cudaMemcpy(dv.get(), vals, n * 4, cudaMemcpyHostToDevice));
  if (key_bytes == 8) {
    DBuf<uint64_t> dk(n, ar);
    CUDA_CHECK(cudaMemcpy(dk.get(), keys, n * 8, cudaMemcpyHostToDevice));
    sort_pairs(dk, dv, n, begin_bit, end_bit, sc, s, ctx);
    CUDA_CHECK(cudaMemcpy(keys, dk.get(), n * 8, cudaMemcpyDeviceToHost));
  } else {
    std::vector<uint32_t> k32(n);
    for (uint64_t i = 0; i < n; i++) k32[i] = (uint32_t)keys[i];
    DBuf<uint32_t> dk(n, ar);
    CUDA_CHECK(cudaMemcpy(dk.get(), k32.data(), n * 4, cudaMemcpyHostToDevice));
    sort_pairs(dk, dv, n, begin_bit, end_bit, sc, s, ctx);
    CUDA_CHECK(cudaMemcpy(k32.data(), dk.get(), n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < n; i++) keys[i] = k32[i];
  }
  CUDA_CHECK(cudaMemcpy(vals, dv.get(), n * 4, cudaMemcpyDeviceToHost));
}

void debug_exclusive_scan(int device, uint64_t *a, uint64_t n_plus_1) {
  check_device(device);
  Arena ar;
  Scratch sc{&ar};
  cudaStream_t s = nullptr;
  DBuf<uint64_t> d(n_plus_1, ar);
  CUDA_CHECK(cudaMemcpy(d.get(), a, n_plus_1 * 8, cudaMemcpyHostToDevice));
  exclusive_scan_u64(d.get(), n_plus_1, sc, s);
  CUDA_CHECK(cudaMemcpy(a, d.get(), n_plus_1 * 8, cudaMemcpyDeviceToHost));
}

// ============================================================ KAT surface
void project_batch(int device, size_t n, const int32_t *req_start, const int32_t *req_end, const impgx_record *records,
                   const uint32_t *runs, const uint64_t *run_offsets, int32_t *out4, uint8_t *ok,
                   uint64_t *out_run_offsets, uint32_t *out_runs, size_t out_runs_cap) {
  check_device(device);
  REQUIRE(req_start && req_end && records && run_offsets && out4 && ok, IMPGX_E_INVALID, "NULL argument");
  if (n == 0) return;
  cudaStream_t s = nullptr;
  Ctx ctx;
  Arena arena;
  std::vector<EntryRec> recs(n);
  std::vector<uint32_t> blk_off(n + 1);
  std::vector<Frontier> fr(n);
  std::vector<LiftTask> tasks(n);
  uint64_t blocks = 0;
  for (size_t i = 0; i < n; i++) {
    const impgx_record &r = records[i];
    uint64_t nr = run_offsets[i + 1] - run_offsets[i];
    REQUIRE(nr < (1ull << 30), IMPGX_E_INVALID, "too many runs");
    blk_off[i] = (uint32_t)blocks;
    EntryRec e;
    e.t_start = r.target_start; e.t_end = r.target_end;
    e.q_start = r.query_start; e.q_end = r.query_end;
    e.query_id = r.query_id;
    e.nruns_flags = ((uint32_t)nr << 2) | (r.strand ? FLAG_STRAND : 0u) | (r.reserved & 1u ? FLAG_REVERSED : 0u);
    e.aln_off = (uint32_t)blocks;
    e.vrank = 0;
    recs[i] = e;
    blocks += aln_sectors((uint32_t)nr);
    fr[i] = Frontier{(uint32_t)i, r.target_id, req_start[i], req_end[i]};
    tasks[i] = LiftTask{(uint32_t)i, (uint32_t)i};
  }
  blk_off[n] = (uint32_t)blocks;
  const uint64_t total_runs = run_offsets[n];
  DBuf<uint32_t> d_raw(std::max<uint64_t>(total_runs, 1), arena), d_blk(n + 1, arena), d_stream(std::max<uint64_t>(blocks, 1) * 8, arena);
  DBuf<uint64_t> d_off(n + 1, arena);
  DBuf<EntryRec> d_rec(n, arena);
  DBuf<Frontier> d_fr(n, arena);
  DBuf<LiftTask> d_tasks(n, arena);
  DBuf<Hit> d_hits(n, arena);
  DBuf<CigarSlice> d_slices(n, arena);
  if (total_runs) CUDA_CHECK(cudaMemcpyAsync(d_raw.get(), runs, total_runs * 4, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(d_off.get(), run_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(d_blk.get(), blk_off.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(d_rec.get(), recs.data(), n * sizeof(EntryRec), cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(d_fr.get(), fr.data(), n * sizeof(Frontier), cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(d_tasks.get(), tasks.data(), n * sizeof(LiftTask), cudaMemcpyHostToDevice, s));
  LAUNCH(k_build_blocks, grid_warps(n), 256, s, d_raw.get(), d_off.get(), d_blk.get(), (uint64_t)n, d_stream.get(),
         (int *)nullptr);
  DevIndexView ix{};
  ix.e_rec = d_rec.get();
  ix.stream = d_stream.get();
  LiftParams lp{};
  lp.clip = 0;
  lp.min_output_len = -1;
  if (!out_runs) {
    CUDA_CHECK(cudaMemsetAsync(d_slices.get(), 0, n * sizeof(CigarSlice), s));
    LAUNCH((k_liftover_ends<false, false, 3>), grid_threads(n), 256, s, ix, d_fr.get(), d_tasks.get(), (uint64_t)n, lp,
           d_hits.get(), (unsigned long long *)nullptr, BucketOut{});
  } else {
    LAUNCH(k_liftover, grid_warps(n), 256, s, ix, d_fr.get(), d_tasks.get(), (uint64_t)n, lp, d_hits.get(),
           d_slices.get(), (unsigned long long *)nullptr);
  }
  std::vector<Hit> hits(n);
  std::vector<CigarSlice> slices(n);
  CUDA_CHECK(cudaMemcpyAsync(hits.data(), d_hits.get(), n * sizeof(Hit), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaMemcpyAsync(slices.data(), d_slices.get(), n * sizeof(CigarSlice), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  std::vector<uint64_t> oo(n + 1, 0);
  for (size_t i = 0; i < n; i++) {
    ok[i] = hits[i].row != INVALID_ID;
    out4[4 * i + 0] = hits[i].q_first;
    out4[4 * i + 1] = hits[i].q_last;
    out4[4 * i + 2] = hits[i].t_first;
    out4[4 * i + 3] = hits[i].t_last;
    oo[i + 1] = oo[i] + slices[i].n_ops;
  }
  if (out_run_offsets) memcpy(out_run_offsets, oo.data(), (n + 1) * 8);
  if (out_runs && oo[n]) {
    REQUIRE(oo[n] <= out_runs_cap, IMPGX_E_INVALID, "out_runs capacity too small");
    DBuf<uint64_t> d_oo(n + 1, arena);
    DBuf<uint32_t> d_out(oo[n], arena);
    CUDA_CHECK(cudaMemcpyAsync(d_oo.get(), oo.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
    LAUNCH(k_emit_cigar, grid_warps(n), 256, s, ix, d_tasks.get(), d_slices.get(), d_oo.get(), (uint64_t)n, d_out.get());
    CUDA_CHECK(cudaMemcpyAsync(out_runs, d_out.get(), oo[n] * 4, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
  }
}

}  // namespace impgx
