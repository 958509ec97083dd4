// common.cuh — error handling, device buffers and shared layouts of libimpgx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/impgx.h"

namespace impgx {

// ----------------------------------------------------------------- errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

#define CUDA_CHECK(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      throw ::impgx::Error(_e == cudaErrorMemoryAllocation ? IMPGX_E_NOMEM : IMPGX_E_CUDA,       \
                           std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                               ":" + std::to_string(__LINE__) + ")");                            \
    }                                                                                            \
  } while (0)

#define REQUIRE(cond, code, msg)                       \
  do {                                                 \
    if (!(cond)) throw ::impgx::Error((code), (msg)); \
  } while (0)

// ----------------------------------------------------------------- layout
constexpr int RUNS_PER_BLOCK = 8;   // K: CIGAR runs per 32-byte block (one DRAM sector) = per checkpoint
constexpr uint32_t FLAG_STRAND = 1u;    // '-' strand (bit 63 of strand_and_data_offset)
constexpr uint32_t FLAG_REVERSED = 2u;  // reversed (bidirectional) entry (bit 62)
constexpr uint32_t INVALID_ID = 0xffffffffu;

// One interval-tree entry (reference QueryMetadata, src/impg.rs:164-174, plus
// the node's first/last). 32 bytes = one DRAM sector.
struct __align__(32) EntryRec {
  int32_t t_start, t_end;    // interval on the indexed (target) sequence
  int32_t q_start, q_end;    // interval on the other sequence
  uint32_t query_id;         // the other sequence
  uint32_t nruns_flags;      // n_runs << 2 | FLAG_REVERSED | FLAG_STRAND
  uint32_t aln_off;          // the alignment's region in the stream, in 32-byte sectors
  uint32_t vrank;            // rank of this entry in the coitrees visit order of its target
};
static_assert(sizeof(EntryRec) == 32, "EntryRec must be one sector");

// Cumulative (target, query) bases consumed before a run block, relative to
// the alignment's first run; one extra per alignment holds the totals.
struct __align__(8) Checkpoint {
  uint32_t t_off, q_off;
};

// Stream layout of one alignment (all in 32-byte sectors, so every block is
// one DRAM sector and checkpoints sit next to the runs they index):
//   [ checkpoints 0..nblk (nblk+1 of them, padded to a multiple of 4) ][ nblk run blocks ]
#if defined(__CUDACC__)
#define IMPGX_HD __host__ __device__ __forceinline__
#else
#define IMPGX_HD inline
#endif
IMPGX_HD uint32_t aln_nblk(uint32_t n_runs) { return (n_runs + RUNS_PER_BLOCK - 1) / RUNS_PER_BLOCK; }
IMPGX_HD uint32_t aln_ck_sectors(uint32_t nblk) { return (nblk + 1 + 3) / 4; }
IMPGX_HD uint32_t aln_sectors(uint32_t n_runs) { return aln_ck_sectors(aln_nblk(n_runs)) + aln_nblk(n_runs); }
IMPGX_HD const Checkpoint *aln_ck(const uint32_t *stream, uint32_t aln_off) {
  return reinterpret_cast<const Checkpoint *>(stream + (uint64_t)aln_off * 8);
}
IMPGX_HD const uint32_t *aln_runs(const uint32_t *stream, uint32_t aln_off, uint32_t nblk) {
  return stream + ((uint64_t)aln_off + aln_ck_sectors(nblk)) * 8;
}

// One stab hit to lift (output of the stab kernel, input of the liftover kernel).
struct __align__(8) LiftTask {
  uint32_t entry;  // index into the entry columns
  uint32_t range;  // index into the frontier of this hop
};

// A frontier range of one hop: stab `seq` with [start,end) on behalf of `row`.
struct __align__(16) Frontier {
  uint32_t row;
  uint32_t seq;
  int32_t start, end;
};

// One lifted hit (AdjustedInterval without CIGAR), 32 bytes.
struct __align__(32) Hit {
  uint32_t row;              // batch row, INVALID_ID if the liftover returned None / was filtered
  uint32_t q_id;
  int32_t q_first, q_last;   // q_first > q_last on the reverse strand
  uint32_t t_id;
  int32_t t_first, t_last;
  uint32_t vrank;            // coitrees visit rank of the entry within its target
};
static_assert(sizeof(Hit) == 32, "Hit must be one sector");

// One result of the direct BED path as it sits in its (row, query sequence) bucket: the liftover
// epilogue writes it, the segment merge reads it once. Row and query sequence are implied by the
// bucket. 32 bytes = one sector.
struct __align__(32) BoxRec {
  uint64_t ord;              // position in the reference's result order of the row (merge_kernels.cuh, make_ord)
  int32_t q_first, q_last;   // q_first > q_last on the reverse strand
  uint32_t t_id;
  int32_t t_first, t_last;
  uint32_t flags;            // BOX_MERGED_A: a stage-A result already (merged on the rank that produced it)
};
constexpr uint32_t BOX_MERGED_A = 1u;
static_assert(sizeof(BoxRec) == 32, "BoxRec must be one sector");

// One merged BED row of a bucket, staged at the bucket's first slots until the rows are compacted.
struct __align__(4) SegOut {
  int32_t q_first, q_last;
  uint32_t t_id;
  int32_t t_first, t_last;
};

struct DevIndexView {
  // entry columns, sorted by (target, start), stable in PAF order
  const int32_t *e_start;
  const int32_t *e_end;
  const int32_t *e_pmax;    // running max of e_end within the target
  const EntryRec *e_rec;
  const uint32_t *e_qid;    // e_rec[i].query_id as a column of its own (bucket counting in the stab scan)
  const uint64_t *tgt_off;  // n_seqs + 1
  const int32_t *seq_len;   // n_seqs
  const uint32_t *stream;   // per alignment: checkpoints, then 8-run blocks (see aln_* helpers)
  uint32_t n_seqs;
  uint64_t n_entries;
};

// ----------------------------------------------------------------- buffers
// Device scratch arena. The query pipeline allocates hundreds of temporaries
// per batch; cudaMallocAsync showed sporadic 100-800 ms stalls when the pool
// remapped memory for multi-GB requests, so scratch comes from slabs owned by
// the index: bump allocation with stack discipline (a freed block is reclaimed
// as soon as everything above it is freed — C++ scopes make that the common
// case), whole-arena reset at the end of a batch. All work of a batch is
// issued on one stream, so reuse in program order is safe.
class Arena {
 public:
  Arena() {}
  Arena(const Arena &) = delete;
  Arena &operator=(const Arena &) = delete;
  ~Arena() {
    for (auto &sl : slabs_) cudaFree(sl.base);
  }
  void *alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (slabs_.empty() || top_ + bytes > slabs_.back().size) grow(bytes);
    Block b{slabs_.size() - 1, top_, bytes, false};
    blocks_.push_back(b);
    top_ += bytes;
    used_ += bytes;
    if (used_ > peak_) peak_ = used_;
    return slabs_.back().base + b.off;
  }
  void free(void *p) {
    if (!p) return;
    for (size_t i = blocks_.size(); i-- > 0;) {
      Block &b = blocks_[i];
      if (slabs_[b.slab].base + b.off == (char *)p) {
        b.freed = true;
        used_ -= b.size;
        break;
      }
    }
    while (!blocks_.empty() && blocks_.back().freed) {
      const Block &b = blocks_.back();
      if (b.slab == slabs_.size() - 1) top_ = b.off;
      blocks_.pop_back();
    }
  }
  // Call with the stream idle. Consolidates the slabs so the next batch of the
  // same shape needs no cudaMalloc.
  void reset() {
    blocks_.clear();
    top_ = 0;
    used_ = 0;
    if (slabs_.size() > 1) {
      size_t total = 0;
      for (auto &sl : slabs_) {
        total += sl.size;
        cudaFree(sl.base);
      }
      slabs_.clear();
      char *p = nullptr;
      if (cudaMalloc((void **)&p, total) == cudaSuccess) slabs_.push_back(Slab{p, total});
      else cudaGetLastError();
    }
  }
  size_t capacity() const {
    size_t t = 0;
    for (auto &sl : slabs_) t += sl.size;
    return t;
  }
  size_t peak() const { return peak_; }

 private:
  struct Slab {
    char *base;
    size_t size;
  };
  struct Block {
    size_t slab, off, size;
    bool freed;
  };
  std::vector<Slab> slabs_;
  std::vector<Block> blocks_;
  size_t top_ = 0, used_ = 0, peak_ = 0;

  void grow(size_t need) {
    size_t sz = std::max<size_t>(need, std::max<size_t>(capacity(), (size_t)256 << 20));
    char *p = nullptr;
    cudaError_t e = cudaMalloc((void **)&p, sz);
    if (e != cudaSuccess && sz > need) {
      cudaGetLastError();
      sz = need;
      e = cudaMalloc((void **)&p, sz);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      throw Error(IMPGX_E_NOMEM, "device scratch allocation of " + std::to_string(sz) + " bytes failed");
    }
    slabs_.push_back(Slab{p, sz});
    top_ = 0;
  }
};

// Typed view of an arena block; frees on scope exit.
template <class T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  Arena *a = nullptr;
  DBuf() {}
  DBuf(size_t count, Arena &arena) { alloc(count, arena); }
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), a(o.a) { o.p = nullptr; o.n = 0; }
  DBuf &operator=(DBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; a = o.a;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count, Arena &arena) {
    release();
    a = &arena;
    n = count;
    p = (T *)arena.alloc(std::max<size_t>(count, 1) * sizeof(T));
  }
  void release() {
    if (p && a) a->free(p);
    p = nullptr;
    n = 0;
  }
  T *get() const { return p; }
  size_t bytes() const { return n * sizeof(T); }
};

template <class T>
struct PinnedBuf {
  T *p = nullptr;
  size_t n = 0;
  PinnedBuf() {}
  explicit PinnedBuf(size_t count) { alloc(count); }
  PinnedBuf(const PinnedBuf &) = delete;
  PinnedBuf &operator=(const PinnedBuf &) = delete;
  ~PinnedBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) CUDA_CHECK(cudaMallocHost((void **)&p, count * sizeof(T)));
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
  }
};

inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace impgx
