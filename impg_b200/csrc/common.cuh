// common.cuh — error handling, device buffers and shared layouts of libimpgx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
constexpr int RUNS_PER_BLOCK = 32;  // K: CIGAR runs per 128-byte block / per checkpoint
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
  uint32_t blk_off;          // first 32-run block of the alignment in the run stream
  uint32_t ck_off;           // first checkpoint (blk_off + alignment ordinal)
};
static_assert(sizeof(EntryRec) == 32, "EntryRec must be one sector");

// Cumulative (target, query) bases consumed before a run block, relative to
// the alignment's first run; one extra per alignment holds the totals.
struct __align__(8) Checkpoint {
  uint32_t t_off, q_off;
};

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

struct DevIndexView {
  // entry columns, sorted by (target, start), stable in PAF order
  const int32_t *e_start;
  const int32_t *e_end;
  const int32_t *e_pmax;    // running max of e_end within the target
  const uint32_t *e_vrank;  // rank in the coitrees visit order of the target's tree
  const EntryRec *e_rec;
  const uint64_t *tgt_off;  // n_seqs + 1
  const int32_t *seq_len;   // n_seqs
  const Checkpoint *ck;
  const uint32_t *runs;     // padded to RUNS_PER_BLOCK per alignment
  uint32_t n_seqs;
  uint64_t n_entries;
};

// ----------------------------------------------------------------- buffers
// Stream-ordered device buffer (cudaMallocAsync pool).
template <class T>
struct DBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DBuf() {}
  DBuf(size_t count, cudaStream_t stream) { alloc(count, stream); }
  DBuf(const DBuf &) = delete;
  DBuf &operator=(const DBuf &) = delete;
  DBuf(DBuf &&o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DBuf &operator=(DBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count, cudaStream_t stream) {
    release();
    s = stream;
    n = count;
    if (count) CUDA_CHECK(cudaMallocAsync((void **)&p, count * sizeof(T), stream));
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
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
