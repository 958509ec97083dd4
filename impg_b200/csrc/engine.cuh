// engine.cuh — shared declarations between engine.cu and api.cu.
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "index.cuh"

namespace impgx {

// per-call counters (become impgx_stats)
struct Ctx {
  uint64_t launches = 0, lift_launches = 0;
  uint64_t stab_ranges = 0, stab_candidates = 0, liftovers = 0, lift_runs = 0, lift_bytes = 0;
  uint64_t lift_touched = 0, lift_rov = 0;
  uint64_t merge_boxes = 0;   // boxes that went through the BED merge
  float merge_kernel_ms = 0;  // device time of the segment-merge launches alone
  uint64_t h2d_bytes = 0, d2h_bytes = 0;
  float lift_ms = 0, stab_ms = 0, fold_ms = 0, merge_ms = 0, exch_ms = 0;
  // host wall-clock per phase (IMPGX_TRACE=1 prints them)
  double w_stab = 0, w_lift = 0, w_order = 0, w_fold = 0, w_assemble = 0, w_merge = 0, w_copy = 0;
};

const char *last_error();
void check_device(int device);
impgx_index *index_build(const impgx_record *recs, size_t n, const uint32_t *runs, const uint64_t *run_offsets,
                         const uint64_t *seq_lens, uint32_t n_seqs, bool bidirectional, int device,
                         const uint32_t *owner = nullptr, uint32_t rank = 0, uint32_t n_ranks = 1);

// Runs the whole batch (chunked into row batches), ranges on host or device.
impgx_results *query_batch(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params &p, bool bed,
                           bool ranges_on_device, bool results_to_host, void *stream, class Comm *comm = nullptr);

uint64_t stab_count_closed(impgx_index *idx, uint32_t target, int32_t start, int32_t end);

// One or more parsed PAF files over a shared SequenceIndex (ids by first appearance, src/seqidx.rs:22-35).
struct PafData {
  std::vector<impgx_record> recs;
  std::vector<uint32_t> runs;
  std::vector<uint64_t> run_off{0};
  std::vector<std::string> names;
  std::vector<uint64_t> lens;
  std::unordered_map<std::string, uint32_t> ids;
  // where the reference would find the CIGAR text again (AlignmentRecord.strand_and_data_offset / data_bytes)
  std::vector<uint64_t> cg_off, cg_len;
  std::vector<uint32_t> file_idx;
  uint32_t n_files = 0;
};
void parse_paf(const std::string &path, PafData &out);
long parse_cigar(const char *s, size_t n, std::vector<uint32_t> &out);

// transform_coordinates_to_original (src/main.rs:4642-4678): "base:start-end" names are reported as `base` with
// the subsequence start added to both coordinates. Returns false (nothing changed) for any other name.
bool to_original_coordinates(const std::string &seq_name, std::string &base, uint32_t &offset);

std::string debug_format_rows(const char *const *names, const uint64_t *lens, uint32_t n_seqs, size_t n, const uint32_t *q_id,
                              const int32_t *q_first, const int32_t *q_last, const uint32_t *t_id, const int32_t *t_first,
                              const int32_t *t_last, const uint64_t *cig_off, const uint32_t *cig, const char *name, int32_t d,
                              int format, bool original_coordinates);
void debug_sort_pairs(int device, uint64_t *keys, uint32_t *vals, uint64_t n, int begin_bit, int end_bit, int key_bytes);
void debug_exclusive_scan(int device, uint64_t *a, uint64_t n_plus_1);

void project_batch(int device, size_t n, const int32_t *req_start, const int32_t *req_end, const impgx_record *records,
                   const uint32_t *runs, const uint64_t *run_offsets, int32_t *out4, uint8_t *ok,
                   uint64_t *out_run_offsets, uint32_t *out_runs, size_t out_runs_cap);

}  // namespace impgx

namespace impgx {
// Pinned host memory recycled across calls (cudaMallocHost is expensive, result
// columns are tens of MB per batch and device->host copies into pageable memory
// run at a fraction of the PCIe rate).
void *pinned_acquire(size_t bytes, size_t *cap_out);
void pinned_release(void *p, size_t cap);

// vector-like result column in pinned memory
template <class T>
class HostCol {
 public:
  HostCol() {}
  HostCol(const HostCol &) = delete;
  HostCol &operator=(const HostCol &) = delete;
  ~HostCol() {
    if (p_) pinned_release(p_, cap_bytes_);
  }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  T *data() { return p_; }
  const T *data() const { return p_; }
  T *begin() { return p_; }
  const T *begin() const { return p_; }
  T &operator[](size_t i) { return p_[i]; }
  const T &operator[](size_t i) const { return p_[i]; }
  void resize(size_t n) {
    if (n * sizeof(T) > cap_bytes_) {
      size_t cap = 0;
      T *q = (T *)pinned_acquire(std::max(n * sizeof(T), cap_bytes_ * 2), &cap);
      if (p_) {
        memcpy(q, p_, n_ * sizeof(T));
        pinned_release(p_, cap_bytes_);
      }
      p_ = q;
      cap_bytes_ = cap;
    }
    n_ = n;
  }
  // capacity for n elements without changing the size (one block instead of a chain of doublings, each of
  // which copies what is there)
  void reserve(size_t n) {
    if (n * sizeof(T) <= cap_bytes_) return;
    const size_t keep = n_;
    resize(n);
    n_ = keep;
  }
  void assign(size_t n, const T &v) {
    resize(n);
    for (size_t i = 0; i < n; i++) p_[i] = v;
  }
  void push_back(const T &v) {
    resize(n_ + 1);
    p_[n_ - 1] = v;
  }

 private:
  T *p_ = nullptr;
  size_t n_ = 0, cap_bytes_ = 0;
};
}  // namespace impgx

struct impgx_results {
  int device = 0;
  bool on_device = false;
  void *stream = nullptr;  // stream the device columns were allocated on
  size_t n_rows = 0, n_results = 0, n_cig = 0;
  bool has_cigar = false;
  // host columns
  impgx::HostCol<uint64_t> row_off, cig_off;
  impgx::HostCol<uint32_t> qid, tid, cig;
  impgx::HostCol<int32_t> qf, ql, tf, tl;
  // device columns (device-resident variant)
  uint64_t *d_row_off = nullptr, *d_cig_off = nullptr;
  uint32_t *d_qid = nullptr, *d_tid = nullptr, *d_cig = nullptr;
  int32_t *d_qf = nullptr, *d_ql = nullptr, *d_tf = nullptr, *d_tl = nullptr;
  ~impgx_results();
};
