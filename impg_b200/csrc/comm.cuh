// comm.cuh — the exchange layer of the target-sharded index.
//
// The sharded pipeline is SPMD: every rank owns the entries, run stream and
// visited sets of its target sequences and runs the same hop loop; between
// hops the ranks exchange (a) lifted hits, routed to the owner of the sequence
// they land on (all-to-all-v), and (b) the next frontier, so that every rank
// knows the global frontier order the reference's result order is built on
// (all-gather-v). Two transports implement it:
//   * NcclComm  — one process per GPU, NCCL over NVLink/NVSwitch (libnccl.so.2
//                 is dlopen'ed: the process usually holds torch's copy already);
//   * LocalComm — the ranks are threads of ONE process (one per device, or
//                 several virtual ranks on one device): peer copies + a barrier.
//                 This is the single-process multi-GPU mode and what the parity
//                 tests use on a 1-GPU box.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <vector>

namespace impgx {

class Comm {
 public:
  virtual ~Comm() {}
  int rank() const { return rank_; }
  int size() const { return size_; }
  virtual const char *kind() const = 0;
  // Every rank contributes `n` u64 values; out[r * n + k] = value k of rank r. Host data;
  // the call synchronises `s` and is a barrier.
  virtual void allgather_u64(const uint64_t *mine, size_t n, uint64_t *out, cudaStream_t s) = 0;
  // Device buffers. send_cnt/send_off (elements) index d_send by destination,
  // recv_cnt/recv_off index d_recv by source. Completes on `s` in stream order;
  // d_send may be released after the call returns (the call synchronises).
  virtual void alltoallv(const void *d_send, const uint64_t *send_cnt, const uint64_t *send_off, void *d_recv,
                         const uint64_t *recv_cnt, const uint64_t *recv_off, size_t elem_bytes, cudaStream_t s) = 0;
  // Every rank contributes n_mine elements; d_recv receives rank r's block at off[r] (cnt[r] elements).
  virtual void allgatherv(const void *d_send, uint64_t n_mine, void *d_recv, const uint64_t *cnt,
                          const uint64_t *off, size_t elem_bytes, cudaStream_t s) = 0;
  // Called by a rank that failed: peers blocked in an exchange are released with an error.
  virtual void abort() {}
  uint64_t bytes_sent = 0, bytes_received = 0, exchanges = 0;

 protected:
  int rank_ = 0, size_ = 1;
};

// NCCL transport. `id` is the 128-byte ncclUniqueId every rank received from rank 0.
void nccl_unique_id(uint8_t id[128]);
Comm *nccl_comm_create(const uint8_t id[128], int rank, int n_ranks, int device);

// In-process transport: returns n_ranks endpoints sharing one rendezvous; each
// must be driven by its own host thread.
std::vector<Comm *> local_comm_group(int n_ranks);

}  // namespace impgx

struct impgx_comm {
  std::unique_ptr<impgx::Comm> c;
};
