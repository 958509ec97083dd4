// comm.cu — transports of the sharded pipeline's exchange steps (host code only).
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"

namespace impgx {

// ------------------------------------------------------------------ NCCL
namespace {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // reuse the copy already mapped into the process (torch's), else the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
#define LOAD(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name))
    LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommDestroy); LOAD(AllGather); LOAD(Send); LOAD(Recv);
    LOAD(GroupStart); LOAD(GroupEnd); LOAD(GetErrorString);
#undef LOAD
  });
  REQUIRE(api.handle && api.GetUniqueId && api.CommInitRank && api.AllGather && api.Send && api.Recv && api.GroupStart &&
              api.GroupEnd,
          IMPGX_E_UNSUPPORTED, "libnccl.so.2 could not be loaded (the sharded index needs NCCL for multi-process runs)");
  return api;
}

#define NCCL_CHECK(expr)                                                                                         \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != ncclSuccess)                                                                                       \
      throw ::impgx::Error(IMPGX_E_CUDA, std::string(#expr) + ": " +                                             \
                                             (nccl().GetErrorString ? nccl().GetErrorString(_r) : "NCCL error")); \
  } while (0)

class NcclComm : public Comm {
 public:
  NcclComm(const uint8_t id[128], int rank, int n_ranks, int device) {
    rank_ = rank;
    size_ = n_ranks;
    device_ = device;
    CUDA_CHECK(cudaSetDevice(device));
    ncclUniqueId uid;
    static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
    memcpy(&uid, id, 128);
    NCCL_CHECK(nccl().CommInitRank(&comm_, n_ranks, uid, rank));
    CUDA_CHECK(cudaMalloc((void **)&d_small_, small_cap_ * 8));
  }
  ~NcclComm() override {
    cudaSetDevice(device_);
    if (comm_ && nccl().CommDestroy) nccl().CommDestroy(comm_);
    cudaFree(d_small_);
  }
  const char *kind() const override { return "nccl"; }

  void allgather_u64(const uint64_t *mine, size_t n, uint64_t *out, cudaStream_t s) override {
    const size_t need = n * ((size_t)size_ + 1);
    if (need > small_cap_) {
      CUDA_CHECK(cudaStreamSynchronize(s));
      cudaFree(d_small_);
      small_cap_ = need * 2;
      CUDA_CHECK(cudaMalloc((void **)&d_small_, small_cap_ * 8));
    }
    uint64_t *d_in = d_small_, *d_out = d_small_ + n;
    CUDA_CHECK(cudaMemcpyAsync(d_in, mine, n * 8, cudaMemcpyHostToDevice, s));
    NCCL_CHECK(nccl().AllGather(d_in, d_out, n, ncclUint64, comm_, s));
    CUDA_CHECK(cudaMemcpyAsync(out, d_out, n * 8 * size_, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    exchanges++;
  }

  void alltoallv(const void *d_send, const uint64_t *send_cnt, const uint64_t *send_off, void *d_recv,
                 const uint64_t *recv_cnt, const uint64_t *recv_off, size_t eb, cudaStream_t s) override {
    const char *src = (const char *)d_send;
    char *dst = (char *)d_recv;
    NCCL_CHECK(nccl().GroupStart());
    for (int p = 0; p < size_; p++) {
      if (p == rank_) continue;
      if (send_cnt[p]) NCCL_CHECK(nccl().Send(src + send_off[p] * eb, send_cnt[p] * eb, ncclUint8, p, comm_, s));
      if (recv_cnt[p]) NCCL_CHECK(nccl().Recv(dst + recv_off[p] * eb, recv_cnt[p] * eb, ncclUint8, p, comm_, s));
      bytes_sent += send_cnt[p] * eb;
      bytes_received += recv_cnt[p] * eb;
    }
    NCCL_CHECK(nccl().GroupEnd());
    if (send_cnt[rank_])
      CUDA_CHECK(cudaMemcpyAsync(dst + recv_off[rank_] * eb, src + send_off[rank_] * eb, send_cnt[rank_] * eb,
                                 cudaMemcpyDeviceToDevice, s));
    exchanges++;
  }

  void allgatherv(const void *d_send, uint64_t n_mine, void *d_recv, const uint64_t *cnt, const uint64_t *off,
                  size_t eb, cudaStream_t s) override {
    char *dst = (char *)d_recv;
    NCCL_CHECK(nccl().GroupStart());
    for (int p = 0; p < size_; p++) {
      if (p == rank_) continue;
      if (n_mine) NCCL_CHECK(nccl().Send(d_send, n_mine * eb, ncclUint8, p, comm_, s));
      if (cnt[p]) NCCL_CHECK(nccl().Recv(dst + off[p] * eb, cnt[p] * eb, ncclUint8, p, comm_, s));
      bytes_sent += n_mine * eb;
      bytes_received += cnt[p] * eb;
    }
    NCCL_CHECK(nccl().GroupEnd());
    if (n_mine) CUDA_CHECK(cudaMemcpyAsync(dst + off[rank_] * eb, d_send, n_mine * eb, cudaMemcpyDeviceToDevice, s));
    exchanges++;
  }

 private:
  ncclComm_t comm_ = nullptr;
  int device_ = 0;
  uint64_t *d_small_ = nullptr;
  size_t small_cap_ = 4096;
};

// ------------------------------------------------------------------ in-process
struct LocalShared {
  int n = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  bool aborted = false;
  std::vector<std::vector<uint64_t>> slots;               // allgather_u64
  std::vector<const void *> ptr;                          // published send buffers
  std::vector<std::vector<uint64_t>> cnt, off;            // published send counts / offsets
  std::vector<int> dev;

  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) throw Error(IMPGX_E_CUDA, "a peer rank of the local group failed");
    const uint64_t g = gen;
    if (++arrived == n) {
      arrived = 0;
      gen++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g || aborted; });
    }
    if (aborted && gen == g) throw Error(IMPGX_E_CUDA, "a peer rank of the local group failed");
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};

class LocalComm : public Comm {
 public:
  LocalComm(std::shared_ptr<LocalShared> sh, int rank) : sh_(std::move(sh)) {
    rank_ = rank;
    size_ = sh_->n;
  }
  ~LocalComm() override { sh_->abort(); }  // a rank leaving releases peers blocked in a barrier
  const char *kind() const override { return "local"; }
  void abort() override { sh_->abort(); }

  void allgather_u64(const uint64_t *mine, size_t n, uint64_t *out, cudaStream_t s) override {
    CUDA_CHECK(cudaStreamSynchronize(s));
    sh_->slots[rank_].assign(mine, mine + n);
    sh_->barrier();
    for (int p = 0; p < size_; p++) {
      REQUIRE(sh_->slots[p].size() == n, IMPGX_E_INVALID, "ranks disagree on an exchange size");
      memcpy(out + (size_t)p * n, sh_->slots[p].data(), n * 8);
    }
    sh_->barrier();
    exchanges++;
  }

  void alltoallv(const void *d_send, const uint64_t *send_cnt, const uint64_t *send_off, void *d_recv,
                 const uint64_t *recv_cnt, const uint64_t *recv_off, size_t eb, cudaStream_t s) override {
    CUDA_CHECK(cudaStreamSynchronize(s));  // my send buffer is complete
    publish(d_send, send_cnt, send_off);
    sh_->barrier();
    char *dst = (char *)d_recv;
    for (int p = 0; p < size_; p++) {
      const uint64_t c = sh_->cnt[p][rank_];
      REQUIRE(c == recv_cnt[p], IMPGX_E_INVALID, "all-to-all counts disagree between ranks");
      if (!c) continue;
      pull(dst + recv_off[p] * eb, (const char *)sh_->ptr[p] + sh_->off[p][rank_] * eb, c * eb, p, s);
      if (p != rank_) bytes_received += c * eb;
    }
    for (int p = 0; p < size_; p++)
      if (p != rank_) bytes_sent += send_cnt[p] * eb;
    CUDA_CHECK(cudaStreamSynchronize(s));
    sh_->barrier();  // peers are done reading my buffer
    exchanges++;
  }

  void allgatherv(const void *d_send, uint64_t n_mine, void *d_recv, const uint64_t *cnt, const uint64_t *off,
                  size_t eb, cudaStream_t s) override {
    CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<uint64_t> c((size_t)size_, n_mine), o((size_t)size_, 0);
    publish(d_send, c.data(), o.data());
    sh_->barrier();
    char *dst = (char *)d_recv;
    for (int p = 0; p < size_; p++) {
      REQUIRE(sh_->cnt[p][rank_] == cnt[p], IMPGX_E_INVALID, "all-gather counts disagree between ranks");
      if (!cnt[p]) continue;
      pull(dst + off[p] * eb, (const char *)sh_->ptr[p], cnt[p] * eb, p, s);
      if (p != rank_) bytes_received += cnt[p] * eb;
    }
    bytes_sent += n_mine * eb * (uint64_t)(size_ - 1);
    CUDA_CHECK(cudaStreamSynchronize(s));
    sh_->barrier();
    exchanges++;
  }

 private:
  std::shared_ptr<LocalShared> sh_;

  void publish(const void *p, const uint64_t *cnt, const uint64_t *off) {
    int dev = 0;
    cudaGetDevice(&dev);
    sh_->ptr[rank_] = p;
    sh_->cnt[rank_].assign(cnt, cnt + size_);
    sh_->off[rank_].assign(off, off + size_);
    sh_->dev[rank_] = dev;
  }
  void pull(void *dst, const void *src, size_t bytes, int peer, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int pdev = sh_->dev[peer];
    if (pdev == dev) CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
    else CUDA_CHECK(cudaMemcpyPeerAsync(dst, dev, src, pdev, bytes, s));
  }
};

}  // namespace

void nccl_unique_id(uint8_t id[128]) {
  ncclUniqueId uid;
  NCCL_CHECK(nccl().GetUniqueId(&uid));
  memcpy(id, &uid, 128);
}

Comm *nccl_comm_create(const uint8_t id[128], int rank, int n_ranks, int device) {
  REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, IMPGX_E_INVALID, "bad rank / n_ranks");
  return new NcclComm(id, rank, n_ranks, device);
}

std::vector<Comm *> local_comm_group(int n_ranks) {
  REQUIRE(n_ranks >= 1 && n_ranks <= 64, IMPGX_E_INVALID, "local group size must be in [1, 64]");
  auto sh = std::make_shared<LocalShared>();
  sh->n = n_ranks;
  sh->slots.resize(n_ranks);
  sh->ptr.assign(n_ranks, nullptr);
  sh->cnt.resize(n_ranks);
  sh->off.resize(n_ranks);
  sh->dev.assign(n_ranks, 0);
  std::vector<Comm *> out;
  for (int r = 0; r < n_ranks; r++) out.push_back(new LocalComm(sh, r));
  return out;
}

}  // namespace impgx
