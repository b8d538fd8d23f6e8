// cube_comm.cuh -- image-to-image transport of libcubegpu.so (replaces the coarray GETs + `sync all` of the reference,
// SURVEY.md sec. 2.4).  Two back ends behind one interface:
//   NcclComm   one process per GPU (the product path): NCCL send/recv over NVLink/NVSwitch.  libnccl.so.2 is opened
//              at run time (dlopen) so that a process that already loaded torch's NCCL shares that copy.
//   LocalComm  all images are host threads of ONE process (the `-fcoarray=single`-style deployment, several images
//              per GPU, and the single-GPU parity tests): device-to-device copies ordered by CUDA events + a host
//              barrier.  Same call sequence, same message matching rule as NCCL (in order per peer pair).
// Messages between two images are matched in issue order, exactly like grouped ncclSend/ncclRecv.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace cube {

struct Comm {
  int rank = 0, size = 1;
  std::string err;
  struct Msg { const void* src; void* dst; size_t bytes; int peer; };
  std::vector<Msg> self_send, self_recv;
  virtual ~Comm() {}
  virtual int group_begin(cudaStream_t st) = 0;
  virtual int send_peer(const void* p, size_t bytes, int peer) = 0;
  virtual int recv_peer(void* p, size_t bytes, int peer) = 0;
  virtual int group_end_peer() = 0;
  // gather `bytes` from every image into out[size*bytes] (host memory), blocking; the caller reduces in image order,
  // like the reference's gather-to-image-1 loops (update_particle.f90:154-166, pm.f90:239-244)
  virtual int allgather_host(const void* in, void* out, size_t bytes, cudaStream_t st) = 0;
  virtual void abort_group() {}

  cudaStream_t st_ = nullptr;
  int begin(cudaStream_t st) { st_ = st; self_send.clear(); self_recv.clear(); return group_begin(st); }
  int send(const void* p, size_t bytes, int peer) {
    if (peer == rank) { self_send.push_back({p, nullptr, bytes, peer}); return 0; }
    return send_peer(p, bytes, peer);
  }
  int recv(void* p, size_t bytes, int peer) {
    if (peer == rank) { self_recv.push_back({nullptr, p, bytes, peer}); return 0; }
    return recv_peer(p, bytes, peer);
  }
  int end() {
    if (self_send.size() != self_recv.size()) { err = "comm: unmatched self messages"; return 1; }
    for (size_t i = 0; i < self_send.size(); i++) {
      if (self_send[i].bytes != self_recv[i].bytes) { err = "comm: self message size mismatch"; return 1; }
      if (self_send[i].bytes && cudaMemcpyAsync(self_recv[i].dst, self_send[i].src, self_send[i].bytes, cudaMemcpyDeviceToDevice, st_) != cudaSuccess) {
        err = "comm: self copy failed"; return 1;
      }
    }
    return group_end_peer();
  }
};

// ---------------------------------------------------------------------------------------------
// NCCL
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy torch (or the host program) already loaded
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot open libnccl.so.2: ") + dlerror(); return false; }
#define CUBE_NCCL_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(lib, "nccl" #f)); if (!f) { err = "libnccl.so.2 lacks nccl" #f; lib = nullptr; return false; }
    CUBE_NCCL_SYM(GetUniqueId) CUBE_NCCL_SYM(CommInitRank) CUBE_NCCL_SYM(CommDestroy) CUBE_NCCL_SYM(CommAbort) CUBE_NCCL_SYM(Send) CUBE_NCCL_SYM(Recv)
    CUBE_NCCL_SYM(AllGather) CUBE_NCCL_SYM(GroupStart) CUBE_NCCL_SYM(GroupEnd) CUBE_NCCL_SYM(GetErrorString)
#undef CUBE_NCCL_SYM
    return true;
  }
};
inline NcclApi& nccl_api() { static NcclApi a; return a; }

struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  char* dbuf = nullptr;  // device staging for allgather_host
  HostStage stage;       // host side of it: mapped memory, no copy engine (cube_common.cuh)
  static constexpr size_t kMaxGather = 256;
  int init(int rank_, int size_, const void* id128) {
    rank = rank_; size = size_;
    NcclApi& a = nccl_api();
    if (!a.load()) { err = a.err; return 1; }
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    ncclResult_t r = a.CommInitRank(&comm, size, id, rank);
    if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + a.GetErrorString(r); return 1; }
    if (cudaMalloc((void**)&dbuf, kMaxGather * (size_t)(size + 1)) != cudaSuccess) { err = "cudaMalloc (comm staging)"; return 1; }
    return 0;
  }
  ~NcclComm() override {
    if (dbuf) cudaFree(dbuf);
    stage.destroy();
    if (comm) nccl_api().CommDestroy(comm);
  }
  int ck(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return 0;
    err = std::string(what) + ": " + nccl_api().GetErrorString(r);
    return 1;
  }
  // a local failure between two collectives: tear the communicator down so that the peers' pending operations end with an error
  // instead of waiting for this image for ever (the reference `stop`s every image)
  void abort_group() override { if (comm) { nccl_api().CommAbort(comm); comm = nullptr; } }
  int group_begin(cudaStream_t) override { return ck(nccl_api().GroupStart(), "ncclGroupStart"); }
  int send_peer(const void* p, size_t bytes, int peer) override { return ck(nccl_api().Send(p, bytes, ncclInt8, peer, comm, st_), "ncclSend"); }
  int recv_peer(void* p, size_t bytes, int peer) override { return ck(nccl_api().Recv(p, bytes, ncclInt8, peer, comm, st_), "ncclRecv"); }
  int group_end_peer() override { return ck(nccl_api().GroupEnd(), "ncclGroupEnd"); }
  int allgather_host(const void* in, void* out, size_t bytes, cudaStream_t st) override {
    if (bytes > kMaxGather) { err = "allgather_host: message too large"; return 1; }
    if (stage.init() != cudaSuccess) { err = "allgather staging"; return 1; }
    if (stage.write(dbuf, in, bytes, st) != cudaSuccess) { err = "allgather H2D"; return 1; }
    if (ck(nccl_api().AllGather(dbuf, dbuf + kMaxGather, bytes, ncclInt8, comm, st), "ncclAllGather")) return 1;
    if (stage.read(out, dbuf + kMaxGather, bytes * size, st) != cudaSuccess) { err = "allgather D2H"; return 1; }
    if (stage.sync(st) != cudaSuccess) { err = "allgather sync"; return 1; }
    return 0;
  }
};

// ---------------------------------------------------------------------------------------------
// in-process group: every image is a host thread of this process
// ---------------------------------------------------------------------------------------------
struct LocalGroup {
  std::mutex m;
  std::condition_variable cv;
  int size = 0, arrived = 0, members = 0;
  long long gen = 0;
  bool failed = false;
  std::vector<std::vector<Comm::Msg>> sends;   // [rank] messages posted in the current exchange
  std::vector<cudaEvent_t> ready;             // [rank] "my send buffers are complete"
  std::vector<std::vector<char>> gather;      // [rank]
  // returns false on failure / time-out (a peer image died): never dead-locks the test suite
  bool barrier() {
    std::unique_lock<std::mutex> lk(m);
    if (failed) return false;
    const long long g = gen;
    if (++arrived == size) { arrived = 0; gen++; cv.notify_all(); return true; }
    const bool ok = cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g || failed; });
    if (!ok || failed) { failed = true; cv.notify_all(); return false; }
    return true;
  }
  void fail() { std::lock_guard<std::mutex> lk(m); failed = true; cv.notify_all(); }
};
inline std::mutex& local_registry_mutex() { static std::mutex m; return m; }
inline std::map<int, std::shared_ptr<LocalGroup>>& local_registry() { static std::map<int, std::shared_ptr<LocalGroup>> r; return r; }

struct LocalComm : Comm {
  std::shared_ptr<LocalGroup> g;
  int key = 0;
  std::vector<Msg> recvs;
  cudaEvent_t ev = nullptr;
  int init(int rank_, int size_, int key_) {
    rank = rank_; size = size_; key = key_;
    {
      std::lock_guard<std::mutex> lk(local_registry_mutex());
      auto& r = local_registry()[key];
      if (!r) { r = std::make_shared<LocalGroup>(); r->size = size; r->sends.resize(size); r->ready.assign(size, nullptr); r->gather.resize(size); }
      if (r->size != size) { err = "local group: inconsistent image count"; return 1; }
      g = r; g->members++;
    }
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { err = "cudaEventCreate"; return 1; }
    g->ready[rank] = ev;
    if (!g->barrier()) { err = "local group: a peer image failed during init"; return 1; }
    return 0;
  }
  ~LocalComm() override {
    if (g) {
      std::lock_guard<std::mutex> lk(local_registry_mutex());
      if (--g->members == 0) local_registry().erase(key);
    }
    if (ev) cudaEventDestroy(ev);
  }
  void abort_group() override { if (g) g->fail(); }
  int group_begin(cudaStream_t) override { g->sends[rank].clear(); recvs.clear(); return 0; }
  int send_peer(const void* p, size_t bytes, int peer) override { g->sends[rank].push_back({p, nullptr, bytes, peer}); return 0; }
  int recv_peer(void* p, size_t bytes, int peer) override { recvs.push_back({nullptr, p, bytes, peer}); return 0; }
  int group_end_peer() override {
    if (cudaEventRecord(ev, st_) != cudaSuccess) { err = "cudaEventRecord"; g->fail(); return 1; }
    if (!g->barrier()) { err = "local group: a peer image failed"; return 1; }
    std::vector<size_t> cursor(size, 0);
    int rc = 0;
    for (const Msg& r : recvs) {
      const std::vector<Msg>& ps = g->sends[r.peer];
      size_t& c = cursor[r.peer];
      while (c < ps.size() && ps[c].peer != rank) c++;
      if (c == ps.size()) { err = "local group: receive without a matching send"; rc = 1; break; }
      if (ps[c].bytes != r.bytes) { err = "local group: message size mismatch"; rc = 1; break; }
      if (cudaStreamWaitEvent(st_, g->ready[r.peer], 0) != cudaSuccess) { err = "cudaStreamWaitEvent"; rc = 1; break; }
      if (r.bytes && cudaMemcpyAsync(r.dst, ps[c].src, r.bytes, cudaMemcpyDefault, st_) != cudaSuccess) { err = "peer copy failed"; rc = 1; break; }
      c++;
    }
    if (!rc && cudaStreamSynchronize(st_) != cudaSuccess) { err = "peer copy sync failed"; rc = 1; }
    if (rc) { g->fail(); return 1; }
    if (!g->barrier()) { err = "local group: a peer image failed"; return 1; }  // send buffers are reusable from here
    return 0;
  }
  int allgather_host(const void* in, void* out, size_t bytes, cudaStream_t) override {
    g->gather[rank].assign((const char*)in, (const char*)in + bytes);
    if (!g->barrier()) { err = "local group: a peer image failed"; return 1; }
    for (int r = 0; r < size; r++) {
      if (g->gather[r].size() != bytes) { err = "local group: allgather size mismatch"; g->fail(); return 1; }
      memcpy((char*)out + (size_t)r * bytes, g->gather[r].data(), bytes);
    }
    if (!g->barrier()) { err = "local group: a peer image failed"; return 1; }
    return 0;
  }
};

}  // namespace cube
