// Collective backends of the multi-GPU single-system path (comm.cuh) and their C-ABI entry points.
// The reference has no distributed path (no MPI / NCCL); BASELINE.json north_star asks for
// "row-block assembly ... joined by an NCCL allreduce, and a block-cyclic Cholesky whose panels are broadcast
// over NVLink" -- these are the collectives behind it.
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "comm.cuh"
#include "common.cuh"

namespace mchrg {

// ------------------------------------------------------------------------------------------------
// NCCL, loaded at run time (the product links no communication library; a process that already holds a
// libnccl.so.2 -- e.g. torch's -- shares it)
// ------------------------------------------------------------------------------------------------
namespace nccl {
struct UniqueId { char internal[128]; };
typedef struct ncclComm* comm_t;
enum { Int8 = 0, Float64 = 8 };  // ncclDataType_t: ncclInt8 = ncclChar = 0, ncclFloat64 = ncclDouble = 8
enum { Sum = 0 };                // ncclRedOp_t

struct Api {
   void* handle = nullptr;
   int (*GetVersion)(int*) = nullptr;
   int (*GetUniqueId)(UniqueId*) = nullptr;
   int (*CommInitRank)(comm_t*, int, UniqueId, int) = nullptr;
   int (*CommDestroy)(comm_t) = nullptr;
   int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
   int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
   int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   const char* (*GetErrorString)(int) = nullptr;
   std::string error;
};

static Api& api() {
   static Api a;
   static std::once_flag once;
   std::call_once(once, [] {
      const char* env = std::getenv("MCHRG_B200_NCCL_LIB");
      if (env && *env) a.handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
      if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process
      if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!a.handle) {
         a.error = std::string("libnccl.so.2 not found (set MCHRG_B200_NCCL_LIB): ") + dlerror();
         return;
      }
      auto sym = [&](const char* name) {
         void* p = dlsym(a.handle, name);
         if (!p && a.error.empty()) a.error = std::string("libnccl lacks ") + name;
         return p;
      };
      a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
      a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
      a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
      a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
      a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
      a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
      a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
      a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
      a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
      a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
   });
   if (!a.error.empty()) fail(MCHRG_B200_ERR_CUDA, "NCCL unavailable: %s", a.error.c_str());
   return a;
}

#define MCHRG_NCCL(call)                                                                                 \
   do {                                                                                                  \
      const int rc__ = (call);                                                                           \
      if (rc__ != 0)                                                                                     \
         ::mchrg::fail(MCHRG_B200_ERR_CUDA, "%s failed: %s", #call, ::mchrg::nccl::api().GetErrorString(rc__)); \
   } while (0)

struct NcclComm : Comm {
   comm_t comm = nullptr;
   ~NcclComm() override {
      if (comm) api().CommDestroy(comm);
   }
   const char* backend() const override { return "nccl"; }
   void bcast(void* buf, size_t bytes, int root, cudaStream_t s) override {
      MCHRG_NCCL(api().Broadcast(buf, buf, bytes, Int8, root, comm, s));
   }
   void allreduce_sum(double* buf, size_t count, cudaStream_t s) override {
      MCHRG_NCCL(api().AllReduce(buf, buf, count, Float64, Sum, comm, s));
   }
   void allgather(const void* send, void* recv, size_t bytes, cudaStream_t s) override {
      MCHRG_NCCL(api().AllGather(send, recv, bytes, Int8, comm, s));
   }
   void group_begin() override { MCHRG_NCCL(api().GroupStart()); }
   void group_end() override { MCHRG_NCCL(api().GroupEnd()); }
};
}  // namespace nccl

// ------------------------------------------------------------------------------------------------
// local backend: the ranks are host threads of one process.  Per collective: publish buffer + "ready" event,
// host rendezvous A, stream-ordered (peer) copies, "done" events, host rendezvous B, the owner of a buffer
// that was read orders its stream after the readers.  The next collective's rendezvous A separates the
// re-recording of the events from the waits of this one.
// ------------------------------------------------------------------------------------------------
struct LocalGroup {
   int n = 0;
   std::atomic<int> count{0};
   std::atomic<uint64_t> gen{0};
   std::atomic<int> broken{0};
   std::vector<const void*> ptr;
   std::vector<cudaEvent_t> ready, done;
   std::vector<int> dev;
   void barrier() {
      const uint64_t g = gen.load(std::memory_order_acquire);
      if (count.fetch_add(1, std::memory_order_acq_rel) + 1 == n) {
         count.store(0, std::memory_order_relaxed);
         gen.fetch_add(1, std::memory_order_release);
         return;
      }
      const auto t0 = std::chrono::steady_clock::now();
      // MCHRG_B200_COMM_TIMEOUT_S: how long a rank waits for its peers before it declares the group broken (default 120)
      const double limit = (double)std::max<long long>(1, env_ll("MCHRG_B200_COMM_TIMEOUT_S", 120));
      for (uint64_t spin = 0; gen.load(std::memory_order_acquire) == g; ++spin) {
         if (broken.load(std::memory_order_relaxed)) fail(MCHRG_B200_ERR_CUDA, "local collective: a peer rank failed");
         if (spin > 2000) std::this_thread::yield();
         if ((spin & 0xffff) == 0xffff &&
             std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit) {
            broken.store(1);
            fail(MCHRG_B200_ERR_CUDA, "local collective timed out: the ranks did not issue the same collectives");
         }
      }
   }
};

__global__ void rank_sum_kernel(const double* __restrict__ parts, size_t count, int n, double* __restrict__ out) {
   const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= count) return;
   double acc = 0.0;
   for (int r = 0; r < n; ++r) acc += parts[(size_t)r * count + i];
   out[i] = acc;
}

struct LocalComm : Comm {
   std::shared_ptr<LocalGroup> g;
   int device = 0;
   cudaEvent_t ready = nullptr, done = nullptr;
   double* scratch = nullptr;
   size_t scratch_count = 0;
   ~LocalComm() override {
      if (g) g->broken.store(1);  // a group is only as alive as all of its members
      cudaSetDevice(device);
      if (ready) cudaEventDestroy(ready);
      if (done) cudaEventDestroy(done);
      if (scratch) cudaFree(scratch);
   }
   const char* backend() const override { return "local"; }
   void copy(void* dst, const void* src, int src_dev, size_t bytes, cudaStream_t s) {
      if (src_dev == device) MCHRG_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
      else MCHRG_CUDA(cudaMemcpyPeerAsync(dst, device, src, src_dev, bytes, s));
   }
   void publish(const void* buf, cudaStream_t s) {
      MCHRG_CUDA(cudaEventRecord(ready, s));
      g->ptr[rank] = buf;
   }
   void bcast(void* buf, size_t bytes, int root, cudaStream_t s) override {
      if (rank == root) publish(buf, s);
      g->barrier();
      if (rank != root) {
         MCHRG_CUDA(cudaStreamWaitEvent(s, g->ready[root], 0));
         copy(buf, g->ptr[root], g->dev[root], bytes, s);
         MCHRG_CUDA(cudaEventRecord(done, s));
      }
      g->barrier();
      if (rank == root)
         for (int r = 0; r < nranks; ++r)
            if (r != root) MCHRG_CUDA(cudaStreamWaitEvent(s, g->done[r], 0));
   }
   void gather_all(const void* send, void* recv, size_t bytes, cudaStream_t s) {
      publish(send, s);
      g->barrier();
      for (int r = 0; r < nranks; ++r) {
         if (r != rank) MCHRG_CUDA(cudaStreamWaitEvent(s, g->ready[r], 0));
         copy((char*)recv + (size_t)r * bytes, g->ptr[r], g->dev[r], bytes, s);
      }
      MCHRG_CUDA(cudaEventRecord(done, s));
      g->barrier();
      for (int r = 0; r < nranks; ++r)  // my send buffer may change only after every rank has read it
         if (r != rank) MCHRG_CUDA(cudaStreamWaitEvent(s, g->done[r], 0));
   }
   void allreduce_sum(double* buf, size_t count, cudaStream_t s) override {
      if (count * nranks > scratch_count) {
         MCHRG_CUDA(cudaStreamSynchronize(s));
         if (scratch) cudaFree(scratch);
         scratch = nullptr;
         scratch_count = count * nranks;
         MCHRG_CUDA(cudaMalloc(&scratch, scratch_count * sizeof(double)));
      }
      gather_all(buf, scratch, count * sizeof(double), s);
      rank_sum_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(scratch, count, nranks, buf);
      MCHRG_CUDA(cudaGetLastError());
   }
   void allgather(const void* send, void* recv, size_t bytes, cudaStream_t s) override { gather_all(send, recv, bytes, s); }
};

static void drop_comm(mchrg_b200_context* ctx) {
   if (!ctx->comm) return;
   cudaSetDevice(ctx->device);
   if (ctx->stream_comm) cudaStreamSynchronize(ctx->stream_comm);
   delete ctx->comm;
   ctx->comm = nullptr;
}

}  // namespace mchrg

using namespace mchrg;

// run `body` on the context's communication stream, ordered after the main stream, and order the main stream
// after it again (entry points below are synchronous with respect to the stream order)
template <class F>
static int comm_call(mchrg_b200_context* ctx, F&& body) {
   return guarded(ctx, [&]() {
      if (!ctx->comm) fail(MCHRG_B200_ERR_ARG, "no communicator installed on this context");
      cudaEvent_t e = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(e, ctx->stream));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream_comm, e, 0));
      body();
      cudaEvent_t e2 = ctx->next_event();
      MCHRG_CUDA(cudaEventRecord(e2, ctx->stream_comm));
      MCHRG_CUDA(cudaStreamWaitEvent(ctx->stream, e2, 0));
      return 0;
   });
}

extern "C" {

int mchrg_b200_comm_unique_id(void* id) {
   if (!id) return MCHRG_B200_ERR_ARG;
   try {
      nccl::UniqueId u;
      MCHRG_NCCL(nccl::api().GetUniqueId(&u));
      std::memcpy(id, &u, sizeof u);
      return 0;
   } catch (const Error& e) {
      set_last_error(e.msg);
      return e.code;
   }
}

int mchrg_b200_comm_init_nccl(mchrg_b200_context* ctx, int32_t rank, int32_t nranks, const void* id) {
   if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) {
      set_last_error("comm_init_nccl: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   std::lock_guard<std::mutex> lock(ctx->mu);
   try {
      ctx->use();
      drop_comm(ctx);
      if (nranks == 1) return 0;
      auto c = std::make_unique<nccl::NcclComm>();
      c->rank = rank;
      c->nranks = nranks;
      nccl::UniqueId u;
      std::memcpy(&u, id, sizeof u);
      MCHRG_NCCL(nccl::api().CommInitRank(&c->comm, nranks, u, rank));
      ctx->comm = c.release();
      return 0;
   } catch (const Error& e) {
      set_last_error(e.msg);
      return e.code;
   }
}

int mchrg_b200_comm_init_local(mchrg_b200_context** ctxs, int32_t n) {
   if (!ctxs || n < 1) {
      set_last_error("comm_init_local: invalid argument");
      return MCHRG_B200_ERR_ARG;
   }
   for (int r = 0; r < n; ++r)
      if (!ctxs[r]) {
         set_last_error("comm_init_local: null context");
         return MCHRG_B200_ERR_ARG;
      }
   try {
      auto g = std::make_shared<LocalGroup>();
      g->n = n;
      g->ptr.assign(n, nullptr);
      g->ready.assign(n, nullptr);
      g->done.assign(n, nullptr);
      g->dev.assign(n, 0);
      for (int r = 0; r < n; ++r) {
         mchrg_b200_context* ctx = ctxs[r];
         std::lock_guard<std::mutex> lock(ctx->mu);
         ctx->use();
         drop_comm(ctx);
         if (n == 1) continue;
         auto c = std::make_unique<LocalComm>();
         c->rank = r;
         c->nranks = n;
         c->device = ctx->device;
         c->g = g;
         MCHRG_CUDA(cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
         MCHRG_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
         g->ready[r] = c->ready;
         g->done[r] = c->done;
         g->dev[r] = ctx->device;
         ctx->comm = c.release();
      }
      return 0;
   } catch (const Error& e) {
      set_last_error(e.msg);
      return e.code;
   }
}

int mchrg_b200_comm_destroy(mchrg_b200_context* ctx) {
   if (!ctx) return MCHRG_B200_ERR_ARG;
   std::lock_guard<std::mutex> lock(ctx->mu);
   drop_comm(ctx);
   return 0;
}

int32_t mchrg_b200_comm_rank(const mchrg_b200_context* ctx) { return (ctx && ctx->comm) ? ctx->comm->rank : 0; }
int32_t mchrg_b200_comm_size(const mchrg_b200_context* ctx) { return (ctx && ctx->comm) ? ctx->comm->nranks : 1; }

int mchrg_b200_comm_bcast(mchrg_b200_context* ctx, void* dev_buf, int64_t bytes, int32_t root) {
   return comm_call(ctx, [&]() { ctx->comm->bcast(dev_buf, (size_t)bytes, root, ctx->stream_comm); });
}
int mchrg_b200_comm_allreduce_sum(mchrg_b200_context* ctx, double* dev_buf, int64_t count) {
   return comm_call(ctx, [&]() { ctx->comm->allreduce_sum(dev_buf, (size_t)count, ctx->stream_comm); });
}
int mchrg_b200_comm_allgather(mchrg_b200_context* ctx, const void* dev_send, void* dev_recv, int64_t bytes) {
   return comm_call(ctx, [&]() { ctx->comm->allgather(dev_send, dev_recv, (size_t)bytes, ctx->stream_comm); });
}

}  // extern "C"
