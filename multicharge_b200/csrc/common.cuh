// Shared host/device infrastructure of libmchrg_b200: error handling, the context (one GPU, one
// stream, grow-only device workspace), host<->device staging of caller arrays.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <cstdlib>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mchrg_b200.h"
#include "comm.cuh"

namespace mchrg {

struct Error {
   int code;
   std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
   char buf[512];
   va_list ap;
   va_start(ap, fmt);
   vsnprintf(buf, sizeof buf, fmt, ap);
   va_end(ap);
   throw Error{code, buf};
}

#define MCHRG_CUDA(call)                                                                            \
   do {                                                                                             \
      cudaError_t err__ = (call);                                                                   \
      if (err__ != cudaSuccess)                                                                     \
         ::mchrg::fail(MCHRG_B200_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,    \
                       cudaGetErrorString(err__));                                                  \
   } while (0)

void set_last_error(const std::string& msg);

constexpr int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// tuning knobs are read from the environment on every call (no process-wide caching: tests and callers may
// change them between calls)
inline long long env_ll(const char* name, long long def) {
   const char* e = std::getenv(name);
   return (e && *e) ? std::atoll(e) : def;
}

}  // namespace mchrg

// One GPU + one stream + a grow-only bump arena.  Calls on one context are serialised by `mu`;
// use one context per host thread for concurrent solves (the reference's solve is re-entrant,
// SURVEY 8b "Threading").
struct mchrg_b200_context {
   int device = 0;
   cudaStream_t stream = nullptr;     // main stream: every entry point is ordered on it
   cudaStream_t stream_hi = nullptr;  // high-priority side stream (Cholesky panel look-ahead)
   cudaStream_t stream_hi2 = nullptr; // second high-priority stream (batch factorisation kernels)
   cudaStream_t stream_aux = nullptr; // second normal-priority stream (batch size classes run concurrently)
   cudaStream_t stream_in = nullptr;  // host->device copies of pipelined batches
   cudaStream_t stream_out = nullptr; // device->host copies of pipelined batches
   cudaStream_t active = nullptr;     // stream MCHRG_LAUNCH currently targets
   std::vector<cudaEvent_t> events;   // reusable, timing disabled
   size_t events_used = 0;
   cudaEvent_t next_event();
   int sm_count = 148;
   double2* erf_tab = nullptr;        // (erf, exp(-x^2)) nodes of erfgauss.cuh, device memory
   // collectives of the multi-GPU single-system path (comm.cuh); nullptr: single GPU.  Every collective is
   // enqueued on stream_comm and tied to the compute streams with events.
   mchrg::Comm* comm = nullptr;
   cudaStream_t stream_comm = nullptr;
   int nranks() const;
   int rank() const;
   size_t smem_optin = 0;
   // kernels whose dynamic shared-memory limit has been raised through THIS context (cudaFuncSetAttribute acts on
   // the current device only, so the record is per context, not per process)
   std::unordered_map<const void*, size_t> func_smem;
   template <class K>
   void ensure_smem(K* kernel, size_t bytes, bool max_carveout = false) {
      auto it = func_smem.find((const void*)kernel);
      if (it != func_smem.end() && it->second >= bytes) return;
      MCHRG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      if (max_carveout)
         MCHRG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)cudaSharedmemCarveoutMaxShared));
      func_smem[(const void*)kernel] = bytes;
   }
   std::mutex mu;
   uint64_t launches = 0;
   // size classes / permutation / scratch layout of the last large device-resident batch (batch.cu, SplitPlan)
   std::shared_ptr<void> batch_plan;
   // MCHRG_B200_TRACE=1: serial per-phase times of the last split-path batch (pair A, factorisation, pair C)
   double trace_ms[3] = {0.0, 0.0, 0.0};
   int trace_launches[3] = {0, 0, 0};

   // bump arena: one cudaMalloc'ed slab, reset at the start of every entry point; when a call
   // needs more than the slab holds, extra slabs are chained and merged on the next reset.
   struct Slab {
      char* base = nullptr;
      size_t size = 0;
      size_t used = 0;
   };
   std::vector<Slab> slabs;
   size_t high_water = 0;

   // pinned staging buffer for small host transfers (status words, scalars)
   void* pinned = nullptr;
   size_t pinned_size = 0;

   // deferred device->host copies of output arrays (executed by finish())
   struct Copyback {
      void* host;
      const void* dev;
      size_t bytes;
   };
   std::vector<Copyback> copybacks;

   void use() const { MCHRG_CUDA(cudaSetDevice(device)); }

   void arena_reset();
   void* arena_alloc(size_t bytes);
   template <class T>
   T* alloc(size_t n) { return static_cast<T*>(arena_alloc(n * sizeof(T))); }
   template <class T>
   T* alloc_zero(size_t n) {
      T* p = alloc<T>(n);
      MCHRG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), stream));
      return p;
   }

   static bool is_device_ptr(const void* p);
   // device-visible view of a caller input array (copies host arrays into the arena)
   template <class T>
   const T* in(const T* p, size_t n) {
      if (p == nullptr || n == 0) return p;
      if (is_device_ptr(p)) return p;
      T* d = alloc<T>(n);
      MCHRG_CUDA(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, stream));
      return d;
   }
   // device buffer standing for a caller output array; host arrays are copied back by finish().
   // `preload` stages the current host content first (for accumulate-into outputs).
   template <class T>
   T* out(T* p, size_t n, bool preload = false) {
      if (p == nullptr || n == 0) return p;
      if (is_device_ptr(p)) return p;
      T* d = alloc<T>(n);
      if (preload) MCHRG_CUDA(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, stream));
      copybacks.push_back({p, d, n * sizeof(T)});
      return d;
   }
   void finish();  // run the copy-backs and synchronise the stream

   void* pinned_buf(size_t bytes);
};

namespace mchrg {

// kernel launch bookkeeping: count + error check
inline void after_launch(mchrg_b200_context* ctx, const char* name) {
   ctx->launches++;
   cudaError_t err = cudaGetLastError();
   if (err != cudaSuccess) fail(MCHRG_B200_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err));
}

#define MCHRG_LAUNCH(ctx, kernel, grid, block, smem, ...)                       \
   do {                                                                          \
      kernel<<<(grid), (block), (smem), (ctx)->active>>>(__VA_ARGS__);           \
      ::mchrg::after_launch((ctx), #kernel);                                     \
   } while (0)

// switches the launch target of MCHRG_LAUNCH for a scope
struct StreamScope {
   mchrg_b200_context* ctx;
   cudaStream_t prev;
   StreamScope(mchrg_b200_context* c, cudaStream_t s) : ctx(c), prev(c->active) { c->active = s; }
   ~StreamScope() { ctx->active = prev; }
};

// Runs `body` under the context lock with a fresh arena and converts exceptions to status codes.
template <class F>
int guarded(mchrg_b200_context* ctx, F&& body) {
   if (ctx == nullptr) {
      set_last_error("null context");
      return MCHRG_B200_ERR_ARG;
   }
   std::lock_guard<std::mutex> lock(ctx->mu);
   try {
      ctx->use();
      ctx->arena_reset();
      ctx->copybacks.clear();
      ctx->active = ctx->stream;
      ctx->events_used = 0;
      int rc = body();
      ctx->finish();
      // a call that outgrew the workspace leaves chained slabs: merge them NOW (this call pays for its own growth), not
      // at the start of the next call, whose time a caller may be measuring
      if (ctx->slabs.size() > 1) ctx->arena_reset();
      return rc;
   } catch (const Error& e) {
      set_last_error(e.msg);
      ctx->copybacks.clear();
      ctx->active = ctx->stream;
      cudaStreamSynchronize(ctx->stream);
      if (ctx->stream_hi) cudaStreamSynchronize(ctx->stream_hi);
      if (ctx->stream_hi2) cudaStreamSynchronize(ctx->stream_hi2);
      if (ctx->stream_aux) cudaStreamSynchronize(ctx->stream_aux);
      if (ctx->stream_in) cudaStreamSynchronize(ctx->stream_in);
      if (ctx->stream_out) cudaStreamSynchronize(ctx->stream_out);
      if (ctx->stream_comm) cudaStreamSynchronize(ctx->stream_comm);
      cudaGetLastError();
      return e.code;
   } catch (const std::exception& e) {
      set_last_error(e.what());
      ctx->copybacks.clear();
      return MCHRG_B200_ERR_ALLOC;
   }
}

}  // namespace mchrg
