// Context, arena and staging implementation (see common.cuh).
#include "common.cuh"
#include "erfgauss.cuh"

namespace mchrg {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

}  // namespace mchrg

using namespace mchrg;

void mchrg_b200_context::arena_reset() {
   if (slabs.size() > 1) {
      // merge: the previous call overflowed the first slab; replace all by one slab of the total
      MCHRG_CUDA(cudaStreamSynchronize(stream));
      size_t total = 0;
      for (auto& s : slabs) {
         total += s.size;
         cudaFree(s.base);
      }
      slabs.clear();
      Slab s;
      s.size = round_up((int64_t)total, 1 << 21);
      if (cudaMalloc(&s.base, s.size) != cudaSuccess) {
         cudaGetLastError();
         s.base = nullptr;
         s.size = 0;
      } else {
         slabs.push_back(s);
      }
   }
   for (auto& s : slabs) s.used = 0;
}

void* mchrg_b200_context::arena_alloc(size_t bytes) {
   bytes = (size_t)round_up((int64_t)(bytes ? bytes : 1), 256);
   for (auto& s : slabs) {
      if (s.used + bytes <= s.size) {
         void* p = s.base + s.used;
         s.used += bytes;
         return p;
      }
   }
   Slab s;
   s.size = (size_t)round_up((int64_t)(bytes > (size_t(64) << 20) ? bytes : (size_t(64) << 20)), 1 << 21);
   cudaError_t err = cudaMalloc(&s.base, s.size);
   if (err != cudaSuccess) {
      cudaGetLastError();
      fail(MCHRG_B200_ERR_ALLOC, "device allocation of %zu bytes failed: %s", s.size, cudaGetErrorString(err));
   }
   s.used = bytes;
   slabs.push_back(s);
   return s.base;
}

bool mchrg_b200_context::is_device_ptr(const void* p) {
   cudaPointerAttributes attr;
   cudaError_t err = cudaPointerGetAttributes(&attr, p);
   if (err != cudaSuccess) {
      cudaGetLastError();
      return false;
   }
   return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

void mchrg_b200_context::finish() {
   for (auto& c : copybacks) MCHRG_CUDA(cudaMemcpyAsync(c.host, c.dev, c.bytes, cudaMemcpyDeviceToHost, stream));
   copybacks.clear();
   MCHRG_CUDA(cudaStreamSynchronize(stream));
}

int mchrg_b200_context::nranks() const { return comm ? comm->nranks : 1; }
int mchrg_b200_context::rank() const { return comm ? comm->rank : 0; }

cudaEvent_t mchrg_b200_context::next_event() {
   if (events_used == events.size()) {
      cudaEvent_t e;
      MCHRG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      events.push_back(e);
   }
   return events[events_used++];
}

void* mchrg_b200_context::pinned_buf(size_t bytes) {
   if (bytes > pinned_size) {
      if (pinned) cudaFreeHost(pinned);
      pinned = nullptr;
      pinned_size = 0;
      size_t want = (size_t)round_up((int64_t)bytes, 4096);
      if (cudaMallocHost(&pinned, want) != cudaSuccess) {
         cudaGetLastError();
         fail(MCHRG_B200_ERR_ALLOC, "pinned host allocation of %zu bytes failed", want);
      }
      pinned_size = want;
   }
   return pinned;
}

extern "C" {

int mchrg_b200_version(void) { return 100; /* 0.1.0 */ }

const char* mchrg_b200_last_error(void) { return g_last_error.c_str(); }

int mchrg_b200_context_create(int device, mchrg_b200_context** out) {
   if (out == nullptr) {
      set_last_error("ctx output pointer is null");
      return MCHRG_B200_ERR_ARG;
   }
   *out = nullptr;
   int count = 0;
   cudaError_t err = cudaGetDeviceCount(&count);
   if (err != cudaSuccess || count == 0) {
      cudaGetLastError();
      set_last_error(std::string("no CUDA device available (libmchrg_b200 has no CPU fallback): ") +
                     (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0"));
      return MCHRG_B200_ERR_CUDA;
   }
   if (device < 0 || device >= count) {
      set_last_error("device index out of range");
      return MCHRG_B200_ERR_ARG;
   }
   auto* ctx = new mchrg_b200_context();
   try {
      ctx->device = device;
      ctx->use();
      cudaDeviceProp prop;
      MCHRG_CUDA(cudaGetDeviceProperties(&prop, device));
      if (prop.major < 10)
         fail(MCHRG_B200_ERR_CUDA, "device %d is sm_%d%d; libmchrg_b200 carries sm_100a code only", device,
              prop.major, prop.minor);
      ctx->sm_count = prop.multiProcessorCount;
      ctx->smem_optin = prop.sharedMemPerBlockOptin;
      int prio_lo = 0, prio_hi = 0;
      MCHRG_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      MCHRG_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_lo));
      MCHRG_CUDA(cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, prio_hi));
      // second priority stream: between the critical-path stream and the bulk work when the device has a level for it
      MCHRG_CUDA(cudaStreamCreateWithPriority(&ctx->stream_hi2, cudaStreamNonBlocking,
                                              (prio_lo - prio_hi >= 2) ? (prio_hi + prio_lo) / 2 : prio_hi));
      MCHRG_CUDA(cudaStreamCreateWithPriority(&ctx->stream_aux, cudaStreamNonBlocking, prio_lo));
      MCHRG_CUDA(cudaStreamCreateWithFlags(&ctx->stream_in, cudaStreamNonBlocking));
      MCHRG_CUDA(cudaStreamCreateWithFlags(&ctx->stream_out, cudaStreamNonBlocking));
      MCHRG_CUDA(cudaStreamCreateWithPriority(&ctx->stream_comm, cudaStreamNonBlocking, prio_hi));
      ctx->active = ctx->stream;
      {
         const std::vector<double2> tab = mchrg::eg::host_table();
         MCHRG_CUDA(cudaMalloc(&ctx->erf_tab, tab.size() * sizeof(double2)));
         MCHRG_CUDA(cudaMemcpy(ctx->erf_tab, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
      }
   } catch (const Error& e) {
      set_last_error(e.msg);
      delete ctx;
      return e.code;
   }
   *out = ctx;
   return 0;
}

void mchrg_b200_context_destroy(mchrg_b200_context* ctx) {
   if (!ctx) return;
   cudaSetDevice(ctx->device);
   if (ctx->stream) {
      cudaStreamSynchronize(ctx->stream);
      cudaStreamDestroy(ctx->stream);
   }
   if (ctx->stream_hi) {
      cudaStreamSynchronize(ctx->stream_hi);
      cudaStreamDestroy(ctx->stream_hi);
   }
   if (ctx->comm) {
      if (ctx->stream_comm) cudaStreamSynchronize(ctx->stream_comm);
      delete ctx->comm;
      ctx->comm = nullptr;
   }
   for (cudaStream_t st : {ctx->stream_aux, ctx->stream_in, ctx->stream_out, ctx->stream_hi2, ctx->stream_comm}) {
      if (st) {
         cudaStreamSynchronize(st);
         cudaStreamDestroy(st);
      }
   }
   ctx->batch_plan.reset();
   for (auto e : ctx->events) cudaEventDestroy(e);
   if (ctx->erf_tab) cudaFree(ctx->erf_tab);
   for (auto& s : ctx->slabs) cudaFree(s.base);
   if (ctx->pinned) cudaFreeHost(ctx->pinned);
   delete ctx;
}

int mchrg_b200_synchronize(mchrg_b200_context* ctx) {
   if (!ctx) return MCHRG_B200_ERR_ARG;
   cudaSetDevice(ctx->device);
   cudaError_t err = cudaStreamSynchronize(ctx->stream);
   if (err != cudaSuccess) {
      set_last_error(cudaGetErrorString(err));
      return MCHRG_B200_ERR_CUDA;
   }
   return 0;
}

void* mchrg_b200_stream(mchrg_b200_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
void* mchrg_b200_side_stream(mchrg_b200_context* ctx) { return ctx ? (void*)ctx->stream_aux : nullptr; }

uint64_t mchrg_b200_launch_count(const mchrg_b200_context* ctx) { return ctx ? ctx->launches : 0; }

int mchrg_b200_batch_trace(const mchrg_b200_context* ctx, double* ms, int32_t* launches) {
   if (!ctx || !ms || !launches) return MCHRG_B200_ERR_ARG;
   for (int k = 0; k < 3; ++k) {
      ms[k] = ctx->trace_ms[k];
      launches[k] = ctx->trace_launches[k];
   }
   return 0;
}

}  // extern "C"
