// Collectives of the multi-GPU single-system path: panel broadcasts of the block-cyclic
// Cholesky, all-reduce of the row-pass partial sums, all-gather of row-sharded results.
// Two backends behind one interface:
//   NCCL   one process per GPU (torchrun / MPI-style launch); libnccl.so.2 is dlopen'ed at run time, the
//          communicator is built from a caller-supplied ncclUniqueId (mchrg_b200_comm_init_nccl)
//   local  one process, one host thread per context (several GPUs, or several contexts on one GPU for the
//          tests): stream-ordered peer copies with a host rendezvous per collective (mchrg_b200_comm_init_local)
// Every rank must issue the same sequence of collectives.  All operations are enqueued on the stream given
// (the library uses the context's dedicated communication stream) and return immediately.
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

namespace mchrg {

struct Comm {
   int rank = 0, nranks = 1;
   virtual ~Comm() {}
   virtual const char* backend() const = 0;
   // buf (bytes) of rank `root` replaces buf on every other rank
   virtual void bcast(void* buf, size_t bytes, int root, cudaStream_t s) = 0;
   // element-wise sum over ranks, in place, same summation order (rank 0, 1, ...) on every rank for the local
   // backend; NCCL's own order otherwise
   virtual void allreduce_sum(double* buf, size_t count, cudaStream_t s) = 0;
   // recv[r * bytes .. (r + 1) * bytes) = send of rank r
   virtual void allgather(const void* send, void* recv, size_t bytes, cudaStream_t s) = 0;
   // several broadcasts as one group (one launch with NCCL); roots / buffers may differ
   virtual void group_begin() {}
   virtual void group_end() {}
};

}  // namespace mchrg
