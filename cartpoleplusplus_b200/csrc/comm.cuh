// NCCL communicator of one data-parallel replica (comm.cu).
#pragma once
#include "common.cuh"

namespace cpp {

int comm_unique_id(void* out128);      // ncclGetUniqueId (rank 0; the host broadcasts the 128 bytes)
int comm_version(int* v);

struct Comm {
  void* comm = nullptr;                // ncclComm_t
  int rank = 0, world = 1;
  bool active() const { return comm != nullptr && world > 1; }
  int init(int rank, int world, const void* id128);
  void destroy();
  // in-place sum over all ranks of n disjoint float ranges, enqueued on s as one grouped NCCL call (capturable)
  int all_reduce_sum(float* const* ptrs, const int64_t* counts, int n, cudaStream_t s) const;
  ~Comm() { destroy(); }
};

}  // namespace cpp
