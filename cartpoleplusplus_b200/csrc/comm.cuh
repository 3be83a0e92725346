// Gradient exchange of one data-parallel replica (comm.cu): our own all-reduce over NVLink peer memory, or NCCL.
#pragma once
#include "common.cuh"

namespace cpp {

int comm_unique_id(void* out128);      // ncclGetUniqueId (rank 0; the host broadcasts the 128 bytes)
int comm_version(int* v);

constexpr int kMaxRanks = 8;           // one NVSwitch box
constexpr int kP2PCtas = 64;           // CTAs of the all-reduce kernel; each owns a contiguous slice of the buffer and its own flags

// Peer-memory exchange state.  Every rank owns ONE cudaMalloc'ed block that its peers map through CUDA IPC:
//   slots [world][n_pad] floats  - slot r receives rank r's gradient buffer (peer stores of r's all-reduce kernel)
//   flags [world][kP2PCtas] u32  - flags[r][c] = e: CTA c of rank r has published its slice of step e
//   acks  [world] u32            - acks[r] = e: rank r has finished READING step e out of its own slots (ours may be overwritten)
//   epoch, done u32              - completed steps of this rank; CTA completion counter
struct P2PBlock {
  float* slots; uint32_t* flags; uint32_t* acks; uint32_t* epoch; uint32_t* done;
};

struct Comm {
  enum Mode { NONE = 0, NCCL = 1, P2P = 2 };
  int mode = NONE;
  void* comm = nullptr;                // ncclComm_t
  int rank = 0, world = 1;
  // ---- P2P
  void* local_block = nullptr;         // this rank's block (cudaMalloc)
  void* peer_block[kMaxRanks] = {};    // every rank's block as mapped here (peer_block[rank] == local_block)
  int64_t n_pad = 0;                   // floats per slot
  size_t block_bytes = 0;

  bool active() const { return world > 1 && mode != NONE; }
  const void* key() const { return mode == NCCL ? comm : local_block; }
  int init(int rank, int world, const void* id128);                       // NCCL
  // P2P, two phases around a host-side exchange of the 64-byte IPC handles (torch.distributed is only the rendezvous)
  int p2p_prepare(int rank, int world, int64_t n_floats, void* handle_out64);
  int p2p_connect(const void* handles /* world x 64 bytes */);
  void destroy();
  // in-place sum of the flat gradient buffer over all replicas, enqueued on s (capturable).  P2P: one kernel (comm.cu);
  // NCCL: one ncclAllReduce.  Bit-identical results on every replica.
  int all_reduce(float* grads, int64_t n, cudaStream_t s) const;
  // in-place sum over all ranks of n disjoint float ranges, enqueued on s as one grouped NCCL call
  int all_reduce_sum(float* const* ptrs, const int64_t* counts, int n, cudaStream_t s) const;
  P2PBlock view(int r) const;
  ~Comm() { destroy(); }
};

}  // namespace cpp
