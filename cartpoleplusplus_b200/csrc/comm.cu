// Data-parallel gradient exchange inside the library: ONE sum all-reduce of the flat gradient buffer per
// grad-step, issued from C++ on the step's own streams so that it is captured into the step's CUDA graph and overlaps
// the last kernels of the backward pass - instead of a host-issued torch.distributed call between two C-ABI calls.
//
// Two transports behind one interface (comm.cuh):
//  * P2P (default on one NVSwitch box): our own one-shot all-reduce over NVLink peer memory.  The 0.93 MB buffer is latency
//    bound - a library all-reduce costs ~25 us at 2 ranks whatever the size - so the exchange is ONE kernel after the last
//    gradient kernel: every CTA owns a slice of the buffer, PUSHES it into the slot this rank has in every peer's block
//    (128-bit peer stores over NVLink), publishes a release flag per peer, acquires the peers' flags for the same slice and
//    sums the world slots LOCALLY in rank order (identical bits on every replica).  Flow control is an ack per rank and step,
//    so one buffer set suffices; every spin has a timeout that traps instead of hanging the GPU.
//    Measured alternatives (profiles/r3/dp_allreduce.md): pushing everything but conv1's gradients early, next to the conv1
//    weight-gradient kernel - by DMA (cudaMemcpyAsync nodes) or as grouped NCCL calls - made the step SLOWER at 2 ranks
//    (the concurrent traffic delays the weight-gradient kernel by more than the exchange costs when it runs alone).
//  * NCCL (cpp_*_comm_init with a unique id): one ncclAllReduce of the whole buffer at the same place.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a Python process that imported torch this resolves to the
// NCCL torch itself loaded, 2.28.9 here; otherwise to the system library) through the handful of entry points below;
// their signatures and the two enum values are those of nccl.h 2.x (ncclFloat32 = 7, ncclSum = 0).
#include <dlfcn.h>
#include <string.h>
#include "comm.cuh"

namespace cpp {

namespace {
struct UniqueId { char internal[128]; };
typedef int (*fn_get_unique_id)(UniqueId*);
typedef int (*fn_comm_init_rank)(void**, int, UniqueId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_get_version)(int*);

struct Api {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_error_string error_string = nullptr;
  fn_get_version get_version = nullptr;
  bool tried = false, ok = false;
};
Api g_api;

int load_api() {
  if (g_api.tried) {
    if (!g_api.ok) { set_error("NCCL is not available in this process (libnccl.so.2 could not be bound)"); return CPP_ERR_NCCL; }
    return CPP_OK;
  }
  g_api.tried = true;
  // RTLD_NOLOAD first: the copy torch already mapped (one NCCL per process - two would not share CUDA IPC state)
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) { set_error("dlopen(libnccl.so.2): %s", dlerror()); return CPP_ERR_NCCL; }
  g_api.handle = h;
#define BIND(member, sym, type)                                                                  \
  g_api.member = reinterpret_cast<type>(dlsym(h, sym));                                          \
  if (g_api.member == nullptr) { set_error("libnccl.so.2 has no symbol %s", sym); return CPP_ERR_NCCL; }
  BIND(get_unique_id, "ncclGetUniqueId", fn_get_unique_id)
  BIND(comm_init_rank, "ncclCommInitRank", fn_comm_init_rank)
  BIND(comm_destroy, "ncclCommDestroy", fn_comm_destroy)
  BIND(all_reduce, "ncclAllReduce", fn_all_reduce)
  BIND(group_start, "ncclGroupStart", fn_group)
  BIND(group_end, "ncclGroupEnd", fn_group)
  BIND(error_string, "ncclGetErrorString", fn_error_string)
  BIND(get_version, "ncclGetVersion", fn_get_version)
#undef BIND
  g_api.ok = true;
  return CPP_OK;
}

#define CPP_CHECK_NCCL(expr)                                                                              \
  do {                                                                                                    \
    const int _r = (expr);                                                                                \
    if (_r != 0) { set_error("%s:%d NCCL error %d: %s", __FILE__, __LINE__, _r, g_api.error_string(_r)); return CPP_ERR_NCCL; } \
  } while (0)
}  // namespace

int comm_unique_id(void* out128) {
  CPP_TRY(load_api());
  UniqueId id;
  CPP_CHECK_NCCL(g_api.get_unique_id(&id));
  memcpy(out128, id.internal, sizeof(id.internal));
  return CPP_OK;
}

int comm_version(int* v) {
  CPP_TRY(load_api());
  CPP_CHECK_NCCL(g_api.get_version(v));
  return CPP_OK;
}

int Comm::init(int rank_, int world_, const void* id128) {
  CPP_REQUIRE(world_ >= 1 && rank_ >= 0 && rank_ < world_, "comm: rank %d of %d", rank_, world_);
  destroy();
  rank = rank_; world = world_;
  if (world == 1) return CPP_OK;
  CPP_REQUIRE(id128 != nullptr, "comm: null unique id");
  CPP_TRY(load_api());
  UniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  CPP_CHECK_NCCL(g_api.comm_init_rank(&comm, world, id, rank));
  mode = NCCL;
  return CPP_OK;
}

void Comm::destroy() {
  if (comm != nullptr && g_api.ok) g_api.comm_destroy(comm);
  comm = nullptr;
  if (local_block != nullptr) {
    cudaDeviceSynchronize();
    for (int r = 0; r < world; ++r)
      if (r != rank && peer_block[r] != nullptr) cudaIpcCloseMemHandle(peer_block[r]);
    cudaFree(local_block);
  }
  local_block = nullptr;
  for (auto& p : peer_block) p = nullptr;
  world = 1; rank = 0; mode = NONE;
}

int Comm::all_reduce_sum(float* const* ptrs, const int64_t* counts, int n, cudaStream_t s) const {
  if (!active()) return CPP_OK;
  CPP_TRY(load_api());
  // several disjoint ranges of the gradient buffer in ONE grouped call (one NCCL kernel)
  if (n > 1) CPP_CHECK_NCCL(g_api.group_start());
  for (int i = 0; i < n; ++i)
    if (counts[i] > 0) CPP_CHECK_NCCL(g_api.all_reduce(ptrs[i], ptrs[i], (size_t)counts[i], /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, s));
  if (n > 1) CPP_CHECK_NCCL(g_api.group_end());
  ++g_launch_count;
  return CPP_OK;
}

// ------------------------------------------------------------------------------------------ P2P over NVLink peer memory
static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

P2PBlock Comm::view(int r) const {
  char* b = reinterpret_cast<char*>(peer_block[r]);
  P2PBlock v;
  size_t off = 0;
  v.slots = reinterpret_cast<float*>(b); off += al256((size_t)world * n_pad * sizeof(float));
  v.flags = reinterpret_cast<uint32_t*>(b + off); off += al256((size_t)world * kP2PCtas * sizeof(uint32_t));
  v.acks = reinterpret_cast<uint32_t*>(b + off); off += al256((size_t)kMaxRanks * sizeof(uint32_t));
  v.epoch = reinterpret_cast<uint32_t*>(b + off);
  v.done = v.epoch + 1;
  return v;
}

int Comm::p2p_prepare(int rank_, int world_, int64_t n_floats, void* handle_out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CPP_REQUIRE(world_ >= 1 && world_ <= kMaxRanks && rank_ >= 0 && rank_ < world_, "p2p: rank %d of %d (at most %d)", rank_, world_, kMaxRanks);
  CPP_REQUIRE(n_floats > 0 && handle_out64 != nullptr, "p2p: bad arguments");
  destroy();
  rank = rank_; world = world_;
  n_pad = round_up(n_floats, 64);
  block_bytes = al256((size_t)world * n_pad * sizeof(float)) + al256((size_t)world * kP2PCtas * sizeof(uint32_t)) +
                al256((size_t)kMaxRanks * sizeof(uint32_t)) + 256;
  CPP_CHECK_CUDA(cudaMalloc(&local_block, block_bytes));
  CPP_CHECK_CUDA(cudaMemset(local_block, 0, block_bytes));
  CPP_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CPP_CHECK_CUDA(cudaIpcGetMemHandle(&h, local_block));
  memcpy(handle_out64, &h, sizeof(h));
  peer_block[rank] = local_block;
  return CPP_OK;
}

int Comm::p2p_connect(const void* handles) {
  CPP_REQUIRE(local_block != nullptr && handles != nullptr, "p2p_connect before p2p_prepare");
  const char* hb = reinterpret_cast<const char*>(handles);
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
    CPP_CHECK_CUDA(cudaIpcOpenMemHandle(&peer_block[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  mode = world > 1 ? P2P : NONE;
  return CPP_OK;
}

namespace {
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until *p has reached e (epochs only grow); a peer that never arrives must fail loudly, not hang the GPU (~4 s)
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t e) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(p) - e) < 0) {
    if (clock64() - t0 > (8ll << 30)) __trap();
  }
}

struct P2PArgs {
  float* grads; int64_t n;             // this rank's flat gradient buffer (n floats, n % 4 == 0)
  P2PBlock peer[kMaxRanks];            // every rank's block as mapped here; peer[rank] is ours
  int rank, world;
  int64_t n_pad;
};

__global__ void __launch_bounds__(512) p2p_all_reduce_kernel(const __grid_constant__ P2PArgs A) {
  const P2PBlock me = A.peer[A.rank];
  const int tid = threadIdx.x, c = blockIdx.x;
  const uint32_t e = *me.epoch + 1;                                   // this step (every CTA reads it before the last one bumps it)
  const int64_t per = ((A.n / 4 + gridDim.x - 1) / gridDim.x) * 4;
  const int64_t c0 = (int64_t)c * per, c1 = c0 + per < A.n ? c0 + per : A.n;
  // 0. the peers have finished reading what we pushed last step: their slots may be overwritten
  if (tid < A.world) spin_until(me.acks + tid, e - 1);
  __syncthreads();
  // 1. push this CTA's slice into slot[rank] of every block (ours included), 128-bit stores
  for (int64_t i = c0 + 4 * tid; i < c1; i += 4 * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(A.grads + i);
    for (int r = 0; r < A.world; ++r) *reinterpret_cast<float4*>(A.peer[r].slots + (int64_t)A.rank * A.n_pad + i) = v;
  }
  // 2. publish: this CTA's slice of step e is complete in every peer's memory.  The CTA barrier orders every thread's stores
  // before the releasing threads, whose system-scope release is cumulative over them (one fence per peer, not one per thread:
  // 32 k membar.sys at once cost more than the whole exchange)
  __syncthreads();
  if (tid < A.world) st_release_sys(A.peer[tid].flags + A.rank * kP2PCtas + c, e);
  // 3. wait for the same slice of every peer
  if (tid < A.world) spin_until(me.flags + tid * kP2PCtas + c, e);
  __syncthreads();
  // 4. sum the slots in rank order (local memory; L1 may hold lines of an earlier step: cache-volatile loads)
  for (int64_t i = c0 + 4 * tid; i < c1; i += 4 * blockDim.x) {
    float4 acc = __ldcv(reinterpret_cast<const float4*>(me.slots + i));
    for (int r = 1; r < A.world; ++r) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(me.slots + (int64_t)r * A.n_pad + i));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(A.grads + i) = acc;
  }
  // 5. the last CTA to finish acks the step to every peer and advances the epoch
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(me.done, 1u) == gridDim.x - 1) {
      *me.done = 0;
      for (int r = 0; r < A.world; ++r) st_release_sys(A.peer[r].acks + A.rank, e);
      *me.epoch = e;
      __threadfence();
    }
  }
}
}  // namespace

int Comm::all_reduce(float* grads, int64_t n, cudaStream_t s) const {
  if (!active()) return CPP_OK;
  if (mode == NCCL) {
    float* p[1] = {grads}; const int64_t cnt[1] = {n};
    return all_reduce_sum(p, cnt, 1, s);
  }
  CPP_REQUIRE(n % 4 == 0 && n <= n_pad && ((uintptr_t)grads & 15) == 0, "p2p: gradient buffer of %lld floats (slot %lld)", (long long)n, (long long)n_pad);
  P2PArgs A{};
  A.grads = grads; A.n = n; A.rank = rank; A.world = world; A.n_pad = n_pad;
  for (int r = 0; r < world; ++r) A.peer[r] = view(r);
  p2p_all_reduce_kernel<<<kP2PCtas, 512, 0, s>>>(A);
  CPP_CHECK_LAUNCH();
  return CPP_OK;
}

}  // namespace cpp
