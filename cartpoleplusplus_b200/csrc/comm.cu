// Data-parallel gradient exchange inside the library: ONE sum all-reduce of the flat gradient buffer per
// grad-step, issued from C++ on the step's own streams so that it is captured into the step's CUDA graph and overlaps
// the last kernels of the backward pass - instead of a host-issued torch.distributed call between two C-ABI calls.
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
  return CPP_OK;
}

void Comm::destroy() {
  if (comm != nullptr && g_api.ok) g_api.comm_destroy(comm);
  comm = nullptr; world = 1; rank = 0;
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

}  // namespace cpp
