// comm.cu — C1: the one exchange step of the engine.  Bundle adjustment shards points (and so
// observations) over ranks; per Gauss-Newton iteration every rank contributes its partial reduced
// camera system and the sum is needed everywhere:  all-reduce(sum) of [S | g | diag(Hcc)] (float32,
// (6C)^2 + 12C values) and of the cost scalar, over NCCL (NVLink 5 / NVSwitch).  Nothing else in the
// hot path communicates: matching, triangulation and PnP shard over pairs / scenes without exchange.
//
// NCCL is bound at run time with dlopen so that the library has no link-time dependency on a
// particular libnccl (a process that already imported torch has torch's bundled libnccl.so.2
// loaded and that one is reused).
#include <dlfcn.h>

#include "ba.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
enum { NCCL_FLOAT = 7, NCCL_DOUBLE = 8, NCCL_SUM = 0 };   // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return SFM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    sfm_set_error("NCCL not found: %s", dlerror());
    return SFM_ERR_NCCL;
  }
#define LOAD(field, sym)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym));       \
  if (!g_nccl.field) { sfm_set_error("NCCL symbol %s missing", sym); return SFM_ERR_NCCL; }
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllReduce, "ncclAllReduce");
  LOAD(GroupStart, "ncclGroupStart");
  LOAD(GroupEnd, "ncclGroupEnd");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  g_nccl.handle = h;
  return SFM_OK;
}

#define SFM_NCCL(call)                                                                           \
  do {                                                                                           \
    int _r = (call);                                                                             \
    if (_r != 0) {                                                                               \
      sfm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r));     \
      return SFM_ERR_NCCL;                                                                       \
    }                                                                                            \
  } while (0)

}  // namespace

extern "C" int sfm_nccl_unique_id(void* out128) {
  SFM_REQUIRE(out128, "sfm_nccl_unique_id: null buffer");
  SFM_TRY(load_nccl());
  nccl_uid_t id;
  SFM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return SFM_OK;
}

extern "C" int sfm_ba_comm_init(sfm_ba* ba, const void* unique_id128, int rank, int world) {
  SFM_REQUIRE(ba && unique_id128 && world >= 1 && rank >= 0 && rank < world, "sfm_ba_comm_init: bad argument");
  SFM_REQUIRE(!ba->comm, "sfm_ba_comm_init: communicator already attached");
  SFM_TRY(load_nccl());
  SFM_CUDA(cudaSetDevice(ba->ctx->device));
  nccl_uid_t id;
  memcpy(&id, unique_id128, sizeof(id));
  nccl_comm_t comm = nullptr;
  SFM_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ba->comm = comm;
  ba->rank = rank;
  ba->world = world;
  return SFM_OK;
}

extern "C" int sfm_ba_comm_destroy(sfm_ba* ba) {
  if (!ba || !ba->comm) return SFM_OK;
  cudaStreamSynchronize(ba->ctx->stream);
  g_nccl.CommDestroy((nccl_comm_t)ba->comm);
  ba->comm = nullptr;
  ba->world = 1;
  ba->rank = 0;
  return SFM_OK;
}

int sfm_ba_allreduce_system(sfm_ba* ba) {
  if (!ba->comm || ba->world == 1) return SFM_OK;
  cudaStream_t st = ba->ctx->stream;
  SFM_NCCL(g_nccl.GroupStart());
  SFM_NCCL(g_nccl.AllReduce(ba->S, ba->S, ba->sys_count, NCCL_FLOAT, NCCL_SUM, (nccl_comm_t)ba->comm, st));
  SFM_NCCL(g_nccl.AllReduce(ba->scal, ba->scal, 1, NCCL_DOUBLE, NCCL_SUM, (nccl_comm_t)ba->comm, st));
  SFM_NCCL(g_nccl.GroupEnd());
  return SFM_OK;
}

int sfm_ba_allreduce_scalars(sfm_ba* ba) {
  if (!ba->comm || ba->world == 1) return SFM_OK;
  SFM_NCCL(g_nccl.AllReduce(ba->scal + 1, ba->scal + 1, 2, NCCL_DOUBLE, NCCL_SUM, (nccl_comm_t)ba->comm, ba->ctx->stream));
  return SFM_OK;
}
