// a11: data-parallel gradient exchange.  The fine-tune step has exactly one collective — the sum of the flat fp32
// adapter + lm_head gradient bucket (6–14 M floats) over the ranks, DDP semantics (the 1/world factor is folded into
// the fused AdamW kernel's grad_scale).  It replaces torch's DDP reducer + ProcessGroupNCCL for the reference's
// trainable set (/root/reference/requirements.txt:1,75).  NCCL is bound at run time with dlopen so that libjl_b200.so
// loads on a box without it; only the five entry points used here are declared (ABI constants of NCCL 2.x).
#include <dlfcn.h>

#include <cstring>
#include <new>

#include <mutex>

#include "common.cuh"

namespace jl {

struct NcclId { char internal[JL_COMM_ID_BYTES]; };
using nccl_comm_t = void*;
constexpr int NCCL_FLOAT32 = 7;   // ncclDataType_t::ncclFloat32
constexpr int NCCL_SUM = 0;       // ncclRedOp_t::ncclSum

struct NcclApi {
  int (*get_unique_id)(NcclId*) = nullptr;
  int (*comm_init_rank)(nccl_comm_t*, int, NcclId, int) = nullptr;
  int (*all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*comm_destroy)(nccl_comm_t) = nullptr;
  const char* (*get_error_string)(int) = nullptr;
  bool ok = false;
  char why[160] = {0};
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the process already uses (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    snprintf(g_nccl.why, sizeof(g_nccl.why), "cannot load libnccl.so.2: %s", dlerror());
    return;
  }
  g_nccl.get_unique_id = reinterpret_cast<decltype(g_nccl.get_unique_id)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.comm_init_rank = reinterpret_cast<decltype(g_nccl.comm_init_rank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.all_reduce = reinterpret_cast<decltype(g_nccl.all_reduce)>(dlsym(h, "ncclAllReduce"));
  g_nccl.comm_destroy = reinterpret_cast<decltype(g_nccl.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.get_error_string = reinterpret_cast<decltype(g_nccl.get_error_string)>(dlsym(h, "ncclGetErrorString"));
  g_nccl.ok = g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.all_reduce && g_nccl.comm_destroy && g_nccl.get_error_string;
  if (!g_nccl.ok) snprintf(g_nccl.why, sizeof(g_nccl.why), "libnccl.so.2 lacks a required symbol");
}

static int need_nccl() {
  std::call_once(g_nccl_once, load_nccl);
  JL_REQUIRE(g_nccl.ok, JL_EUNSUPPORTED, "comm: %s", g_nccl.why);
  return JL_OK;
}

}  // namespace jl

struct jl_comm {
  jl::nccl_comm_t nccl;
  int32_t rank, world;
};

#define JL_NCCL(call, what)                                                                          \
  do {                                                                                               \
    const int r__ = (call);                                                                          \
    JL_REQUIRE(r__ == 0, JL_ECUDA, "comm: %s failed: %s", what, jl::g_nccl.get_error_string(r__));   \
  } while (0)

extern "C" {

int jl_comm_unique_id(void* id_out) {
  JL_REQUIRE(id_out != nullptr, JL_EINVAL, "comm: null id buffer");
  if (int rc = jl::need_nccl()) return rc;
  jl::NcclId id;
  JL_NCCL(jl::g_nccl.get_unique_id(&id), "ncclGetUniqueId");
  memcpy(id_out, id.internal, JL_COMM_ID_BYTES);
  return JL_OK;
}

int jl_comm_init(const void* id, int32_t rank, int32_t world, jl_comm** out) {
  JL_REQUIRE(id != nullptr && out != nullptr, JL_EINVAL, "comm: null argument");
  JL_REQUIRE(world >= 1 && rank >= 0 && rank < world, JL_EINVAL, "comm: bad rank %d of %d", rank, world);
  if (int rc = jl::check_device()) return rc;
  if (int rc = jl::need_nccl()) return rc;
  jl::NcclId nid;
  memcpy(nid.internal, id, JL_COMM_ID_BYTES);
  jl::nccl_comm_t c = nullptr;
  JL_NCCL(jl::g_nccl.comm_init_rank(&c, world, nid, rank), "ncclCommInitRank");
  jl_comm* comm = new (std::nothrow) jl_comm{c, rank, world};
  if (comm == nullptr) {
    jl::g_nccl.comm_destroy(c);
    JL_REQUIRE(false, JL_EINVAL, "comm: out of host memory");
  }
  *out = comm;
  return JL_OK;
}

int jl_comm_allreduce(jl_comm* comm, float* buf, size_t n_f32, void* stream) {
  JL_REQUIRE(comm != nullptr && comm->nccl != nullptr, JL_EINVAL, "comm: null communicator");
  JL_REQUIRE(buf != nullptr || n_f32 == 0, JL_EINVAL, "comm: null buffer");
  if (n_f32 == 0 || comm->world == 1) return JL_OK;      // a single rank already holds the sum
  JL_NCCL(jl::g_nccl.all_reduce(buf, buf, n_f32, jl::NCCL_FLOAT32, jl::NCCL_SUM, comm->nccl, static_cast<cudaStream_t>(stream)),
          "ncclAllReduce");
  return JL_OK;
}

int jl_comm_rank(const jl_comm* comm, int32_t* rank, int32_t* world) {
  JL_REQUIRE(comm != nullptr, JL_EINVAL, "comm: null communicator");
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  return JL_OK;
}

int jl_comm_destroy(jl_comm* comm) {
  if (comm == nullptr) return JL_OK;
  if (comm->nccl != nullptr && jl::g_nccl.ok) JL_NCCL(jl::g_nccl.comm_destroy(comm->nccl), "ncclCommDestroy");
  delete comm;
  return JL_OK;
}

}  // extern "C"
