// comm.cpp -- NCCL plumbing for row-sharded runs (one process per GPU).
//
// The only collective of the path is the all-gather of each rank's block of result rows after
// a matrix-vector product (and, once per assembly / solve, of alpha and of the band rows).
// NCCL is loaded with dlopen so that libwbem.so itself has no link-time dependency on it:
// single-GPU users never touch it, and inside a PyTorch process the already-loaded
// libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "internal.h"

typedef struct
{
  char internal[128];
} nccl_uid_t;
typedef int (*fn_get_uid)(nccl_uid_t *);
typedef int (*fn_comm_init_rank)(void **, int, nccl_uid_t, int);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_gather)(const void *, void *, size_t, int /*dtype*/, void *, cudaStream_t);
typedef const char *(*fn_err_string)(int);

struct NcclApi
{
  void *handle = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init_rank init_rank = nullptr;
  fn_comm_destroy destroy = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_err_string err_string = nullptr;
};

static NcclApi *load_nccl(std::string *err)
{
  static NcclApi api;
  if (api.handle) return &api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names)
    {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
  if (!api.handle)
    {
      *err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return nullptr;
    }
  api.get_uid = (fn_get_uid)dlsym(api.handle, "ncclGetUniqueId");
  api.init_rank = (fn_comm_init_rank)dlsym(api.handle, "ncclCommInitRank");
  api.destroy = (fn_comm_destroy)dlsym(api.handle, "ncclCommDestroy");
  api.all_gather = (fn_all_gather)dlsym(api.handle, "ncclAllGather");
  api.err_string = (fn_err_string)dlsym(api.handle, "ncclGetErrorString");
  if (!api.get_uid || !api.init_rank || !api.destroy || !api.all_gather)
    {
      *err = "libnccl.so.2 lacks a required symbol";
      api.handle = nullptr;
      return nullptr;
    }
  return &api;
}

int wbem_nccl_unique_id(void *id128, std::string *err)
{
  NcclApi *api = load_nccl(err);
  if (!api) return -5;
  nccl_uid_t id;
  const int rc = api->get_uid(&id);
  if (rc)
    {
      *err = std::string("ncclGetUniqueId: ") + (api->err_string ? api->err_string(rc) : "error");
      return -5;
    }
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int wbem_nccl_init(wbem_ctx *ctx, const void *id128)
{
  if (ctx->p.world_size <= 1) return 0;
  std::string err;
  NcclApi *api = load_nccl(&err);
  if (!api) WBEM_FAIL(ctx, -5, "%s", err.c_str());
  ctx->nccl = api;
  nccl_uid_t id;
  memcpy(&id, id128, sizeof(id));
  void *comm = nullptr;
  const int rc = api->init_rank(&comm, ctx->p.world_size, id, ctx->p.rank);
  if (rc) WBEM_FAIL(ctx, -5, "ncclCommInitRank: %s", api->err_string ? api->err_string(rc) : "error");
  ctx->nccl_comm = comm;
  return 0;
}

void wbem_nccl_destroy(wbem_ctx *ctx)
{
  if (ctx->nccl && ctx->nccl_comm) ctx->nccl->destroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

int wbem_nccl_allgather(wbem_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank)
{
  // ncclInt8 = 0 (ncclChar): byte-wise gather keeps one entry point for doubles and band rows
  const int rc = ctx->nccl->all_gather(send, recv, bytes_per_rank, 0, ctx->nccl_comm, ctx->stream);
  if (rc) WBEM_FAIL(ctx, -5, "ncclAllGather: %s", ctx->nccl->err_string ? ctx->nccl->err_string(rc) : "error");
  return 0;
}
