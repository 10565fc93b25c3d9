// group.cpp -- single-process multi-GPU context (wbem_params.n_gpus > 1).
//
// The reference is one process with one thread (source/main.cc:26); BEMProblem<3>::solve is called
// from FreeSurface (source/free_surface.cc:6099-6109) and cannot be launched once per GPU.  So ONE
// wbem_ctx handle owns P row blocks ("shards"): every shard is a complete per-rank context (its own
// device, stream, matrices rows, replicated vectors), and every collective entry point of the C ABI
// is handed to P persistent worker threads -- one per shard, so kernel launches for the P devices
// are issued in parallel and each shard's GMRES loop polls its own device.  The caller's thread
// only posts the job and waits.  The workers run the very code the one-process-per-GPU mode runs
// and meet in two places:
//   wbem_group_allgather  stream-ordered peer copies (events + a host barrier), used once per
//                         assembly / preconditioner set-up;
//   the fused mat-vec     k_bem_gemv stores its rows into every shard's gather buffer through
//                         peer access (cudaDeviceEnablePeerAccess; no IPC handles in one process).
// A shard that fails releases the barrier so its peers return instead of waiting forever.
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>

#include "internal.h"

struct WbemGroup
{
  int P = 0;
  std::vector<wbem_ctx *> shard; // shard[0] is the handle the caller holds
  std::vector<std::thread> workers;
  // job hand-off
  std::mutex m;
  std::condition_variable cv_job, cv_done;
  int (*fn)(wbem_ctx *, void *) = nullptr;
  void *arg = nullptr;
  unsigned long long job_id = 0;
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
  // barrier the workers meet in (abortable)
  std::mutex bm;
  std::condition_variable bcv;
  int b_count = 0;
  unsigned long long b_gen = 0;
  std::atomic<bool> failed{false};
  // exchange slots of the collectives
  std::vector<void *> xch;
  std::vector<cudaEvent_t> ev_ready, ev_done;
  std::atomic<int> vote{1};
};

static thread_local bool tl_in_worker = false;

bool wbem_group_forward(const wbem_ctx *ctx) { return ctx && ctx->group && !tl_in_worker; }
bool wbem_is_root(const wbem_ctx *ctx) { return !ctx->group || ctx->p.rank == 0; }
int wbem_group_size(const wbem_ctx *ctx) { return ctx->group ? ctx->group->P : 1; }
wbem_ctx *wbem_group_shard(const wbem_ctx *ctx, int r) { return ctx->group ? ctx->group->shard[r] : const_cast<wbem_ctx *>(ctx); }

static void worker_main(WbemGroup *g, int r)
{
  tl_in_worker = true;
  cudaSetDevice(g->shard[r]->dev);
  unsigned long long seen = 0;
  for (;;)
    {
      int (*fn)(wbem_ctx *, void *);
      void *arg;
      {
        std::unique_lock<std::mutex> lk(g->m);
        g->cv_job.wait(lk, [&] { return g->quit || g->job_id != seen; });
        if (g->quit) return;
        seen = g->job_id;
        fn = g->fn;
        arg = g->arg;
      }
      int rc;
      try
        {
          rc = fn(g->shard[r], arg);
        }
      catch (const std::exception &e)
        {
          g->shard[r]->err = std::string("exception in a shard worker: ") + e.what();
          rc = -4;
        }
      if (rc < 0)
        { // peers waiting for this shard in the barrier must not wait forever
          std::lock_guard<std::mutex> lk(g->bm);
          g->failed = true;
          g->bcv.notify_all();
        }
      {
        std::lock_guard<std::mutex> lk(g->m);
        g->rc[r] = rc;
        if (--g->pending == 0) g->cv_done.notify_all();
      }
    }
}

int wbem_group_run_impl(wbem_ctx *leader, int (*fn)(wbem_ctx *, void *), void *arg)
{
  WbemGroup *g = leader->group;
  {
    std::lock_guard<std::mutex> lk(g->bm);
    g->failed = false;
    g->b_count = 0;
  }
  std::unique_lock<std::mutex> lk(g->m);
  g->fn = fn;
  g->arg = arg;
  g->pending = g->P;
  g->job_id++;
  g->cv_job.notify_all();
  g->cv_done.wait(lk, [&] { return g->pending == 0; });
  // first fatal code wins (and its message moves to the handle the caller holds); else the
  // solver's "not converged" (> 0), which every shard reports alike
  int out = 0;
  for (int r = 0; r < g->P; ++r)
    if (g->rc[r] < 0 && g->rc[r] != -7)
      {
        out = g->rc[r];
        if (r != 0) leader->err = "row block " + std::to_string(r) + ": " + g->shard[r]->err;
        return out;
      }
  for (int r = 0; r < g->P; ++r)
    if (g->rc[r] < 0)
      {
        if (r != 0) leader->err = "row block " + std::to_string(r) + ": " + g->shard[r]->err;
        return g->rc[r];
      }
  for (int r = 0; r < g->P; ++r)
    if (g->rc[r] > 0) out = g->rc[r];
  return out;
}

int wbem_group_barrier(wbem_ctx *ctx)
{
  WbemGroup *g = ctx->group;
  if (!g) return 0;
  std::unique_lock<std::mutex> lk(g->bm);
  if (g->failed) WBEM_FAIL(ctx, -7, "another row block of this context failed");
  const unsigned long long gen = g->b_gen;
  if (++g->b_count == g->P)
    {
      g->b_count = 0;
      g->b_gen++;
      g->bcv.notify_all();
      return 0;
    }
  g->bcv.wait(lk, [&] { return g->b_gen != gen || g->failed.load(); });
  if (g->b_gen == gen) WBEM_FAIL(ctx, -7, "another row block of this context failed");
  return 0;
}

// every shard holds a buffer of P blocks with its own block filled: pull the other P-1 blocks
// from their owners (peer copies on this shard's stream, ordered by events)
int wbem_group_allgather(wbem_ctx *ctx, void *d_buf, size_t bytes_per_rank)
{
  WbemGroup *g = ctx->group;
  const int P = g->P, r = ctx->p.rank;
  cudaStream_t st = ctx->stream;
  g->xch[r] = d_buf;
  CUDA_OK(ctx, cudaEventRecord(g->ev_ready[r], st));
  int rc = wbem_group_barrier(ctx); // all pointers and "block ready" events are posted
  if (rc) return rc;
  for (int k = 1; k < P; ++k)
    {
      const int q = (r + k) % P; // staggered: not every shard pulls from shard 0 first
      CUDA_OK(ctx, cudaStreamWaitEvent(st, g->ev_ready[q], 0));
      CUDA_OK(ctx, cudaMemcpyAsync((char *)d_buf + bytes_per_rank * q, (const char *)g->xch[q] + bytes_per_rank * q,
                                   bytes_per_rank, cudaMemcpyDefault, st));
    }
  CUDA_OK(ctx, cudaEventRecord(g->ev_done[r], st));
  rc = wbem_group_barrier(ctx);
  if (rc) return rc;
  // nobody may overwrite its block again before every peer has read it
  for (int q = 0; q < P; ++q)
    if (q != r) CUDA_OK(ctx, cudaStreamWaitEvent(st, g->ev_done[q], 0));
  return 0;
}

// after every wbem_set_topology: let k_bem_gemv store into the peers' gather buffers
int wbem_group_p2p_setup(wbem_ctx *ctx)
{
  WbemGroup *g = ctx->group;
  const int P = g->P, r = ctx->p.rank;
  ctx->p2p_ready = false;
  if (P > WBEM_MAX_PEERS) return 0;
  g->xch[r] = ctx->d_p2p;
  g->vote = 1;
  int rc = wbem_group_barrier(ctx);
  if (rc) return rc;
  bool ok = ctx->d_p2p != nullptr;
  for (int q = 0; q < P && ok; ++q)
    {
      const int dq = g->shard[q]->dev;
      if (q != r && dq == ctx->dev && !ctx->p.fused_gather_on_shared_device) ok = false;
      if (q != r && dq != ctx->dev)
        {
          int can = 0;
          if (cudaDeviceCanAccessPeer(&can, ctx->dev, dq) != cudaSuccess || !can) ok = false;
          else
            {
              const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
              if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
              cudaGetLastError();
            }
        }
      ctx->peer_base[q] = (double *)g->xch[q];
      ctx->peer_opened[q] = false; // nothing to close: plain pointers of this process
    }
  if (!ok) g->vote = 0;
  rc = wbem_group_barrier(ctx);
  if (rc) return rc;
  ctx->p2p_ready = g->vote.load() == 1;
  return wbem_group_barrier(ctx); // the vote is read by everybody before the next one resets it
}

int wbem_group_create(const wbem_params *p, wbem_ctx **out, std::string *err)
{
  const int P = p->n_gpus;
  if (P > WBEM_MAX_PEERS)
    {
      *err = "n_gpus exceeds the supported number of row blocks";
      return -1;
    }
  WbemGroup *g = new WbemGroup();
  g->P = P;
  g->rc.assign(P, 0);
  g->xch.assign(P, nullptr);
  for (int r = 0; r < P; ++r)
    {
      wbem_params q = *p;
      q.n_gpus = 0;
      q.rank = r;
      q.world_size = P;
      q.device = p->devices[r] >= 0 ? p->devices[r] : r;
      wbem_ctx *s = nullptr;
      const int rc = wbem_create_single(&q, &s, err);
      if (rc)
        {
          for (wbem_ctx *t : g->shard) wbem_destroy_single(t);
          delete g;
          return rc;
        }
      s->p.n_gpus = P;
      s->p.fused_gather_on_shared_device = p->fused_gather_on_shared_device;
      s->group = g;
      g->shard.push_back(s);
    }
  g->ev_ready.resize(P);
  g->ev_done.resize(P);
  for (int r = 0; r < P; ++r)
    {
      cudaSetDevice(g->shard[r]->dev);
      cudaEventCreateWithFlags(&g->ev_ready[r], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&g->ev_done[r], cudaEventDisableTiming);
    }
  cudaSetDevice(g->shard[0]->dev);
  for (int r = 0; r < P; ++r) g->workers.emplace_back(worker_main, g, r);
  *out = g->shard[0];
  return 0;
}

int wbem_group_destroy(wbem_ctx *leader)
{
  WbemGroup *g = leader->group;
  // drain every device before any buffer a peer may still be writing into goes away
  wbem_group_run(leader, [](wbem_ctx *s) -> int {
    cudaSetDevice(s->dev);
    if (s->stream) cudaStreamSynchronize(s->stream);
    return 0;
  });
  {
    std::lock_guard<std::mutex> lk(g->m);
    g->quit = true;
    g->cv_job.notify_all();
  }
  for (std::thread &t : g->workers) t.join();
  for (int r = 0; r < g->P; ++r)
    {
      cudaSetDevice(g->shard[r]->dev);
      cudaEventDestroy(g->ev_ready[r]);
      cudaEventDestroy(g->ev_done[r]);
    }
  for (int r = g->P - 1; r >= 0; --r)
    {
      g->shard[r]->group = nullptr;
      wbem_destroy_single(g->shard[r]);
    }
  delete g;
  return 0;
}
