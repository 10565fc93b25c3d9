// operator.cu -- the dense operator applications of the GMRES solve.
//
//   BEMProblem<3>::vmult        dst = -D (s.x) + N (o.x) + alpha.o.x   (source/bem_problem.cc:620-670)
//   BEMProblem<3>::compute_rhs  dst = -(N (s.t) + alpha.s.t) + D (o.t) (source/bem_problem.cc:673-707)
//   ConstrainedOperator::vmult  constrained rows overwritten          (include/constrained_matrix.h:73-86)
//
// s = surface_nodes, o = other_nodes.  Because the multiplier of N (resp. D) is exactly zero
// wherever its mask is zero, only the 64-column chunks of each matrix whose mask is non-zero
// are streamed (chunk lists built in wbem_set_masks): with complementary 0/1 masks one
// operator application reads ~8 N^2 bytes instead of the reference's 16 N^2.
// Each warp owns 4 rows: the x chunk is loaded once and reused for 4 streamed 128-bit
// matrix loads per lane.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "internal.h"

#define GEMV_WARPS 8   // warps per CTA
#define GEMV_MAXR 5    // rows one warp streams in one pass

__device__ __forceinline__ double2 ld_stream(const double *p)
{
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// acc[r] += sum over the listed 64-column chunks of M[rows[r], :] . x.  NR rows share every x
// chunk; NC chunks are in flight per iteration, NR * NC ~ 8..10 independent 128-bit loads per
// lane whatever the number of rows a warp owns (few rows per warp on row-sharded runs).
template <int NR, int NC>
__device__ __forceinline__ void gemv_pass(const double *__restrict__ M, const double *__restrict__ x,
                                          const uint32_t *__restrict__ list, int n, uint32_t ld,
                                          const uint32_t rows[GEMV_MAXR], int lane, double acc[GEMV_MAXR])
{
  const double *base[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) base[r] = M + (size_t)rows[r] * ld + lane * 2;
  int i = 0;
  for (; i + NC <= n; i += NC)
    {
      uint32_t c[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) c[k] = list[i + k] * 64;
      double2 m[NC][NR];
#pragma unroll
      for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int r = 0; r < NR; ++r) m[k][r] = ld_stream(base[r] + c[k]);
      double2 xv[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) xv[k] = *reinterpret_cast<const double2 *>(x + c[k] + lane * 2);
#pragma unroll
      for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int r = 0; r < NR; ++r)
          {
            acc[r] = fma(m[k][r].x, xv[k].x, acc[r]);
            acc[r] = fma(m[k][r].y, xv[k].y, acc[r]);
          }
    }
  for (; i < n; ++i)
    {
      const uint32_t c0 = list[i] * 64;
      const double2 x0 = *reinterpret_cast<const double2 *>(x + c0 + lane * 2);
#pragma unroll
      for (int r = 0; r < NR; ++r)
        {
          const double2 m0 = ld_stream(base[r] + c0);
          acc[r] = fma(m0.x, x0.x, acc[r]);
          acc[r] = fma(m0.y, x0.y, acc[r]);
        }
    }
}

struct GemvArgs
{
  const int *stop; // GMRES iterations enqueued ahead: non-zero = skip the streaming (the completion protocol still runs)
  const double *M1, *x1, *M2, *x2, *alpha, *xdiag;
  const uint32_t *list1, *list2;
  int n1, n2;
  double s1, s2, sdiag;
  uint32_t ld, nloc, row0;
  const uint32_t *row_list; // local rows to compute (constrained rows are skipped), or null
  uint32_t n_rows;          // entries of row_list (= nloc when null)
  double *y;
  // fused all-gather over NVLink peer memory (n_peers > 1): every result row is stored straight
  // into the gather buffer of every rank; the last CTA to finish raises this rank's flag there
  int n_peers, rank;
  double *peer_base[WBEM_MAX_PEERS];
  size_t buf_off, flag_off; // in doubles from peer_base: this epoch's buffer / the flag array
  unsigned long long epoch, done_target;
  unsigned long long *done_counter;
};

__device__ __forceinline__ void gemv_store_fwd(const GemvArgs &a, uint32_t lrow, double v, int lane);

template <int NR, int NC>
__device__ __forceinline__ void gemv_rows(const GemvArgs &a, uint32_t r0, int lane)
{
  double a1[GEMV_MAXR], a2[GEMV_MAXR];
#pragma unroll
  for (int r = 0; r < GEMV_MAXR; ++r) a1[r] = a2[r] = 0.0;
  uint32_t rows[GEMV_MAXR];
#pragma unroll
  for (int r = 0; r < NR; ++r) rows[r] = a.row_list ? a.row_list[r0 + r] : r0 + r;
  gemv_pass<NR, NC>(a.M1, a.x1, a.list1, a.n1, a.ld, rows, lane, a1);
  gemv_pass<NR, NC>(a.M2, a.x2, a.list2, a.n2, a.ld, rows, lane, a2);
#pragma unroll
  for (int r = 0; r < NR; ++r)
    {
      double v = a.s1 * a1[r] + a.s2 * a2[r];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      gemv_store_fwd(a, rows[r], v, lane);
    }
}

__device__ __forceinline__ void gemv_store(const GemvArgs &a, uint32_t lrow, double v, int lane);
__device__ __forceinline__ void gemv_store_fwd(const GemvArgs &a, uint32_t lrow, double v, int lane)
{
  gemv_store(a, lrow, v, lane);
}
__device__ __forceinline__ void gemv_store(const GemvArgs &a, uint32_t lrow, double v, int lane)
{
  const uint32_t g = a.row0 + lrow;
  v += a.sdiag * a.alpha[g] * a.xdiag[g];
  if (a.n_peers > 1)
    {
      if (lane < a.n_peers) a.peer_base[lane][a.buf_off + g] = v; // lane q -> rank q (NVLink store)
    }
  else if (lane == 0)
    a.y[lrow] = v;
}

// one row, its column chunks split over the 8 warps of the CTA (used for the remainder rows)
__device__ __forceinline__ void gemv_row_split(const GemvArgs &a, uint32_t idx, int warp, int lane, double *red)
{
  double a1[GEMV_MAXR], a2[GEMV_MAXR];
  a1[0] = a2[0] = 0.0;
  uint32_t rows[GEMV_MAXR];
  rows[0] = a.row_list ? a.row_list[idx] : idx;
  const int b1 = a.n1 * warp / GEMV_WARPS, e1 = a.n1 * (warp + 1) / GEMV_WARPS;
  const int b2 = a.n2 * warp / GEMV_WARPS, e2 = a.n2 * (warp + 1) / GEMV_WARPS;
  gemv_pass<1, 8>(a.M1, a.x1, a.list1 + b1, e1 - b1, a.ld, rows, lane, a1);
  gemv_pass<1, 8>(a.M2, a.x2, a.list2 + b2, e2 - b2, a.ld, rows, lane, a2);
  double v = a.s1 * a1[0] + a.s2 * a2[0];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0)
    {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < GEMV_WARPS; ++w) t += red[w]; // fixed order: deterministic
      gemv_store(a, rows[0], t, lane);
    }
  __syncthreads();
}

// y[r] = s1 * (M1[r,:] . x1) + s2 * (M2[r,:] . x2) + sdiag * alpha[row0+r] * xdiag[row0+r]
// One resident wave of CTAs (grid = SMs x occupancy).  Every warp streams the same number q of
// whole rows; the n_rows - q*warps remainder rows are split column-wise over the 8 warps of a
// CTA, so all warps finish together whatever the row count (no tail wave, no 3-rows-vs-2 skew
// on row-sharded runs).
__global__ void __launch_bounds__(GEMV_WARPS * 32, 3) k_bem_gemv(const GemvArgs a)
{
  __shared__ double red[GEMV_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t G = gridDim.x, c = blockIdx.x;
  const uint32_t Wt = G * GEMV_WARPS;
  const bool idle = a.stop && *a.stop != 0;
  const uint32_t q = idle ? 0u : a.n_rows / Wt, rem = idle ? 0u : a.n_rows - q * Wt;
  uint32_t r0 = (c * GEMV_WARPS + warp) * q;
  const uint32_t r1 = r0 + q;
  while (r0 < r1)
    {
      const uint32_t left = r1 - r0;
      // passes of at most GEMV_MAXR rows, split evenly (e.g. 6 -> 3 + 3, 9 -> 5 + 4)
      const uint32_t passes = (left + GEMV_MAXR - 1) / GEMV_MAXR;
      const uint32_t take = (left + passes - 1) / passes;
      switch (take)
        {
        case 1: gemv_rows<1, 8>(a, r0, lane); break;
        case 2: gemv_rows<2, 4>(a, r0, lane); break;
        case 3: gemv_rows<3, 3>(a, r0, lane); break;
        case 4: gemv_rows<4, 2>(a, r0, lane); break;
        default: gemv_rows<5, 2>(a, r0, lane); break;
        }
      r0 += take;
    }
  for (uint32_t j = c; j < rem; j += G) gemv_row_split(a, q * Wt + j, warp, lane, red);
  if (a.n_peers > 1)
    { // completion: the last CTA of this launch publishes "rank's rows of epoch e are in place"
      __shared__ bool s_last;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          __threadfence_system();
          const unsigned long long prev = atomicAdd(a.done_counter, 1ull);
          s_last = (prev + 1ull == a.done_target);
        }
      __syncthreads();
      if (s_last && threadIdx.x < a.n_peers)
        {
          __threadfence_system();
          unsigned long long *flag =
            reinterpret_cast<unsigned long long *>(a.peer_base[threadIdx.x] + a.flag_off) + a.rank;
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(a.epoch) : "memory");
        }
    }
}

// multipliers: x1p[colpos[i]] = m1[i] src[i], x2p[colpos[i]] = m2[i] src[i], xdiag[i] = m1[i] src[i].
// src_rw != null: the source vector is first scaled in place (the normalisation of the newest Krylov
// vector, folded into the mat-vec that consumes it)
__global__ void k_prep_multipliers(uint32_t N, const double *__restrict__ src,
                                   const double *__restrict__ m1, const double *__restrict__ m2,
                                   const uint32_t *__restrict__ colpos, double *__restrict__ x1p,
                                   double *__restrict__ x2p, double *__restrict__ xdiag, double scale = 1.0,
                                   double *src_rw = nullptr, const double *scale_ptr = nullptr, const int *stop = nullptr)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (stop && *stop != 0) return;
  if (scale_ptr) scale = *scale_ptr;
  double s;
  if (src_rw)
    {
      s = src_rw[i] * scale;
      src_rw[i] = s;
    }
  else
    s = src[i];
  const double a = m1[i] * s, b = m2[i] * s;
  const uint32_t c = colpos[i];
  x1p[c] = a;
  x2p[c] = b;
  xdiag[i] = a;
}

// dst = gathered rows, with constrained rows replaced by src_i - sum c_ik src_k
__global__ void k_epilogue(uint32_t N, const double *y, const double *__restrict__ src,
                           const int32_t *__restrict__ line_of, const uint32_t *__restrict__ cptr,
                           const uint32_t *__restrict__ ccol, const double *__restrict__ cval,
                           double shift, double *__restrict__ dst, const unsigned long long *flags, int n_peers,
                           unsigned long long epoch, unsigned int *timeout_flag, const int *stop = nullptr)
{
  if (flags)
    { // fused gather: wait until every rank has raised its flag for this epoch.  The wait is
      // bounded (20 s): a peer that died must surface as an error, not as a hung device
      if (threadIdx.x < n_peers)
        {
          unsigned long long v, t0 = 0, t1;
          unsigned int spins = 0;
          for (;;)
            {
              asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
              if (v >= epoch) break;
              if ((++spins & 0x3ffu) == 0)
                {
                  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                  if (!t0) t0 = t1;
                  if (t1 - t0 > 20000000000ull)
                    {
                      atomicExch(timeout_flag, 1u);
                      break;
                    }
                }
            }
        }
      __syncthreads();
    }
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (stop && *stop != 0) return;
  double v = y[i] + shift;
  if (line_of)
    {
      const int l = line_of[i];
      if (l >= 0)
        {
          v = src[i];
          for (uint32_t k = cptr[l]; k < cptr[l + 1]; ++k) v -= cval[k] * src[ccol[k]];
        }
    }
  dst[i] = v;
}

// single-CTA l2 norm (used only by the pure-Neumann shift of vmult, :667-668)
__global__ void k_norm2_single(uint32_t N, const double *__restrict__ v, double *__restrict__ out)
{
  __shared__ double red[32];
  double s = 0;
  for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) s = fma(v[i], v[i], s);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32)
    {
      s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (threadIdx.x == 0) out[0] = sqrt(s);
    }
}

__global__ void k_add_neg_scalar(uint32_t N, double *__restrict__ v, const double *__restrict__ s)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) v[i] -= s[0];
}

int wbem_allgather_rows(wbem_ctx *ctx, double *d_buf)
{
  if (ctx->p.world_size <= 1) return 0;
  return wbem_allgather_bytes(ctx, d_buf, sizeof(double) * (size_t)ctx->chunk);
}

int wbem_allgather_bytes(wbem_ctx *ctx, void *d_buf, size_t bytes_per_rank)
{
  if (ctx->p.world_size <= 1) return 0;
  if (ctx->group) return wbem_group_allgather(ctx, d_buf, bytes_per_rank); // one process: peer copies
  if (!ctx->nccl_comm)
    {
      // diagnostics only: time one rank's share of a sharded run on a single GPU
      if (getenv("WBEM_DIAG_NO_COMM")) return 0;
      WBEM_FAIL(ctx, -5, "world_size=%d but wbem_comm_init was not called", ctx->p.world_size);
    }
  const char *base = (const char *)d_buf;
  return wbem_nccl_allgather(ctx, base + bytes_per_rank * ctx->p.rank, d_buf, bytes_per_rank);
}

// after a synchronisation point: did a fused gather give up waiting for a peer?
int wbem_check_gather_timeout(wbem_ctx *ctx)
{
  if (!ctx->d_gather_timeout) return 0;
  unsigned int f = 0;
  CUDA_OK(ctx, cudaMemcpyAsync(&f, ctx->d_gather_timeout, sizeof(f), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (f) WBEM_FAIL(ctx, -5, "fused mat-vec gather: a peer row block never delivered its rows (20 s)");
  return 0;
}

// chunk lists (see header comment): d_list_o = chunks with any other_nodes != 0, d_list_s =
// chunks with any surface_nodes != 0, in storage-column order; built by wbem_set_masks.
int wbem_apply_operator(wbem_ctx *ctx, int mode, const double *d_src, double *d_dst, bool constrained)
{
  return wbem_apply_operator_ex(ctx, mode, d_src, d_dst, constrained, 1.0, nullptr, false);
}

// src_scale_rw != null: scale that vector in place by src_scale first (it is d_src).
// with_precond: d_dst = M^-1 (operator d_src) -- for the sparse approximate inverse the operator's
// epilogue and the preconditioner are ONE kernel (wbem_spai_apply_fused), else they run one after the other.
int wbem_apply_operator_ex(wbem_ctx *ctx, int mode, const double *d_src, double *d_dst, bool constrained, double src_scale,
                           double *src_scale_rw, bool with_precond)
{
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "operator applied before wbem_assemble");
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "operator applied before wbem_set_masks");
  if (!ctx->have_alpha) WBEM_FAIL(ctx, -3, "operator applied before alpha was computed");
  const uint32_t *list_o = ctx->d_list_o, *list_s = ctx->d_list_s;
  const int n_o = ctx->n_list_o, n_s = ctx->n_list_s;
  const double *m1 = mode == 0 ? ctx->d_other : ctx->d_surf; // multiplier mask of N
  const double *m2 = mode == 0 ? ctx->d_surf : ctx->d_other; // multiplier mask of D
  k_prep_multipliers<<<(N + 255) / 256, 256, 0, st>>>(N, d_src, m1, m2, ctx->d_colpos, ctx->d_xn,
                                                     ctx->d_xd, ctx->d_xdiag, src_scale, src_scale_rw,
                                                     src_scale_rw ? ctx->op_scale_ptr : nullptr, ctx->op_stop_ptr);
  ctx->launches++;
  double *yloc = ctx->d_yloc + (size_t)ctx->p.rank * ctx->chunk;
  const int P = ctx->p.world_size;
  const bool use_p2p = P > 1 && ctx->p2p_ready && !((mode == 0) && ctx->pure_neumann);
  const size_t p2p_len = (size_t)ctx->chunk * P, p2p_half = (size_t)WBEM_MULTI_MAX * p2p_len;
  if (use_p2p) ctx->p2p_epoch++;
  const double *ygather = use_p2p ? ctx->d_p2p + (ctx->p2p_epoch & 1ull) * p2p_half : ctx->d_yloc;
  if (ctx->nloc || use_p2p)
    {
      ctx->timer.begin(T_GEMV);
      int &ctas_per_sm = ctx->gemv_ctas_per_sm, &n_sm = ctx->n_sm;
      if (!ctas_per_sm)
        {
          CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_bem_gemv, GEMV_WARPS * 32, 0));
          CUDA_OK(ctx, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->dev));
          if (ctas_per_sm < 1) ctas_per_sm = 1;
        }
      GemvArgs ga;
      ga.stop = ctx->op_stop_ptr;
      ga.M1 = ctx->d_Nm;
      ga.x1 = ctx->d_xn;
      ga.M2 = ctx->d_Dm;
      ga.x2 = ctx->d_xd;
      ga.alpha = ctx->d_alpha;
      ga.xdiag = ctx->d_xdiag;
      ga.ld = ctx->ld;
      ga.nloc = ctx->nloc;
      ga.row0 = ctx->row0;
      ga.y = yloc;
      // rows a constraint line overwrites afterwards need not be streamed at all
      // (the pure-Neumann shift needs the norm over ALL rows of the raw product, :667-668)
      const bool skip = constrained && ctx->n_lines > 0 && ctx->d_free_rows && !(mode == 0 && ctx->pure_neumann);
      ga.row_list = skip ? ctx->d_free_rows : nullptr;
      ga.n_rows = skip ? ctx->n_free_rows : ctx->nloc;
      if (mode == 0)
        {
          ga.list1 = list_o; ga.n1 = n_o; ga.s1 = 1.0;
          ga.list2 = list_s; ga.n2 = n_s; ga.s2 = -1.0;
          ga.sdiag = 1.0;
        }
      else
        {
          ga.list1 = list_s; ga.n1 = n_s; ga.s1 = -1.0;
          ga.list2 = list_o; ga.n2 = n_o; ga.s2 = 1.0;
          ga.sdiag = -1.0;
        }
      uint32_t grid = (uint32_t)(n_sm * ctas_per_sm);
      if (ga.n_rows == 0) grid = 0;
      if (grid > ga.n_rows && ga.n_rows > 0) grid = ga.n_rows; // tiny problems: one split row per CTA
      ga.n_peers = 1;
      ga.rank = ctx->p.rank;
      if (use_p2p)
        {
          if (grid == 0) grid = 1; // a rank without rows still has to raise its flag
          ga.n_peers = P;
          for (int q = 0; q < P; ++q) ga.peer_base[q] = ctx->peer_base[q];
          ga.buf_off = (ctx->p2p_epoch & 1ull) * p2p_half;
          ga.flag_off = 2 * p2p_half;
          ga.epoch = ctx->p2p_epoch;
          ctx->gemv_done_total += grid;
          ga.done_target = ctx->gemv_done_total;
          ga.done_counter = ctx->d_done_counter;
        }
      if (grid)
        {
          k_bem_gemv<<<grid, GEMV_WARPS * 32, 0, st>>>(ga);
          ctx->launches++;
        }
      ctx->tm.gemv_bytes_last = 8.0 * 64.0 * (double)(n_o + n_s) * (double)ga.n_rows;
      ctx->timer.end();
    }
  if (P > 1 && !use_p2p)
    {
      ctx->timer.begin(T_ALLGATHER);
      int rc = wbem_allgather_rows(ctx, ctx->d_yloc);
      ctx->timer.end();
      if (rc) return rc;
    }
  if (use_p2p && ctx->group && ctx->p.fused_gather_on_shared_device)
    { // test mode, row blocks sharing a device: two streams of one device may share a hardware
      // queue, and an epilogue spinning for a mat-vec queued behind it would never end.  So every
      // block launches its mat-vec before any block launches its (waiting) epilogue.
      const int brc = wbem_group_barrier(ctx);
      if (brc) return brc;
    }
  const bool shift = (mode == 0) && ctx->pure_neumann;
  if (shift)
    {
      k_norm2_single<<<1, 1024, 0, st>>>(N, ctx->d_yloc, ctx->d_h + 512);
      ctx->launches++;
      k_add_neg_scalar<<<(N + 255) / 256, 256, 0, st>>>(N, ctx->d_yloc, ctx->d_h + 512);
      ctx->launches++;
    }
  const bool con = constrained && ctx->n_lines > 0;
  const unsigned long long *flags =
    use_p2p ? reinterpret_cast<const unsigned long long *>(ctx->d_p2p + 2 * p2p_half) : nullptr;
  if (with_precond && ctx->p.precond_kind == 1 && ctx->p.preconditioner_band > 0)
    { // epilogue (gather wait, constrained rows) + sparse approximate inverse in one kernel
      EpilogueArgs ea;
      ea.y = ygather;
      ea.src = d_src;
      ea.line_of = con ? ctx->d_con_line_of : nullptr;
      ea.cptr = ctx->d_con_ptr;
      ea.ccol = ctx->d_con_col;
      ea.cval = ctx->d_con_val;
      ea.flags = flags;
      ea.n_peers = P;
      ea.epoch = ctx->p2p_epoch;
      ea.timeout_flag = ctx->d_gather_timeout;
      ea.stop = ctx->op_stop_ptr;
      return wbem_spai_apply_fused(ctx, ea, d_dst);
    }
  double *d_mid = with_precond ? ctx->d_tmp[0] : d_dst;
  k_epilogue<<<(N + 255) / 256, 256, 0, st>>>(N, ygather, d_src, con ? ctx->d_con_line_of : nullptr, ctx->d_con_ptr,
                                             ctx->d_con_col, ctx->d_con_val, 0.0, d_mid, flags, P, ctx->p2p_epoch,
                                             ctx->d_gather_timeout, ctx->op_stop_ptr);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  if (with_precond) return wbem_apply_preconditioner(ctx, d_mid, d_dst);
  return 0;
}

// ---------------------------------------------------------------------------------------
// Block mat-vec: the same operator applied to nb <= WBEM_MULTI_MAX vectors while the matrices are
// streamed ONCE (the J.v pattern of FreeSurface::jacobian, reference source/free_surface.cc:4918-4993:
// one inner GMRES per outer Krylov vector, unchanged matrices).  One CTA per group of MV_NR rows;
// its 8 warps split the column chunks, every lane keeps MV_NR x NB accumulators; fixed-order
// reduction through shared memory.  Bytes per launch = those of ONE single-vector application.
// ---------------------------------------------------------------------------------------
#define MV_MAXR 4 // rows one warp streams in one pass of the block mat-vec
struct GemvMultiArgs
{
  GemvArgs g;            // x1 / x2 / xdiag / y hold the FIRST vector; the others follow at the strides below.
                         // The signs s1 / s2 / sdiag are already folded into the multiplier vectors.
  int nb;
  size_t xstride;        // doubles between consecutive multiplier vectors
  size_t ystride;        // doubles between consecutive result vectors (local results or gather slabs)
};

// acc[r][k] += sum over the listed chunks [b, e) of M[rows[r], :] . x_k ; NR rows share every x chunk,
// NC chunks in flight (NR * NC independent 128-bit matrix loads per lane)
template <int NR, int NC, int NB>
__device__ __forceinline__ void gemv_multi_pass(const double *__restrict__ M, const double *__restrict__ x, size_t xstride,
                                                const uint32_t *__restrict__ list, int b, int e, uint32_t ld,
                                                const uint32_t rows[MV_MAXR], int lane, double acc[MV_MAXR][NB])
{
  const double *base[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) base[r] = M + (size_t)rows[r] * ld + lane * 2;
  int i = b;
  for (; i + NC <= e; i += NC)
    {
      uint32_t c[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) c[j] = list[i + j] * 64;
      double2 m[NC][NR];
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int r = 0; r < NR; ++r) m[j][r] = ld_stream(base[r] + c[j]);
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int k = 0; k < NB; ++k)
          {
            const double2 xv = *reinterpret_cast<const double2 *>(x + k * xstride + c[j] + lane * 2);
#pragma unroll
            for (int r = 0; r < NR; ++r)
              {
                acc[r][k] = fma(m[j][r].x, xv.x, acc[r][k]);
                acc[r][k] = fma(m[j][r].y, xv.y, acc[r][k]);
              }
          }
    }
  for (; i < e; ++i)
    {
      const uint32_t c0 = list[i] * 64;
      double2 m0[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) m0[r] = ld_stream(base[r] + c0);
#pragma unroll
      for (int k = 0; k < NB; ++k)
        {
          const double2 xv = *reinterpret_cast<const double2 *>(x + k * xstride + c0 + lane * 2);
#pragma unroll
          for (int r = 0; r < NR; ++r)
            {
              acc[r][k] = fma(m0[r].x, xv.x, acc[r][k]);
              acc[r][k] = fma(m0[r].y, xv.y, acc[r][k]);
            }
        }
    }
}

template <int NB>
__device__ __forceinline__ void gemv_multi_store(const GemvMultiArgs &ma, uint32_t lrow, int k, double v)
{
  const GemvArgs &a = ma.g;
  const uint32_t g = a.row0 + lrow;
  v += a.alpha[g] * a.xdiag[k * ma.xstride + g];
  if (a.n_peers > 1)
    for (int q = 0; q < a.n_peers; ++q) a.peer_base[q][a.buf_off + k * ma.ystride + g] = v; // NVLink stores
  else
    a.y[k * ma.ystride + lrow] = v;
}

template <int NR, int NC, int NB>
__device__ __forceinline__ void gemv_multi_rows(const GemvMultiArgs &ma, uint32_t r0, int lane)
{
  const GemvArgs &a = ma.g;
  double acc[MV_MAXR][NB];
#pragma unroll
  for (int r = 0; r < MV_MAXR; ++r)
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[r][k] = 0.0;
  uint32_t rows[MV_MAXR];
#pragma unroll
  for (int r = 0; r < NR; ++r) rows[r] = a.row_list ? a.row_list[r0 + r] : r0 + r;
  gemv_multi_pass<NR, NC, NB>(a.M1, a.x1, ma.xstride, a.list1, 0, a.n1, a.ld, rows, lane, acc);
  gemv_multi_pass<NR, NC, NB>(a.M2, a.x2, ma.xstride, a.list2, 0, a.n2, a.ld, rows, lane, acc);
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int k = 0; k < NB; ++k)
      {
        double v = acc[r][k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == (k & 31) && k < ma.nb) gemv_multi_store<NB>(ma, rows[r], k, v);
      }
}

// Same work split as k_bem_gemv: one resident wave, every warp streams the same number q of whole
// rows (passes of <= MV_MAXR rows sharing each x chunk), the remainder rows are split column-wise
// over the 8 warps of a CTA.
template <int NB>
__global__ void __launch_bounds__(GEMV_WARPS * 32, 2) k_bem_gemv_multi(const GemvMultiArgs ma)
{
  const GemvArgs &a = ma.g;
  __shared__ double red[GEMV_WARPS][NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t G = gridDim.x, c = blockIdx.x;
  const uint32_t Wt = G * GEMV_WARPS;
  const bool idle = a.stop && *a.stop != 0;
  const uint32_t q = idle ? 0u : a.n_rows / Wt, rem = idle ? 0u : a.n_rows - q * Wt;
  uint32_t r0 = (c * GEMV_WARPS + warp) * q;
  const uint32_t r1 = r0 + q;
  while (r0 < r1)
    {
      const uint32_t left = r1 - r0;
      const uint32_t passes = (left + MV_MAXR - 1) / MV_MAXR;
      const uint32_t take = (left + passes - 1) / passes;
      switch (take)
        {
        case 1: gemv_multi_rows<1, 4, NB>(ma, r0, lane); break;
        case 2: gemv_multi_rows<2, 2, NB>(ma, r0, lane); break;
        case 3: gemv_multi_rows<3, 2, NB>(ma, r0, lane); break;
        default: gemv_multi_rows<4, 2, NB>(ma, r0, lane); break;
        }
      r0 += take;
    }
  for (uint32_t j = c; j < rem; j += G)
    { // one row, its column chunks split over the 8 warps
      uint32_t rows[MV_MAXR];
      const uint32_t idx = q * Wt + j;
      rows[0] = a.row_list ? a.row_list[idx] : idx;
      double acc[MV_MAXR][NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) acc[0][k] = 0.0;
      const int b1 = a.n1 * warp / GEMV_WARPS, e1 = a.n1 * (warp + 1) / GEMV_WARPS;
      const int b2 = a.n2 * warp / GEMV_WARPS, e2 = a.n2 * (warp + 1) / GEMV_WARPS;
      gemv_multi_pass<1, 4, NB>(a.M1, a.x1, ma.xstride, a.list1, b1, e1, a.ld, rows, lane, acc);
      gemv_multi_pass<1, 4, NB>(a.M2, a.x2, ma.xstride, a.list2, b2, e2, a.ld, rows, lane, acc);
#pragma unroll
      for (int k = 0; k < NB; ++k)
        {
          double v = acc[0][k];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if (lane == 0) red[warp][k] = v;
        }
      __syncthreads();
      if ((int)threadIdx.x < ma.nb)
        {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < GEMV_WARPS; ++w) t += red[w][threadIdx.x]; // fixed order: deterministic
          gemv_multi_store<NB>(ma, rows[0], threadIdx.x, t);
        }
      __syncthreads();
    }
  if (a.n_peers > 1)
    { // completion: the last CTA of this launch publishes "rank's rows of epoch e are in place"
      __shared__ bool s_last;
      __syncthreads();
      if (threadIdx.x == 0)
        {
          __threadfence_system();
          const unsigned long long prev = atomicAdd(a.done_counter, 1ull);
          s_last = (prev + 1ull == a.done_target);
        }
      __syncthreads();
      if (s_last && threadIdx.x < a.n_peers)
        {
          __threadfence_system();
          unsigned long long *flag =
            reinterpret_cast<unsigned long long *>(a.peer_base[threadIdx.x] + a.flag_off) + a.rank;
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(a.epoch) : "memory");
        }
    }
}

// multipliers of the block mat-vec, signs folded in: x1p = s1 m1 src, x2p = s2 m2 src, xdiag = sdiag m1 src
__global__ void k_prep_multipliers_signed(uint32_t N, const double *__restrict__ src, const double *__restrict__ m1,
                                          const double *__restrict__ m2, const uint32_t *__restrict__ colpos,
                                          double s1, double s2, double sdiag, double *__restrict__ x1p,
                                          double *__restrict__ x2p, double *__restrict__ xdiag)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double s = src[i];
  const double a = m1[i] * s, b = m2[i] * s;
  const uint32_t c = colpos[i];
  x1p[c] = s1 * a;
  x2p[c] = s2 * b;
  xdiag[i] = sdiag * a;
}

// d_dst[b] = operator(d_src[b]) for b < nb, one pass over the matrices.  The multiplier vectors are
// kept in ctx->multi (allocated by gmres.cu): xn / xd / xdiag slabs of ld doubles each.

int wbem_apply_operator_multi(wbem_ctx *ctx, int mode, int nb, const double *const *d_src, double *const *d_dst,
                              bool constrained)
{
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  if (nb < 1 || nb > WBEM_MULTI_MAX) WBEM_FAIL(ctx, -1, "block mat-vec takes 1..%d vectors", WBEM_MULTI_MAX);
  if (!ctx->assembled || !ctx->have_masks || !ctx->have_alpha) WBEM_FAIL(ctx, -3, "operator applied before assembly / masks / alpha");
  MultiWork *mw = wbem_multi_work(ctx);
  if (!mw) return -2;
  const double *m1 = mode == 0 ? ctx->d_other : ctx->d_surf; // multiplier mask of N
  const double *m2 = mode == 0 ? ctx->d_surf : ctx->d_other; // multiplier mask of D
  const size_t ld = ctx->ld;
  const double sg1 = mode == 0 ? 1.0 : -1.0, sg2 = -sg1, sgd = sg1; // vmult: +N -D +alpha ; rhs: -N +D -alpha
  for (int b = 0; b < nb; ++b)
    {
      k_prep_multipliers_signed<<<(N + 255) / 256, 256, 0, st>>>(N, d_src[b], m1, m2, ctx->d_colpos, sg1, sg2, sgd,
                                                                mw->d_xn + b * ld, mw->d_xd + b * ld, mw->d_xdiag + b * ld);
      ctx->launches++;
    }
  const int P = ctx->p.world_size;
  const bool shift = (mode == 0) && ctx->pure_neumann;
  const bool use_p2p = P > 1 && ctx->p2p_ready && !shift;
  const size_t p2p_len = (size_t)ctx->chunk * P, p2p_half = (size_t)WBEM_MULTI_MAX * p2p_len;
  if (use_p2p) ctx->p2p_epoch++;
  if (!ctx->d_ymulti)
    {
      CUDA_OK(ctx, cudaMalloc((void **)&ctx->d_ymulti, sizeof(double) * WBEM_MULTI_MAX * (p2p_len + 64)));
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_ymulti, 0, sizeof(double) * WBEM_MULTI_MAX * (p2p_len + 64), st));
    }
  const size_t ystride = use_p2p ? p2p_len : p2p_len + 64;
  const double *ygather = use_p2p ? ctx->d_p2p + (ctx->p2p_epoch & 1ull) * p2p_half : ctx->d_ymulti;
  {
    ctx->timer.begin(T_GEMV);
    const int variant = nb <= 2 ? 0 : (nb <= 4 ? 1 : 2);
    int &ctas_per_sm = ctx->gemv_multi_ctas_per_sm[variant], &n_sm = ctx->n_sm;
    if (!ctas_per_sm)
      {
        if (variant == 0)
          CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_bem_gemv_multi<2>, GEMV_WARPS * 32, 0));
        else if (variant == 1)
          CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_bem_gemv_multi<4>, GEMV_WARPS * 32, 0));
        else
          CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_bem_gemv_multi<WBEM_MULTI_MAX>,
                                                                     GEMV_WARPS * 32, 0));
        CUDA_OK(ctx, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->dev));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
      }
    GemvMultiArgs ma;
    GemvArgs &ga = ma.g;
    ga.stop = nullptr;
    ga.M1 = ctx->d_Nm;
    ga.x1 = mw->d_xn;
    ga.M2 = ctx->d_Dm;
    ga.x2 = mw->d_xd;
    ga.alpha = ctx->d_alpha;
    ga.xdiag = mw->d_xdiag;
    ga.ld = ctx->ld;
    ga.nloc = ctx->nloc;
    ga.row0 = ctx->row0;
    ga.y = ctx->d_ymulti + (size_t)ctx->p.rank * ctx->chunk;
    const bool skip = constrained && ctx->n_lines > 0 && ctx->d_free_rows && !shift;
    ga.row_list = skip ? ctx->d_free_rows : nullptr;
    ga.n_rows = skip ? ctx->n_free_rows : ctx->nloc;
    if (mode == 0)
      {
        ga.list1 = ctx->d_list_o; ga.n1 = ctx->n_list_o; ga.s1 = 1.0;
        ga.list2 = ctx->d_list_s; ga.n2 = ctx->n_list_s; ga.s2 = -1.0;
        ga.sdiag = 1.0;
      }
    else
      {
        ga.list1 = ctx->d_list_s; ga.n1 = ctx->n_list_s; ga.s1 = -1.0;
        ga.list2 = ctx->d_list_o; ga.n2 = ctx->n_list_o; ga.s2 = 1.0;
        ga.sdiag = -1.0;
      }
    ma.nb = nb;
    ma.xstride = ld;
    ma.ystride = ystride;
    uint32_t grid = (uint32_t)(n_sm * ctas_per_sm);
    if (ga.n_rows == 0) grid = 0;
    if (grid > ga.n_rows && ga.n_rows > 0) grid = ga.n_rows;
    ga.n_peers = 1;
    ga.rank = ctx->p.rank;
    if (use_p2p)
      {
        if (grid == 0) grid = 1;
        ga.n_peers = P;
        for (int q = 0; q < P; ++q) ga.peer_base[q] = ctx->peer_base[q];
        ga.buf_off = (ctx->p2p_epoch & 1ull) * p2p_half;
        ga.flag_off = 2 * p2p_half;
        ga.epoch = ctx->p2p_epoch;
        ctx->gemv_done_total += grid;
        ga.done_target = ctx->gemv_done_total;
        ga.done_counter = ctx->d_done_counter;
      }
    if (grid)
      {
        if (nb <= 2)
          k_bem_gemv_multi<2><<<grid, GEMV_WARPS * 32, 0, st>>>(ma);
        else if (nb <= 4)
          k_bem_gemv_multi<4><<<grid, GEMV_WARPS * 32, 0, st>>>(ma);
        else
          k_bem_gemv_multi<WBEM_MULTI_MAX><<<grid, GEMV_WARPS * 32, 0, st>>>(ma);
        ctx->launches++;
      }
    ctx->tm.gemv_bytes_last = 8.0 * 64.0 * (double)(ctx->n_list_o + ctx->n_list_s) * (double)ga.n_rows;
    ctx->timer.end();
  }
  if (use_p2p && ctx->group && ctx->p.fused_gather_on_shared_device)
    {
      const int brc = wbem_group_barrier(ctx);
      if (brc) return brc;
    }
  if (P > 1 && !use_p2p)
    {
      ctx->timer.begin(T_ALLGATHER);
      for (int b = 0; b < nb; ++b)
        {
          const int rc = wbem_allgather_rows(ctx, ctx->d_ymulti + b * ystride);
          if (rc) return rc;
        }
      ctx->timer.end();
    }
  const bool con = constrained && ctx->n_lines > 0;
  for (int b = 0; b < nb; ++b)
    {
      double *yb = const_cast<double *>(ygather) + b * ystride;
      if (shift)
        {
          k_norm2_single<<<1, 1024, 0, st>>>(N, yb, ctx->d_h + 512);
          k_add_neg_scalar<<<(N + 255) / 256, 256, 0, st>>>(N, yb, ctx->d_h + 512);
          ctx->launches += 2;
        }
      k_epilogue<<<(N + 255) / 256, 256, 0, st>>>(
        N, yb, d_src[b], con ? ctx->d_con_line_of : nullptr, ctx->d_con_ptr, ctx->d_con_col, ctx->d_con_val, 0.0, d_dst[b],
        use_p2p ? reinterpret_cast<const unsigned long long *>(ctx->d_p2p + 2 * p2p_half) : nullptr, P, ctx->p2p_epoch,
        ctx->d_gather_timeout);
      ctx->launches++;
    }
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}
