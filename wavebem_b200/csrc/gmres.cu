// gmres.cu -- BEMProblem<3>::solve_system (reference source/bem_problem.cc:821-895):
//   alpha, compute_rhs, distribute_rhs, band preconditioner, restarted left-preconditioned
//   GMRES (deal.II SolverGMRES with AdditionalData(100): 98 Krylov vectors per cycle,
//   absolute tolerance on the preconditioned residual estimate), unpack into phi / dphi_dn.
//
// Krylov vectors live in device memory (replicated on every rank, so no all-reduce is ever
// needed); the Hessenberg column, the Givens rotations and the stopping test are host
// scalars, fetched with one small D2H copy per iteration.  Orthogonalisation is classical
// Gram-Schmidt applied twice (CGS2) -- two batched kernels per pass instead of the 2k vector
// kernels of deal.II's modified Gram-Schmidt loop; same Arnoldi relation to rounding.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "internal.h"

// ---------------------------------------------------------------------------------------
// vector kernels
// ---------------------------------------------------------------------------------------
#define DOT_SEGS 8
#define GM_KMAX WBEM_GMRES_KMAX // largest Krylov basis (row stride of the small device work arrays)
// partial[seg*GM_KMAX + i] = V[i] . w over segment seg of the vector  (grid = k x DOT_SEGS;
// fixed reduction order inside a CTA, the segments are summed in order by the consumer)
__global__ void __launch_bounds__(256)
  k_dots(uint32_t N, const double *__restrict__ V, size_t ldv, const double *__restrict__ w,
         double *__restrict__ partial, const int *stop = nullptr)
{
  __shared__ double red[8];
  if (stop && *stop != 0) return; // iteration enqueued ahead of a finished solve
  const double *v = V + (size_t)blockIdx.x * ldv;
  const uint32_t seg = blockIdx.y;
  const uint32_t len = (N + DOT_SEGS - 1) / DOT_SEGS;
  const uint32_t b = seg * len, e = min(N, b + len);
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  uint32_t i = b + threadIdx.x;
  for (; i + 768 < e; i += 1024)
    {
      const double v0 = v[i], v1 = v[i + 256], v2 = v[i + 512], v3 = v[i + 768];
      const double w0 = w[i], w1 = w[i + 256], w2 = w[i + 512], w3 = w[i + 768];
      s0 = fma(v0, w0, s0);
      s1 = fma(v1, w1, s1);
      s2 = fma(v2, w2, s2);
      s3 = fma(v3, w3, s3);
    }
  for (; i < e; i += 256) s0 = fma(v[i], w[i], s0);
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
    {
      double t = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k];
      partial[seg * GM_KMAX + blockIdx.x] = t;
    }
}

// One Arnoldi step's small algebra on the device (one warp; lane 0 does the arithmetic): the new
// Hessenberg column from the accumulated projections hacc[0..dim) and the per-CTA sums of squares of the
// orthogonalised vector, the Givens rotations (deal.II SolverGMRES::givens_rotation, as restated for the
// host below -- same operations in the same order, no contraction, so both give the same bits), the
// residual estimate and the stop decision.  With it the host can enqueue several iterations before it looks.
//   gm: [0] 1/h[dim] for the mat-vec that consumes the new vector, [1] rho; then gamma[ntmp+2],
//   ci[ntmp+2], si[ntmp+2], Hs[..][ntmp] (rotated columns)
struct HessArgs
{
  double *gm;            // null: no Hessenberg step behind this launch
  int *ctl;              // [0] stop state, [1] inner iterations done in this cycle
  unsigned int *counter; // CTAs of the launch that have finished (the last one runs the step)
  const double *hacc;
  int inner, ntmp, accumulated, max_steps;
  double tol;
};

struct HessShared
{
  double part[128], h[GM_KMAX + 8], c[GM_KMAX + 8], s[GM_KMAX + 8];
};

__device__ __forceinline__ void hessenberg_step(const HessArgs &a, const double *nrm2, unsigned nparts, HessShared &sh, int lane)
{
  const int inner = a.inner, ntmp = a.ntmp, dim = inner + 1;
  double *gm = a.gm;
  double *gamma = gm + 2, *ci = gamma + (ntmp + 2), *si = ci + (ntmp + 2), *H = si + (ntmp + 2) + (size_t)inner * ntmp;
  for (int i = lane; i < dim; i += 32)
    {
      sh.h[i] = __ldcg(a.hacc + i); // (written by other CTAs of this launch: L2, not a stale L1 line)
      sh.c[i] = __ldcg(ci + i);
      sh.s[i] = __ldcg(si + i);
    }
  const double g_in = __ldcg(gamma + inner);
  double ss = 0;
  for (unsigned base = 0; base < nparts; base += 128)
    {
      __syncwarp();
      for (unsigned i = lane; i < 128 && base + i < nparts; i += 32) sh.part[i] = __ldcg(nrm2 + base + i);
      __syncwarp();
      if (lane == 0)
        for (unsigned i = 0; i < 128 && base + i < nparts; ++i) ss = __dadd_rn(ss, sh.part[i]);
    }
  __syncwarp();
  if (lane != 0) return;
  const double hd = sqrt(ss);
  gm[0] = 1.0 / hd;
  double lo = sh.h[0];
  for (int i = 0; i < inner; ++i)
    { // rotations of the previous columns
      const double s = sh.s[i], c = sh.c[i], t = lo, nx = sh.h[i + 1];
      H[i] = __dadd_rn(__dmul_rn(c, t), __dmul_rn(s, nx));
      lo = __dadd_rn(__dmul_rn(-s, t), __dmul_rn(c, nx));
    }
  const double r = 1.0 / sqrt(__dadd_rn(__dmul_rn(lo, lo), __dmul_rn(hd, hd))); // the new rotation
  const double sn = __dmul_rn(hd, r), cs = __dmul_rn(lo, r);
  si[inner] = sn;
  ci[inner] = cs;
  H[inner] = __dadd_rn(__dmul_rn(cs, lo), __dmul_rn(sn, hd));
  gamma[dim] = __dmul_rn(-sn, g_in);
  gamma[inner] = __dmul_rn(g_in, cs);
  const double rho = fabs(__dmul_rn(-sn, g_in));
  gm[1] = rho;
  int state = (rho <= a.tol) ? 1 : ((a.accumulated >= a.max_steps) ? 2 : 0);
  if (!(rho == rho)) state = 2;
  a.ctl[1] = dim;
  __threadfence();
  a.ctl[0] = state;
}

// h[i] = sum_seg partial ; w -= sum_i h[i] V[i] ; hacc[i] (+)= h[i] (block 0) ;
// nrm2[blockIdx.x] = sum over this CTA's entries of w_j^2 (after the update)
__global__ void __launch_bounds__(128)
  k_project_out(uint32_t N, int k, const double *__restrict__ V, size_t ldv,
                const double *__restrict__ partial, double *__restrict__ w, double *__restrict__ hacc,
                int accumulate, double *__restrict__ nrm2, const int *stop = nullptr, const HessArgs hs = HessArgs{})
{
  extern __shared__ double sh[];
  __shared__ double red[4];
  __shared__ bool s_last;
  __shared__ HessShared s_hess;
  if (stop && *stop != 0) return;
  for (int i = threadIdx.x; i < k; i += blockDim.x)
    {
      double t = 0;
#pragma unroll
      for (int sg = 0; sg < DOT_SEGS; ++sg) t += partial[sg * GM_KMAX + i];
      sh[i] = t;
    }
  __syncthreads();
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (j < N)
    {
      // four independent partial sums: the k loads per element are issued in batches of 8
      double p0 = 0, p1 = 0, p2 = 0, p3 = 0;
      const double *vj = V + j;
      int i = 0;
      for (; i + 8 <= k; i += 8)
        {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = vj[(size_t)(i + u) * ldv];
          p0 = fma(sh[i + 0], v[0], p0);
          p1 = fma(sh[i + 1], v[1], p1);
          p2 = fma(sh[i + 2], v[2], p2);
          p3 = fma(sh[i + 3], v[3], p3);
          p0 = fma(sh[i + 4], v[4], p0);
          p1 = fma(sh[i + 5], v[5], p1);
          p2 = fma(sh[i + 6], v[6], p2);
          p3 = fma(sh[i + 7], v[7], p3);
        }
      for (; i < k; ++i) p0 = fma(sh[i], vj[(size_t)i * ldv], p0);
      s = w[j] - ((p0 + p1) + (p2 + p3));
      w[j] = s;
    }
  if (blockIdx.x == 0 && hacc)
    for (int i = threadIdx.x; i < k; i += blockDim.x) hacc[i] = accumulate ? hacc[i] + sh[i] : sh[i];
  double q = s * s;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0)
    {
      double t = 0;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) t += red[kk];
      nrm2[blockIdx.x] = t;
    }
  if (hs.gm)
    { // second pass: the CTA that finishes last does the Arnoldi step's small algebra (no extra launch)
      __syncthreads();
      if (threadIdx.x == 0)
        {
          __threadfence();
          s_last = (atomicAdd(hs.counter, 1u) == gridDim.x - 1);
        }
      __syncthreads();
      if (s_last && threadIdx.x < 32)
        {
          __threadfence();
          if (threadIdx.x == 0) *hs.counter = 0;
          hessenberg_step(hs, nrm2, gridDim.x, s_hess, threadIdx.x);
        }
    }
}

// out[0] = ||v||_2   (single CTA, fixed order)
__global__ void __launch_bounds__(1024) k_norm2(uint32_t N, const double *__restrict__ v, double *out, double *out_inv = nullptr)
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
      s = red[threadIdx.x];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (threadIdx.x == 0)
        {
          out[0] = sqrt(s);
          if (out_inv) out_inv[0] = 1.0 / sqrt(s);
        }
    }
}

__global__ void k_scale(uint32_t N, double *__restrict__ v, double a)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) v[i] *= a;
}
// r = b - p
__global__ void k_residual(uint32_t N, const double *__restrict__ b, const double *p,
                           double *r)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) r[i] = b[i] - p[i];
}
// x += sum_i y[i] V[i]
__global__ void __launch_bounds__(256)
  k_update_solution(uint32_t N, int k, const double *__restrict__ V, size_t ldv,
                    const double *__restrict__ y, double *__restrict__ x)
{
  extern __shared__ double sh[];
  for (int i = threadIdx.x; i < k; i += blockDim.x) sh[i] = y[i];
  __syncthreads();
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < N)
    {
      double s = x[j];
      for (int i = 0; i < k; ++i) s = fma(sh[i], V[(size_t)i * ldv + j], s);
      x[j] = s;
    }
}
__global__ void k_distribute_rhs(uint32_t n_lines, const uint32_t *__restrict__ lines,
                                 const double *__restrict__ inhom, double *__restrict__ rhs)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_lines) rhs[lines[k]] = inhom[k];
}
__global__ void k_unpack(uint32_t N, const double *__restrict__ surf, const double *__restrict__ sol,
                         double *__restrict__ phi, double *__restrict__ dphi)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (surf[i] == 0)
    phi[i] = sol[i];
  else
    dphi[i] = sol[i];
}

// ---------------------------------------------------------------------------------------
// band preconditioner (reference source/bem_problem.cc:1107-1149)
// band row r holds columns i = r - hb + 1 .. r + hb  (r in [i - hb, i + hb))
// ---------------------------------------------------------------------------------------
__global__ void k_extract_band(uint32_t N, uint32_t nloc, uint32_t row0, uint32_t ld, int band,
                               const double *__restrict__ Nm, const double *__restrict__ Dm,
                               const double *__restrict__ alpha, const double *__restrict__ surf,
                               const int32_t *__restrict__ line_of,
                               const uint32_t *__restrict__ colpos, double *__restrict__ out)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r = t / band;
  const int k = t - r * band;
  if (r >= nloc) return;
  const int hb = band / 2;
  const long g = (long)row0 + r;
  const long i = g - hb + 1 + k;
  double v = 0.0;
  if (i >= 0 && i < (long)N)
    {
      const bool row_con = line_of && line_of[g] >= 0;
      if (row_con)
        v = (i == g) ? 1.0 : 0.0;
      else if (surf[i] == 0)
        {
          v = Nm[(size_t)r * ld + colpos[i]];
          if (i == g) v += alpha[i];
        }
      else
        v = -Dm[(size_t)r * ld + colpos[i]];
    }
  out[(size_t)r * band + k] = v;
}

// host banded LU with partial pivoting (row interchanges inside the band), LAPACK gbtrf
// layout: column-major band with kl extra rows of fill.
namespace
{
struct HostBand
{
  int n, kl, ku, ld;
  double *ab;
  int *piv;
  inline double &at(int i, int j) { return ab[(size_t)j * ld + (kl + ku + i - j)]; }
};

int host_band_factor(HostBand &B)
{
  const int n = B.n;
  int reach = 0; // last column touched by row interchanges so far
  for (int j = 0; j < n; ++j)
    {
      const int below = std::min(B.kl, n - 1 - j);
      int p = 0;
      double best = std::fabs(B.at(j, j));
      for (int k = 1; k <= below; ++k)
        {
          const double a = std::fabs(B.at(j + k, j));
          if (a > best)
            {
              best = a;
              p = k;
            }
        }
      B.piv[j] = j + p;
      if (best == 0.0) return j + 1;
      reach = std::max(reach, std::min(j + B.ku + p, n - 1));
      if (p)
        for (int c = j; c <= reach; ++c) std::swap(B.at(j, c), B.at(j + p, c));
      const double inv = 1.0 / B.at(j, j);
      for (int k = 1; k <= below; ++k) B.at(j + k, j) *= inv;
      for (int c = j + 1; c <= reach; ++c)
        {
          const double u = B.at(j, c);
          if (u == 0.0) continue;
          for (int k = 1; k <= below; ++k) B.at(j + k, c) -= B.at(j + k, j) * u;
        }
    }
  return 0;
}

void host_band_solve(HostBand &B, double *x)
{
  const int n = B.n, span = B.kl + B.ku;
  for (int j = 0; j < n; ++j)
    {
      const int below = std::min(B.kl, n - 1 - j);
      if (B.piv[j] != j) std::swap(x[j], x[B.piv[j]]);
      const double xj = x[j];
      for (int k = 1; k <= below; ++k) x[j + k] -= B.at(j + k, j) * xj;
    }
  for (int j = n - 1; j >= 0; --j)
    {
      x[j] /= B.at(j, j);
      const double xj = x[j];
      for (int i = std::max(0, j - span); i < j; ++i) x[i] -= B.at(i, j) * xj;
    }
}
} // namespace

int wbem_build_preconditioner(wbem_ctx *ctx)
{
  const int band = ctx->p.preconditioner_band;
  if (band <= 0)
    {
      ctx->precond_ready = true;
      return 0;
    }
  if (ctx->precond_ready && ctx->precond_version == ctx->op_version) return 0;
  if (ctx->p.precond_kind == 1)
    { // local-inverse sparse approximate inverse (spai.cu) instead of the band LU
      const int rc = wbem_spai_setup(ctx);
      if (rc) return rc;
      ctx->precond_ready = true;
      ctx->precond_version = ctx->op_version;
      return 0;
    }
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  double *loc = ctx->d_band + (size_t)ctx->p.rank * ctx->chunk * band;
  if (ctx->nloc)
    {
      const size_t total = (size_t)ctx->nloc * band;
      k_extract_band<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        N, ctx->nloc, ctx->row0, ctx->ld, band, ctx->d_Nm, ctx->d_Dm, ctx->d_alpha, ctx->d_surf,
        ctx->n_lines ? ctx->d_con_line_of : nullptr, ctx->d_colpos, loc);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  int rc = wbem_allgather_bytes(ctx, ctx->d_band, sizeof(double) * (size_t)ctx->chunk * band);
  if (rc) return rc;

  // the device variant (precond.cu) pivots inside 64 x 64 diagonal blocks only; the reference's sparse
  // LU pivots across the whole band.  Should a diagonal / Schur block come out singular there, this
  // factorisation falls back to the pivoted band LU on the host instead of failing the solve.
  ctx->precond_host_active = ctx->p.precond_on_host != 0;
  if (!ctx->precond_host_active)
    {
      rc = wbem_device_precond_factor(ctx);
      if (rc == -6)
        ctx->precond_host_active = true;
      else if (rc)
        return rc;
    }
  if (ctx->precond_host_active)
    {
      std::vector<double> hb((size_t)N * band);
      CUDA_OK(ctx, cudaMemcpyAsync(hb.data(), ctx->d_band, sizeof(double) * (size_t)N * band,
                                   cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      const int half = band / 2;
      HostBand B;
      B.n = (int)N;
      B.kl = half;      // rows below the diagonal: r - i <= half - 1
      B.ku = half;      // rows above: i - r <= half
      B.ld = 2 * B.kl + B.ku + 1;
      ctx->h_lu.assign((size_t)N * B.ld, 0.0);
      ctx->h_piv.assign(N, 0);
      B.ab = ctx->h_lu.data();
      B.piv = ctx->h_piv.data();
      for (uint32_t r = 0; r < N; ++r)
        for (int k = 0; k < band; ++k)
          {
            const long i = (long)r - half + 1 + k;
            if (i >= 0 && i < (long)N) B.at((int)r, (int)i) = hb[(size_t)r * band + k];
          }
      const int info = host_band_factor(B);
      if (info) WBEM_FAIL(ctx, -6, "band preconditioner is singular at column %d", info - 1);
      ctx->h_kl = B.kl;
      ctx->h_ku = B.ku;
      ctx->h_ldab = B.ld;
    }
  ctx->precond_ready = true;
  ctx->precond_version = ctx->op_version;
  return 0;
}

int wbem_apply_preconditioner(wbem_ctx *ctx, const double *d_in, double *d_out)
{
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  if (ctx->p.preconditioner_band <= 0)
    {
      if (d_in != d_out)
        CUDA_OK(ctx, cudaMemcpyAsync(d_out, d_in, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
  if (!ctx->precond_ready) WBEM_FAIL(ctx, -3, "preconditioner applied before it was assembled");
  if (ctx->p.precond_kind == 1) return wbem_spai_apply(ctx, d_in, d_out);
  if (ctx->precond_host_active)
    {
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_pinned, d_in, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      HostBand B;
      B.n = (int)N;
      B.kl = ctx->h_kl;
      B.ku = ctx->h_ku;
      B.ld = ctx->h_ldab;
      B.ab = ctx->h_lu.data();
      B.piv = ctx->h_piv.data();
      host_band_solve(B, ctx->h_pinned);
      CUDA_OK(ctx, cudaMemcpyAsync(d_out, ctx->h_pinned, sizeof(double) * N, cudaMemcpyHostToDevice, st));
      return 0;
    }
  return wbem_device_precond_solve(ctx, d_in, d_out);
}

// ---------------------------------------------------------------------------------------
// GMRES
// ---------------------------------------------------------------------------------------
namespace
{
inline void givens(std::vector<double> &h, std::vector<double> &b, std::vector<double> &ci,
                   std::vector<double> &si, int col)
{
  for (int i = 0; i < col; ++i)
    {
      const double s = si[i], c = ci[i], t = h[i];
      h[i] = c * t + s * h[i + 1];
      h[i + 1] = -s * t + c * h[i + 1];
    }
  const double r = 1.0 / std::sqrt(h[col] * h[col] + h[col + 1] * h[col + 1]);
  si[col] = h[col + 1] * r;
  ci[col] = h[col] * r;
  h[col] = ci[col] * h[col] + si[col] * h[col + 1];
  b[col + 1] = -si[col] * b[col];
  b[col] *= ci[col];
}
} // namespace


int wbem_solve_system_device(wbem_ctx *ctx, double *d_phi, double *d_dphi_dn, const double *d_bc,
                             int *iters_out, double *last_res_out)
{
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  const unsigned nb = (N + 255) / 256, nb128 = (N + 127) / 128;
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "solve_system before assemble_system");
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "solve_system before wbem_set_masks");
  EvTimer &g_timer = ctx->timer;
  g_timer.st = st;
  g_timer.on = true;
  g_timer.reset();
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[4], st));
  int rc;
  // alpha (compute_alpha, :833) -- a function of the assembled Neumann matrix only
  if (!ctx->have_alpha)
    {
      g_timer.begin(T_ALPHA);
      rc = wbem_launch_alpha(ctx);
      g_timer.end();
      if (rc) return rc;
    }
  // wbem_gmres: the caller supplies system_rhs as it stands (d_bc = that vector, no unpack)
  const bool rhs_given = (d_phi == nullptr);
  if (rhs_given)
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_rhs, d_bc, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
  // system_rhs (:839) and constrained rows (:845)
  if (!rhs_given)
    {
      g_timer.begin(T_RHS);
      rc = wbem_apply_operator(ctx, 1, d_bc, ctx->d_rhs, false);
      g_timer.end();
      if (rc) return rc;
    }
  if (ctx->p.auto_constraints && !rhs_given)
    { // compute_constraints(constraints, tmp_rhs) (:845)
      g_timer.begin(T_CONSTRAINTS);
      rc = wbem_compute_constraints_device(ctx, d_bc);
      g_timer.end();
      if (rc) return rc;
    }
  if (ctx->n_lines && !rhs_given)
    {
      k_distribute_rhs<<<(ctx->n_lines + 255) / 256, 256, 0, st>>>(ctx->n_lines, ctx->d_con_lines,
                                                                   ctx->d_con_inhom, ctx->d_rhs);
      ctx->launches++;
    }
  // preconditioner (:851)
  g_timer.begin(T_PRECOND_SETUP);
  rc = wbem_build_preconditioner(ctx);
  g_timer.end();
  if (rc) return rc;
  if (ctx->group)
    { // Row blocks that share a device: a block still inside its set-up may sit in a
      // device-synchronising call (cudaMalloc) -- it must not find a peer's mat-vec epilogue
      // already spinning for rows it has yet to launch.  One host rendezvous per solve.
      if ((rc = wbem_group_barrier(ctx))) return rc;
    }

  // solver.solve(cc, sol, system_rhs, preconditioner) (:853), x0 = 0 (:830)
  const int ntmp = ctx->p.gmres_n_tmp_vectors;
  const int m = ntmp - 2;
  const double tol = ctx->p.gmres_tol;
  const int max_steps = ctx->p.gmres_max_steps;
  const size_t ldv = ctx->ld;
  double *V = ctx->d_V, *p = ctx->d_tmp[0], *x = ctx->d_sol;
  double *d_h = ctx->d_h + 1024;     // ctx->d_h[256] initial norm, [512] pure-Neumann norm; d_h: [0,KMAX) y
  double *d_part = d_h + GM_KMAX;    // [DOT_SEGS][KMAX] dot-product partials
  double *d_nrm = d_part + DOT_SEGS * GM_KMAX;  // [0,nb128) per-CTA |w|^2, directly followed by
  double *d_hacc = d_nrm + nb128;               // [0,KMAX) the accumulated Hessenberg column
  double *hp = ctx->h_pinned;
  // device-side Arnoldi state (k_hessenberg): the host enqueues `ahead` (4) iterations at a time and only
  // then reads the stop state back; kernels of iterations past the stop return at once.  The sparse
  // approximate inverse path is all early-exit kernels; the band / host preconditioners keep one
  // synchronisation per iteration.
  const size_t gm_need = 2 + 3 * (size_t)(ntmp + 2) + (size_t)ntmp * ntmp;
  if (ctx->gm_doubles < gm_need)
    {
      if (ctx->d_gm) cudaFree(ctx->d_gm);
      ctx->d_gm = nullptr;
      CUDA_OK(ctx, cudaMalloc((void **)&ctx->d_gm, sizeof(double) * gm_need));
      ctx->gm_doubles = gm_need;
    }
  if (!ctx->d_gm_ctl) CUDA_OK(ctx, cudaMalloc((void **)&ctx->d_gm_ctl, sizeof(int) * 8));
  double *gm = ctx->d_gm;
  int *ctl = ctx->d_gm_ctl;
  double *gm_gamma = gm + 2, *gm_H = gm + 2 + 3 * (size_t)(ntmp + 2);
  const bool spai_path = ctx->p.precond_kind == 1 && ctx->p.preconditioner_band > 0 && !ctx->precond_host_active;
  const int ahead = spai_path ? 4 : 1;
  CUDA_OK(ctx, cudaMemsetAsync(x, 0, sizeof(double) * N, st));
  std::vector<double> H((size_t)ntmp * ntmp, 0.0), gamma(ntmp + 1), y(ntmp + 1);
  int accumulated = 0, state = 0;
  double rho = 0;
  bool x_is_zero = true;
  g_timer.begin(T_GMRES);
  int n_gemv = 0, n_idle = 0; // n_idle: iterations enqueued past the stop (their kernels return at once)
  do
    {
      double *v0 = V;
      if (x_is_zero)
        CUDA_OK(ctx, cudaMemcpyAsync(p, ctx->d_rhs, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
      else
        {
          rc = wbem_apply_operator(ctx, 0, x, p, true);
          ++n_gemv;
          if (rc) return rc;
          k_residual<<<nb, 256, 0, st>>>(N, ctx->d_rhs, p, p);
          ctx->launches++;
        }
      g_timer.begin(T_PRECOND_APPLY);
      rc = wbem_apply_preconditioner(ctx, p, v0);
      g_timer.end();
      if (rc) return rc;
      // rho = |v0|; its inverse (the normalisation folded into the first mat-vec) stays on the device
      CUDA_OK(ctx, cudaMemsetAsync(ctl, 0, sizeof(int) * 8, st));
      k_norm2<<<1, 1024, 0, st>>>(N, v0, ctx->d_h + 256, gm);
      ctx->launches++;
      CUDA_OK(ctx, cudaMemcpyAsync(gm_gamma, ctx->d_h + 256, sizeof(double), cudaMemcpyDeviceToDevice, st)); // gamma[0] = rho
      CUDA_OK(ctx, cudaMemcpyAsync(hp, ctx->d_h + 256, sizeof(double), cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      rho = hp[0];
      state = (rho <= tol) ? 1 : ((accumulated >= max_steps) ? 2 : 0);
      if (!(rho == rho)) state = 2;
      if (state != 0) break;
      int dim = 0;
      ctx->op_scale_ptr = gm;
      ctx->op_stop_ptr = ctl;
      while (dim < m && state == 0)
        {
          int batch = std::min(ahead, m - dim);
          batch = std::min(batch, std::max(1, max_steps - accumulated));
          for (int j = 0; j < batch; ++j)
            {
              const int inner = dim + j;
              double *vv = V + (size_t)(inner + 1) * ldv;
              // vv = M^-1 (A V[inner]): operator, its epilogue and the preconditioner; V[inner] is
              // normalised in place by the mat-vec's prologue (factor gm[0], left by the step before)
              rc = wbem_apply_operator_ex(ctx, 0, V + (size_t)inner * ldv, vv, true, 1.0, V + (size_t)inner * ldv, true);
              ++n_gemv;
              if (rc)
                {
                  ctx->op_scale_ptr = nullptr;
                  ctx->op_stop_ptr = nullptr;
                  return rc;
                }
              // CGS2
              for (int pass = 0; pass < 2; ++pass)
                {
                  HessArgs hs{};
                  if (pass == 1)
                    { // the Arnoldi step's small algebra rides on the last CTA of the second projection
                      hs.gm = gm;
                      hs.ctl = ctl;
                      hs.counter = reinterpret_cast<unsigned int *>(ctl + 4);
                      hs.hacc = d_hacc;
                      hs.inner = inner;
                      hs.ntmp = ntmp;
                      hs.accumulated = accumulated + j + 1;
                      hs.max_steps = max_steps;
                      hs.tol = tol;
                    }
                  k_dots<<<dim3(inner + 1, DOT_SEGS), 256, 0, st>>>(N, V, ldv, vv, d_part, ctl);
                  k_project_out<<<nb128, 128, sizeof(double) * (inner + 1), st>>>(N, inner + 1, V, ldv, d_part, vv, d_hacc, pass,
                                                                                 d_nrm, ctl, hs);
                  ctx->launches += 2;
                }
            }
          // one look per batch: stop state, iterations done in this cycle, residual estimate
          CUDA_OK(ctx, cudaMemcpyAsync(hp, gm + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
          CUDA_OK(ctx, cudaMemcpyAsync(hp + 8, ctl, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
          CUDA_OK(ctx, cudaStreamSynchronize(st));
          const int *hctl = reinterpret_cast<const int *>(hp + 8);
          const int done = hctl[1] - dim;
          if (done <= 0 || done > batch)
            {
              ctx->op_scale_ptr = nullptr;
              ctx->op_stop_ptr = nullptr;
              WBEM_FAIL(ctx, -2, "GMRES: inconsistent device iteration count (%d after %d)", hctl[1], dim);
            }
          accumulated += done;
          n_idle += batch - done;
          dim = hctl[1];
          state = hctl[0];
          rho = hp[0];
        }
      ctx->op_scale_ptr = nullptr;
      ctx->op_stop_ptr = nullptr;
      // H y = gamma (rotated columns and right-hand side from the device), x += V y
      const size_t hneed = (size_t)(ntmp + 2) + (size_t)dim * ntmp;
      std::vector<double> hbig;
      double *hb = hp;
      if (hneed > ctx->pinned_doubles)
        { // long restart lengths on small problems: pageable staging
          hbig.resize(hneed);
          hb = hbig.data();
        }
      CUDA_OK(ctx, cudaMemcpyAsync(hb, gm_gamma, sizeof(double) * (dim + 1), cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaMemcpyAsync(hb + ntmp + 2, gm_H, sizeof(double) * (size_t)dim * ntmp, cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      for (int i = 0; i <= dim; ++i) gamma[i] = hb[i];
      for (int c = 0; c < dim; ++c)
        for (int i = 0; i <= c; ++i) H[(size_t)i * ntmp + c] = hb[ntmp + 2 + (size_t)c * ntmp + i];
      for (int i = dim - 1; i >= 0; --i)
        {
          double s = gamma[i];
          for (int k = i + 1; k < dim; ++k) s -= H[(size_t)i * ntmp + k] * y[k];
          y[i] = s / H[(size_t)i * ntmp + i];
        }
      for (int i = 0; i < dim; ++i) hp[i] = y[i];
      CUDA_OK(ctx, cudaMemcpyAsync(d_h, hp, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
      // the last Krylov vector of the cycle was never normalised, nor is it used
      k_update_solution<<<nb, 256, sizeof(double) * dim, st>>>(N, dim, V, ldv, d_h, x);
      ctx->launches++;
      CUDA_OK(ctx, cudaStreamSynchronize(st)); // hp is reused next cycle
      x_is_zero = false;
    }
  while (state == 0);
  g_timer.end();
  // unpack (:869-879)
  if (!rhs_given)
    {
      k_unpack<<<nb, 256, 0, st>>>(N, ctx->d_surf, x, d_phi, d_dphi_dn);
      ctx->launches++;
    }
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[5], st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  CUDA_OK(ctx, cudaGetLastError());
  if ((rc = wbem_check_gather_timeout(ctx))) return rc;
  double sums[T_NTAGS];
  int counts[T_NTAGS];
  g_timer.resolve(sums, counts, T_NTAGS);
  g_timer.on = false;
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
  ctx->tm.rhs_ms = sums[T_RHS];
  ctx->tm.precond_setup_ms = sums[T_PRECOND_SETUP];
  ctx->tm.gmres_ms = sums[T_GMRES];
  if (counts[T_ALPHA]) ctx->tm.alpha_ms = sums[T_ALPHA];
  ctx->tm.gemv_ms_sum = sums[T_GEMV];
  ctx->tm.precond_apply_ms_sum = sums[T_PRECOND_APPLY];
  ctx->tm.allgather_ms_sum = sums[T_ALLGATHER];
  ctx->tm.constraints_ms = sums[T_CONSTRAINTS];
  ctx->tm.solve_system_total_ms = ms;
  ctx->tm.gmres_iters = accumulated;
  ctx->tm.gemv_calls = counts[T_GEMV] - n_idle; // mat-vecs that streamed the matrices
  (void)n_gemv;
  if (iters_out) *iters_out = accumulated;
  if (last_res_out) *last_res_out = rho;
  return state == 1 ? 0 : 1;
}

// ---------------------------------------------------------------------------------------
// wbem_solve_system_multi: nrhs right-hand sides, ONE pass over the matrices per GMRES iteration.
//
// FreeSurface::jacobian (reference source/free_surface.cc:4918-4993) runs one full inner GMRES
// (bem.solve_system, :4993) per Jacobian-vector product the outer SPGMR asks for
// (source/dae_time_integrator.cc:375-377), always on the SAME matrices: each of those solves
// re-streams 8 N^2 bytes per iteration.  Here up to WBEM_MULTI_MAX systems advance in lock step:
// every system keeps its own Krylov basis, Hessenberg matrix and stopping test (it is the same
// left-preconditioned GMRES as above, system by system), but the operator is applied to all
// active systems by ONE block mat-vec (k_bem_gemv_multi).  A system that has converged leaves the
// block; the others go on.
// ---------------------------------------------------------------------------------------
namespace
{
struct MultiState
{
  MultiWork w = {nullptr, nullptr, nullptr};
  double *d_V = nullptr;                                // [MAX][ntmp-1][ld]
  double *d_p = nullptr, *d_x = nullptr, *d_rhs = nullptr; // [MAX][ld]
  double *d_inhom = nullptr;                            // [MAX][inhom_cap]
  double *d_small = nullptr;                            // [MAX][small_stride]: y | dot partials | |w|^2 partials | h
  double *d_norm = nullptr;                             // [MAX]
  double *h_pin = nullptr;                              // pinned [MAX][pin_stride]
  double *d_io = nullptr;                               // staging of the caller's [nrhs][N] arrays (grow-only)
  size_t io_cap = 0;
  size_t small_stride = 0, pin_stride = 0, inhom_cap = 0;
  uint32_t ld = 0;
  int ntmp = 0;
};
} // namespace

void wbem_multi_free(wbem_ctx *ctx)
{
  MultiState *m = reinterpret_cast<MultiState *>(ctx->multi);
  if (!m) return;
  for (double *p : {m->w.d_xn, m->w.d_xd, m->w.d_xdiag, m->d_V, m->d_p, m->d_x, m->d_rhs, m->d_inhom, m->d_small, m->d_norm, m->d_io})
    if (p) cudaFree(p);
  if (m->h_pin) cudaFreeHost(m->h_pin);
  delete m;
  ctx->multi = nullptr;
}

static MultiState *multi_state(wbem_ctx *ctx)
{
  MultiState *m = reinterpret_cast<MultiState *>(ctx->multi);
  if (m && m->ld == ctx->ld && m->ntmp == ctx->p.gmres_n_tmp_vectors) return m;
  wbem_multi_free(ctx);
  m = new MultiState();
  ctx->multi = m;
  m->ld = ctx->ld;
  m->ntmp = ctx->p.gmres_n_tmp_vectors;
  const size_t ld = ctx->ld, B = WBEM_MULTI_MAX;
  const size_t nb128 = (ctx->N + 127) / 128;
  m->small_stride = GM_KMAX + DOT_SEGS * GM_KMAX + nb128 + GM_KMAX + 8;
  m->pin_stride = nb128 + GM_KMAX + 8;
  bool ok = true;
  auto al = [&](double **p, size_t n) { ok = ok && cudaMalloc((void **)p, sizeof(double) * n) == cudaSuccess; };
  al(&m->w.d_xn, B * ld);
  al(&m->w.d_xd, B * ld);
  al(&m->w.d_xdiag, B * ld);
  al(&m->d_V, B * (size_t)(m->ntmp - 1) * ld);
  al(&m->d_p, B * ld);
  al(&m->d_x, B * ld);
  al(&m->d_rhs, B * ld);
  al(&m->d_small, B * m->small_stride);
  al(&m->d_norm, B);
  ok = ok && cudaMallocHost((void **)&m->h_pin, sizeof(double) * B * m->pin_stride) == cudaSuccess;
  if (ok)
    { // padding entries [N, ld) of the multiplier slabs are read by the mat-vec: zero for good
      cudaMemsetAsync(m->w.d_xn, 0, sizeof(double) * B * ld, ctx->stream);
      cudaMemsetAsync(m->w.d_xd, 0, sizeof(double) * B * ld, ctx->stream);
      cudaMemsetAsync(m->w.d_xdiag, 0, sizeof(double) * B * ld, ctx->stream);
    }
  if (!ok)
    {
      cudaGetLastError();
      wbem_multi_free(ctx);
      ctx->err = "wbem_solve_system_multi: out of device memory for the block work vectors";
      return nullptr;
    }
  return m;
}

// device staging for the host arrays of wbem_solve_system_multi: allocated once and kept (a cudaMalloc /
// cudaFree pair per call synchronises the device and can stall for hundreds of milliseconds)
double *wbem_multi_io(wbem_ctx *ctx, size_t doubles)
{
  MultiState *m = multi_state(ctx);
  if (!m) return nullptr;
  if (doubles > m->io_cap)
    {
      if (m->d_io) cudaFree(m->d_io);
      m->d_io = nullptr;
      m->io_cap = 0;
      if (cudaMalloc((void **)&m->d_io, sizeof(double) * doubles) != cudaSuccess)
        {
          cudaGetLastError();
          ctx->err = "wbem_solve_system_multi: out of device memory for the staging arrays";
          return nullptr;
        }
      m->io_cap = doubles;
    }
  return m->d_io;
}

MultiWork *wbem_multi_work(wbem_ctx *ctx)
{
  MultiState *m = multi_state(ctx);
  return m ? &m->w : nullptr;
}

static int solve_group(wbem_ctx *ctx, int nb, double *d_phi, double *d_dphi_dn, const double *d_bc, int *iters_out,
                       double *res_out, int *rc_out)
{
  cudaStream_t st = ctx->stream;
  const uint32_t N = ctx->N;
  const unsigned nbk = (N + 255) / 256, nb128 = (N + 127) / 128;
  MultiState *ms = multi_state(ctx);
  if (!ms) return -2;
  const size_t ld = ctx->ld, ldv = ctx->ld;
  const int ntmp = ctx->p.gmres_n_tmp_vectors, m = ntmp - 2;
  const double tol = ctx->p.gmres_tol;
  const int max_steps = ctx->p.gmres_max_steps;
  int rc;
  const double *src[WBEM_MULTI_MAX];
  double *dst[WBEM_MULTI_MAX];
  // system_rhs of every system (:839): one block application of the rhs operator
  for (int b = 0; b < nb; ++b)
    {
      src[b] = d_bc + (size_t)b * N;
      dst[b] = ms->d_rhs + b * ld;
    }
  if ((rc = wbem_apply_operator_multi(ctx, 1, nb, src, dst, false))) return rc;
  // constrained rows (:845): the lines are the same for every system, the inhomogeneities follow tmp_rhs
  bool own_inhom = false;
  if (ctx->p.auto_constraints)
    {
      for (int b = 0; b < nb; ++b)
        {
          if ((rc = wbem_compute_constraints_device(ctx, d_bc + (size_t)b * N))) return rc;
          if (ctx->n_lines > ms->inhom_cap)
            {
              if (b != 0) WBEM_FAIL(ctx, -4, "constraint lines changed between the systems of one block");
              if (ms->d_inhom) cudaFree(ms->d_inhom);
              ms->inhom_cap = ctx->n_lines;
              CUDA_OK(ctx, cudaMalloc((void **)&ms->d_inhom, sizeof(double) * WBEM_MULTI_MAX * ms->inhom_cap));
            }
          if (ctx->n_lines)
            CUDA_OK(ctx, cudaMemcpyAsync(ms->d_inhom + b * ms->inhom_cap, ctx->d_con_inhom, sizeof(double) * ctx->n_lines,
                                         cudaMemcpyDeviceToDevice, st));
        }
      own_inhom = true;
    }
  if (ctx->n_lines)
    for (int b = 0; b < nb; ++b)
      {
        k_distribute_rhs<<<(ctx->n_lines + 255) / 256, 256, 0, st>>>(
          ctx->n_lines, ctx->d_con_lines, own_inhom ? ms->d_inhom + b * ms->inhom_cap : ctx->d_con_inhom, ms->d_rhs + b * ld);
        ctx->launches++;
      }
  if ((rc = wbem_build_preconditioner(ctx))) return rc;
  if (ctx->group && (rc = wbem_group_barrier(ctx))) return rc;

  // per-system GMRES state (host)
  struct Sys
  {
    std::vector<double> H, gamma, ci, si, h, y;
    int accumulated = 0, state = 0, dim = 0;
    double rho = 0;
    bool in_cycle = false;
  };
  std::vector<Sys> sys(nb);
  for (Sys &q : sys)
    {
      q.H.assign((size_t)ntmp * ntmp, 0.0);
      q.gamma.assign(ntmp + 1, 0.0);
      q.ci.assign(ntmp + 1, 0.0);
      q.si.assign(ntmp + 1, 0.0);
      q.h.assign(ntmp + 1, 0.0);
      q.y.assign(ntmp + 1, 0.0);
    }
  auto Vb = [&](int b) { return ms->d_V + (size_t)b * (ntmp - 1) * ldv; };
  auto small = [&](int b) { return ms->d_small + (size_t)b * ms->small_stride; };
  CUDA_OK(ctx, cudaMemsetAsync(ms->d_x, 0, sizeof(double) * WBEM_MULTI_MAX * ld, st));
  bool first_cycle = true;
  for (;;)
    {
      std::vector<int> act;
      for (int b = 0; b < nb; ++b)
        if (sys[b].state == 0) act.push_back(b);
      if (act.empty()) break;
      // residuals of the active systems -> V[0]
      if (first_cycle)
        for (int b : act)
          CUDA_OK(ctx, cudaMemcpyAsync(ms->d_p + b * ld, ms->d_rhs + b * ld, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
      else
        {
          for (size_t k = 0; k < act.size(); ++k)
            {
              src[k] = ms->d_x + act[k] * ld;
              dst[k] = ms->d_p + act[k] * ld;
            }
          if ((rc = wbem_apply_operator_multi(ctx, 0, (int)act.size(), src, dst, true))) return rc;
          for (int b : act)
            {
              k_residual<<<nbk, 256, 0, st>>>(N, ms->d_rhs + b * ld, ms->d_p + b * ld, ms->d_p + b * ld);
              ctx->launches++;
            }
        }
      for (int b : act)
        {
          if ((rc = wbem_apply_preconditioner(ctx, ms->d_p + b * ld, Vb(b)))) return rc;
          k_norm2<<<1, 1024, 0, st>>>(N, Vb(b), ms->d_norm + b);
          ctx->launches++;
        }
      CUDA_OK(ctx, cudaMemcpyAsync(ms->h_pin, ms->d_norm, sizeof(double) * WBEM_MULTI_MAX, cudaMemcpyDeviceToHost, st));
      CUDA_OK(ctx, cudaStreamSynchronize(st));
      for (int b : act)
        {
          Sys &q = sys[b];
          q.rho = ms->h_pin[b];
          q.state = (q.rho <= tol) ? 1 : ((q.accumulated >= max_steps) ? 2 : 0);
          if (!(q.rho == q.rho)) q.state = 2;
          q.dim = 0;
          q.in_cycle = q.state == 0;
          if (q.state == 0)
            {
              q.gamma[0] = q.rho;
              k_scale<<<nbk, 256, 0, st>>>(N, Vb(b), 1.0 / q.rho);
              ctx->launches++;
            }
        }
      for (int inner = 0; inner < m; ++inner)
        {
          act.clear();
          for (int b = 0; b < nb; ++b)
            if (sys[b].state == 0 && sys[b].in_cycle) act.push_back(b);
          if (act.empty()) break;
          for (size_t k = 0; k < act.size(); ++k)
            {
              src[k] = Vb(act[k]) + (size_t)inner * ldv;
              dst[k] = ms->d_p + act[k] * ld;
            }
          if ((rc = wbem_apply_operator_multi(ctx, 0, (int)act.size(), src, dst, true))) return rc;
          const int dim = inner + 1;
          for (int b : act)
            {
              double *vv = Vb(b) + (size_t)(inner + 1) * ldv;
              if ((rc = wbem_apply_preconditioner(ctx, ms->d_p + b * ld, vv))) return rc;
              double *d_part = small(b) + GM_KMAX, *d_nrm = d_part + DOT_SEGS * GM_KMAX, *d_hacc = d_nrm + nb128;
              for (int pass = 0; pass < 2; ++pass)
                {
                  k_dots<<<dim3(dim, DOT_SEGS), 256, 0, st>>>(N, Vb(b), ldv, vv, d_part);
                  k_project_out<<<nb128, 128, sizeof(double) * dim, st>>>(N, dim, Vb(b), ldv, d_part, vv, d_hacc, pass, d_nrm);
                  ctx->launches += 2;
                }
              CUDA_OK(ctx, cudaMemcpyAsync(ms->h_pin + b * ms->pin_stride, d_nrm, sizeof(double) * (nb128 + dim),
                                           cudaMemcpyDeviceToHost, st));
            }
          CUDA_OK(ctx, cudaStreamSynchronize(st)); // ONE host round trip per block iteration
          for (int b : act)
            {
              Sys &q = sys[b];
              const double *hp = ms->h_pin + b * ms->pin_stride;
              ++q.accumulated;
              for (int i = 0; i < dim; ++i) q.h[i] = hp[nb128 + i];
              double ss = 0;
              for (unsigned i = 0; i < nb128; ++i) ss += hp[i];
              q.h[dim] = std::sqrt(ss);
              k_scale<<<nbk, 256, 0, st>>>(N, Vb(b) + (size_t)(inner + 1) * ldv, 1.0 / q.h[dim]);
              ctx->launches++;
              givens(q.h, q.gamma, q.ci, q.si, inner);
              for (int i = 0; i < dim; ++i) q.H[(size_t)i * ntmp + inner] = q.h[i];
              q.rho = std::fabs(q.gamma[dim]);
              q.dim = dim;
              q.state = (q.rho <= tol) ? 1 : ((q.accumulated >= max_steps) ? 2 : 0);
              if (!(q.rho == q.rho)) q.state = 2;
            }
        }
      // H y = gamma, x += V y for every system that took part in this cycle
      for (int b = 0; b < nb; ++b)
        {
          Sys &q = sys[b];
          if (!q.in_cycle) continue;
          q.in_cycle = false;
          const int dim = q.dim;
          if (dim == 0) continue;
          for (int i = dim - 1; i >= 0; --i)
            {
              double t = q.gamma[i];
              for (int k = i + 1; k < dim; ++k) t -= q.H[(size_t)i * ntmp + k] * q.y[k];
              q.y[i] = t / q.H[(size_t)i * ntmp + i];
            }
          double *hp = ms->h_pin + b * ms->pin_stride;
          for (int i = 0; i < dim; ++i) hp[i] = q.y[i];
          CUDA_OK(ctx, cudaMemcpyAsync(small(b), hp, sizeof(double) * dim, cudaMemcpyHostToDevice, st));
          k_update_solution<<<nbk, 256, sizeof(double) * dim, st>>>(N, dim, Vb(b), ldv, small(b), ms->d_x + b * ld);
          ctx->launches++;
        }
      CUDA_OK(ctx, cudaStreamSynchronize(st)); // the pinned staging is reused by the next cycle
      first_cycle = false;
    }
  // unpack (:869-879)
  for (int b = 0; b < nb; ++b)
    {
      k_unpack<<<nbk, 256, 0, st>>>(N, ctx->d_surf, ms->d_x + b * ld, d_phi + (size_t)b * N, d_dphi_dn + (size_t)b * N);
      ctx->launches++;
      if (iters_out) iters_out[b] = sys[b].accumulated;
      if (res_out) res_out[b] = sys[b].rho;
      if (sys[b].state != 1) *rc_out = 1;
    }
  return 0;
}

int wbem_solve_system_multi_device(wbem_ctx *ctx, int nrhs, double *d_phi, double *d_dphi_dn, const double *d_bc,
                                   int *iters, double *last_res)
{
  if (!ctx->assembled) WBEM_FAIL(ctx, -3, "solve_system before assemble_system");
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "solve_system before wbem_set_masks");
  if (nrhs < 1) return 0;
  cudaStream_t st = ctx->stream;
  ctx->timer.st = st;
  ctx->timer.on = true;
  ctx->timer.reset();
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[4], st));
  int rc;
  if (!ctx->have_alpha && (rc = wbem_launch_alpha(ctx))) return rc;
  int not_converged = 0, max_iters = 0;
  for (int g0 = 0; g0 < nrhs; g0 += WBEM_MULTI_MAX)
    {
      const int nb = std::min(WBEM_MULTI_MAX, nrhs - g0);
      rc = solve_group(ctx, nb, d_phi + (size_t)g0 * ctx->N, d_dphi_dn + (size_t)g0 * ctx->N, d_bc + (size_t)g0 * ctx->N,
                       iters ? iters + g0 : nullptr, last_res ? last_res + g0 : nullptr, &not_converged);
      if (rc) return rc;
    }
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[5], st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  CUDA_OK(ctx, cudaGetLastError());
  if ((rc = wbem_check_gather_timeout(ctx))) return rc;
  double sums[T_NTAGS];
  int counts[T_NTAGS];
  ctx->timer.resolve(sums, counts, T_NTAGS);
  ctx->timer.on = false;
  float msf = 0;
  cudaEventElapsedTime(&msf, ctx->ev[4], ctx->ev[5]);
  if (iters)
    for (int b = 0; b < nrhs; ++b) max_iters = std::max(max_iters, iters[b]);
  ctx->tm.gemv_ms_sum = sums[T_GEMV];
  ctx->tm.gemv_calls = counts[T_GEMV];
  ctx->tm.allgather_ms_sum = sums[T_ALLGATHER];
  ctx->tm.solve_system_total_ms = msf;
  ctx->tm.gmres_iters = max_iters;
  return not_converged;
}
