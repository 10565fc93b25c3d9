// precond.cu -- device-resident form of the reference's band preconditioner.
//
// The reference factorises band_system (|row - col| < 50 band of the merged operator,
// source/bem_problem.cc:1107-1149) with UMFPACK on the host and does two sparse triangular
// solves per GMRES iteration.  Here the same band matrix is viewed as BLOCK TRIDIAGONAL with
// 64 x 64 blocks (band/2 <= 64, so all entries fall in adjacent blocks) and factorised by
// the block Thomas algorithm on the device:
//
//     S_0 = A_0,   L_k = B_{k-1} S_{k-1}^{-1},   S_k = A_k - L_k C_{k-1},   U_k = S_k^{-1} C_k
//     forward   y_k = b_k - L_k y_{k-1}
//     diagonal  z_k = S_k^{-1} y_k                     (independent blocks: K CTAs)
//     backward  x_k = z_k - U_k x_{k+1}
//
// (A = diagonal, B = sub-, C = super-diagonal blocks).  Diagonal blocks are inverted by
// Gauss-Jordan with partial pivoting inside the block.  M^{-1} v is the same vector as the
// reference's LU solve up to rounding.  The two sequential sweeps run in one CTA that streams
// the 32 KB blocks through a 3-stage ring of TMA bulk copies.
#include <cstdio>

#include "internal.h"

#define BS 64           // block size
#define BS2 (BS * BS)
#define LDS_ (BS + 1)   // padded shared-memory leading dimension

struct DevPrecond
{
  uint32_t K = 0;
  double *SinvT = nullptr, *Lt = nullptr, *Ut = nullptr; // [K][64*64], column-major blocks
  double *work = nullptr;                                // [K*64]
  int *info = nullptr;
};

__device__ __forceinline__ double band_entry(const double *__restrict__ bandm, uint32_t N, int band,
                                             long r, long i)
{
  if (r >= (long)N || i >= (long)N) return (r == i) ? 1.0 : 0.0; // identity padding
  const long k = i - r + band / 2 - 1;
  return (k >= 0 && k < band) ? bandm[(size_t)r * band + k] : 0.0;
}

__device__ void load_block(double *dst, const double *__restrict__ bandm, uint32_t N, int band,
                           uint32_t bi, uint32_t bj)
{
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x)
    {
      const int a = idx / BS, b = idx % BS;
      dst[a * LDS_ + b] = band_entry(bandm, N, band, (long)bi * BS + a, (long)bj * BS + b);
    }
}

// C = A * B (64x64, padded smem), optionally C = D - A*B.  256 threads, 4x4 per thread.
__device__ void gemm64(double *C, const double *A, const double *B, const double *D)
{
  const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int kk = 0; kk < BS; ++kk)
    {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = A[(4 * ty + i) * LDS_ + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[kk * LDS_ + 4 * tx + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      {
        const int o = (4 * ty + i) * LDS_ + 4 * tx + j;
        C[o] = D ? D[o] - acc[i][j] : acc[i][j];
      }
}

// in-place inverse of the 64x64 matrix S (padded smem) by Gauss-Jordan with partial pivoting
__device__ void invert64(double *S, int *piv, int *info, int blk)
{
  __shared__ int s_p;
  __shared__ double s_pivinv;
  const int tid = threadIdx.x;
  for (int c = 0; c < BS; ++c)
    {
      if (tid < 32)
        { // arg-max |S[r][c]|, r >= c
          double best = -1.0;
          int bi = c;
          for (int r = c + tid; r < BS; r += 32)
            {
              const double v = fabs(S[r * LDS_ + c]);
              if (v > best)
                {
                  best = v;
                  bi = r;
                }
            }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1)
            {
              const double ob = __shfl_xor_sync(0xffffffffu, best, off);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
              if (ob > best || (ob == best && oi < bi))
                {
                  best = ob;
                  bi = oi;
                }
            }
          if (tid == 0)
            {
              s_p = bi;
              piv[c] = bi;
              if (best == 0.0) atomicExch(info, blk * BS + c + 1);
            }
        }
      __syncthreads();
      const int p = s_p;
      if (p != c && tid < BS)
        {
          const double t = S[c * LDS_ + tid];
          S[c * LDS_ + tid] = S[p * LDS_ + tid];
          S[p * LDS_ + tid] = t;
        }
      __syncthreads();
      if (tid == 0)
        {
          s_pivinv = 1.0 / S[c * LDS_ + c];
          S[c * LDS_ + c] = 1.0;
        }
      __syncthreads();
      if (tid < BS) S[c * LDS_ + tid] *= s_pivinv;
      __syncthreads();
      // eliminate column c from every other row: thread -> (row r = tid/4 + 64*?..)
      {
        const int r = tid >> 2, q = tid & 3; // 64 rows x 4 column quarters
        double f = 0.0;
        if (r != c) f = S[r * LDS_ + c];
        __syncthreads();
        if (r != c)
          {
            if (q == 0) S[r * LDS_ + c] = 0.0;
          }
        __syncthreads();
        if (r != c && f != 0.0)
          for (int j = q * 16; j < q * 16 + 16; ++j) S[r * LDS_ + j] = fma(-f, S[c * LDS_ + j], S[r * LDS_ + j]);
        __syncthreads();
      }
    }
  // undo the row interchanges: columns of the inverse, in reverse order
  for (int c = BS - 1; c >= 0; --c)
    {
      const int p = piv[c];
      if (p != c && tid < BS)
        {
          const double t = S[tid * LDS_ + c];
          S[tid * LDS_ + c] = S[tid * LDS_ + p];
          S[tid * LDS_ + p] = t;
        }
      __syncthreads();
    }
}

// store a padded smem block to global, transposed (column-major): out[b*64 + a] = M[a][b]
__device__ void store_block_T(double *__restrict__ out, const double *M)
{
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x)
    {
      const int b = idx / BS, a = idx % BS;
      out[idx] = M[a * LDS_ + b];
    }
}

__global__ void __launch_bounds__(256, 1)
  k_bt_factor(uint32_t N, uint32_t K, int band, const double *__restrict__ bandm,
              double *__restrict__ SinvT, double *__restrict__ Lt, double *__restrict__ Ut, int *info)
{
  extern __shared__ double sm[];
  double *S = sm, *Bm = S + BS * LDS_, *Cm = Bm + BS * LDS_, *W = Cm + BS * LDS_;
  __shared__ int piv[BS];
  load_block(S, bandm, N, band, 0, 0);
  __syncthreads();
  for (uint32_t k = 0; k < K; ++k)
    {
      invert64(S, piv, info, (int)k);
      store_block_T(SinvT + (size_t)k * BS2, S);
      if (k + 1 < K)
        {
          load_block(Bm, bandm, N, band, k + 1, k);
          load_block(Cm, bandm, N, band, k, k + 1);
          __syncthreads();
          gemm64(W, S, Cm, nullptr); // U_k = S_k^{-1} C_k
          __syncthreads();
          store_block_T(Ut + (size_t)k * BS2, W);
          __syncthreads();
          gemm64(W, Bm, S, nullptr); // L_{k+1} = B_k S_k^{-1}
          __syncthreads();
          store_block_T(Lt + (size_t)(k + 1) * BS2, W);
          load_block(Bm, bandm, N, band, k + 1, k + 1); // A_{k+1} (Bm is free: W holds L)
          __syncthreads();
          gemm64(S, W, Cm, Bm); // S_{k+1} = A_{k+1} - L_{k+1} C_k
          __syncthreads();
        }
    }
}

// ---- sequential sweeps --------------------------------------------------------------------
__device__ __forceinline__ uint32_t p_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p_mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void p_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 p_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(p_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void p_mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do
    {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(p_smem_u32(bar)), "r"(parity)
                   : "memory");
    }
  while (!done);
}

#define NSTAGE 3
// forward (dir = +1):  v_0 = w_0,       v_k = w_k - M_k v_{k-1},  k = 1..K-1,   M = Lt
// backward (dir = -1): v_{K-1}=w_{K-1}, v_k = w_k - M_k v_{k+1},  k = K-2..0,   M = Ut
// Mt blocks are column-major: Mt[j*64 + r] = M[r][j].  w may alias v.
__global__ void __launch_bounds__(256, 1)
  k_bt_sweep(uint32_t N, uint32_t K, int dir, const double *__restrict__ Mt, const double *w, double *v)
{
  extern __shared__ __align__(128) double sm[];
  double *buf = sm;                                   // [NSTAGE][4096]
  double *vprev = buf + NSTAGE * BS2;                 // [2][64]
  double *part = vprev + 2 * BS;                      // [4][64]
  uint64_t *bar = reinterpret_cast<uint64_t *>(part + 4 * BS); // [NSTAGE]
  const int tid = threadIdx.x;
  const uint32_t nsteps = K - 1;
  auto blk_of = [&](uint32_t s) -> uint32_t { return dir > 0 ? s + 1 : K - 2 - s; };
  if (tid == 0)
    {
      for (int s = 0; s < NSTAGE; ++s) p_mbar_init(&bar[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
  if (tid == 0)
    for (uint32_t s = 0; s < NSTAGE && s < nsteps; ++s)
      {
        p_mbar_expect_tx(&bar[s], BS2 * sizeof(double));
        p_bulk_g2s(buf + s * BS2, Mt + (size_t)blk_of(s) * BS2, BS2 * sizeof(double), &bar[s]);
      }
  // first block: copy through
  {
    const uint32_t k0 = dir > 0 ? 0 : K - 1;
    if (tid < BS)
      {
        const size_t g = (size_t)k0 * BS + tid;
        const double x = g < N ? w[g] : 0.0;
        if (g < N) v[g] = x;
        vprev[tid] = x;
      }
  }
  __syncthreads();
  const int g4 = tid >> 6, r = tid & 63;
  for (uint32_t s = 0; s < nsteps; ++s)
    {
      const int stage = s % NSTAGE;
      const uint32_t k = blk_of(s);
      p_mbar_wait(&bar[stage], (s / NSTAGE) & 1);
      const double *M = buf + stage * BS2;
      const double *vp = vprev + (s & 1) * BS;
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int j = 0; j < 16; j += 2)
        {
          a0 = fma(M[(16 * g4 + j) * BS + r], vp[16 * g4 + j], a0);
          a1 = fma(M[(16 * g4 + j + 1) * BS + r], vp[16 * g4 + j + 1], a1);
        }
      part[g4 * BS + r] = a0 + a1;
      __syncthreads();
      if (tid < BS)
        {
          const size_t g = (size_t)k * BS + tid;
          const double x = (g < N ? w[g] : 0.0) - ((part[tid] + part[BS + tid]) + (part[2 * BS + tid] + part[3 * BS + tid]));
          if (g < N) v[g] = x;
          vprev[((s + 1) & 1) * BS + tid] = x;
        }
      // this stage's buffer is free for step s + NSTAGE (all threads passed the barrier above)
      if (tid == 0 && s + NSTAGE < nsteps)
        {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          p_mbar_expect_tx(&bar[stage], BS2 * sizeof(double));
          p_bulk_g2s(buf + stage * BS2, Mt + (size_t)blk_of(s + NSTAGE) * BS2, BS2 * sizeof(double), &bar[stage]);
        }
      __syncthreads();
    }
}

// z_k = S_k^{-1} y_k for every block (one CTA of 64 threads per block); y may alias z
__global__ void __launch_bounds__(BS)
  k_bt_diag(uint32_t N, const double *__restrict__ SinvT, const double *y, double *z)
{
  __shared__ double sy[BS];
  const uint32_t k = blockIdx.x;
  const int r = threadIdx.x;
  const size_t g = (size_t)k * BS + r;
  sy[r] = g < N ? y[g] : 0.0;
  __syncthreads();
  const double *M = SinvT + (size_t)k * BS2;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
  for (int j = 0; j < BS; j += 4)
    {
      a0 = fma(M[(j + 0) * BS + r], sy[j + 0], a0);
      a1 = fma(M[(j + 1) * BS + r], sy[j + 1], a1);
      a2 = fma(M[(j + 2) * BS + r], sy[j + 2], a2);
      a3 = fma(M[(j + 3) * BS + r], sy[j + 3], a3);
    }
  if (g < N) z[g] = (a0 + a1) + (a2 + a3);
}

static bool g_attr_done = false;

int wbem_device_precond_factor(wbem_ctx *ctx)
{
  const int band = ctx->p.preconditioner_band;
  if (band / 2 > BS) WBEM_FAIL(ctx, -1, "preconditioner band/2 must be <= %d", BS);
  DevPrecond *dp = reinterpret_cast<DevPrecond *>(ctx->dev_precond);
  const uint32_t K = (ctx->N + BS - 1) / BS;
  if (!dp || dp->K != K)
    {
      wbem_device_precond_free(ctx);
      dp = new DevPrecond();
      dp->K = K;
      ctx->dev_precond = dp;
      CUDA_OK(ctx, cudaMalloc((void **)&dp->SinvT, sizeof(double) * (size_t)K * BS2));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->Lt, sizeof(double) * (size_t)K * BS2));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->Ut, sizeof(double) * (size_t)K * BS2));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->work, sizeof(double) * (size_t)K * BS));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->info, sizeof(int)));
    }
  if (!g_attr_done)
    {
      CUDA_OK(ctx, cudaFuncSetAttribute(k_bt_factor, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * 4 * BS * LDS_)));
      CUDA_OK(ctx, cudaFuncSetAttribute(k_bt_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * (NSTAGE * BS2 + 6 * BS) + 64)));
      g_attr_done = true;
    }
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemsetAsync(dp->info, 0, sizeof(int), st));
  k_bt_factor<<<1, 256, sizeof(double) * 4 * BS * LDS_, st>>>(ctx->N, K, band, ctx->d_band, dp->SinvT,
                                                             dp->Lt, dp->Ut, dp->info);
  ctx->launches++;
  int info = 0;
  CUDA_OK(ctx, cudaMemcpyAsync(&info, dp->info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  CUDA_OK(ctx, cudaGetLastError());
  if (info) WBEM_FAIL(ctx, -6, "band preconditioner: singular diagonal block at row %d", info - 1);
  return 0;
}

int wbem_device_precond_solve(wbem_ctx *ctx, const double *d_in, double *d_out)
{
  DevPrecond *dp = reinterpret_cast<DevPrecond *>(ctx->dev_precond);
  if (!dp) WBEM_FAIL(ctx, -3, "device preconditioner not factorised");
  cudaStream_t st = ctx->stream;
  const size_t smem = sizeof(double) * (NSTAGE * BS2 + 6 * BS) + 64;
  if (dp->K > 1)
    {
      k_bt_sweep<<<1, 256, smem, st>>>(ctx->N, dp->K, +1, dp->Lt, d_in, d_out);
      k_bt_diag<<<dp->K, BS, 0, st>>>(ctx->N, dp->SinvT, d_out, d_out);
      k_bt_sweep<<<1, 256, smem, st>>>(ctx->N, dp->K, -1, dp->Ut, d_out, d_out);
      ctx->launches += 3;
    }
  else
    {
      k_bt_diag<<<1, BS, 0, st>>>(ctx->N, dp->SinvT, d_in, d_out);
      ctx->launches++;
    }
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

void wbem_device_precond_free(wbem_ctx *ctx)
{
  DevPrecond *dp = reinterpret_cast<DevPrecond *>(ctx->dev_precond);
  if (!dp) return;
  cudaFree(dp->SinvT);
  cudaFree(dp->Lt);
  cudaFree(dp->Ut);
  cudaFree(dp->work);
  cudaFree(dp->info);
  delete dp;
  ctx->dev_precond = nullptr;
}
