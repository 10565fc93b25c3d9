// precond.cu -- device-resident form of the reference's band preconditioner.
//
// The reference factorises band_system (the |row - col| < 50 band of the merged operator,
// source/bem_problem.cc:1107-1149) with UMFPACK on the host and runs two sequential sparse
// triangular solves per GMRES iteration.  Here the same band matrix is viewed as BLOCK
// TRIDIAGONAL with 64 x 64 blocks (band/2 <= 64, so every entry falls in adjacent blocks;
// the tail is padded with an identity) and solved by BLOCK CYCLIC REDUCTION:
//
//   level l, active positions 0..n-1 (original block index = position << l):
//     eliminate the odd positions p:   x_p = D_p^-1 (b_p - L_p x_{p-1} - U_p x_{p+1})
//     kept (even) positions q:         P-_q = L_q D_{q-1}^-1,  P+_q = U_q D_{q+1}^-1
//                                      D'_q = D_q - P-_q U_{q-1} - P+_q L_{q+1}
//                                      L'_q = -P-_q L_{q-1},   U'_q = -P+_q U_{q+1}
//                                      b'_q = b_q - P-_q b_{q-1} - P+_q b_{q+1}
//   until one block is left.  log2(K) levels, every level fully parallel over blocks: the
//   factorisation is ~9 rounds of independent 64x64 inversions/GEMMs instead of a chain of K
//   of them, and one application of M^-1 streams ~5 K blocks once (HBM-bound) instead of
//   walking a K-long dependency chain.  Diagonal blocks are inverted by Gauss-Jordan with
//   partial (row) pivoting inside the block.  M^-1 v equals the reference's LU solve up to
//   rounding (checked against the oracle's pivoted band LU in tests/test_gpu_parity.py).
#include <cstdio>
#include <vector>

#include "internal.h"

#define BS 64
#define BS2 (BS * BS)
#define LDP (BS + 1) // padded shared-memory leading dimension

struct BcrLevel
{
  uint32_t n, n_elim, n_kept;
  size_t off_dinvT, off_lT, off_uT; // [n_elim] blocks each (column-major = transposed)
  size_t off_pmT, off_ppT;          // [n_kept] blocks each
};

struct DevPrecond
{
  uint32_t K = 0;
  std::vector<BcrLevel> lev;
  double *pool = nullptr;     // saved blocks of every level + the last inverse
  size_t off_last = 0, pool_blocks = 0;
  double *Lw[2] = {nullptr, nullptr}, *Dw[2] = {nullptr, nullptr}, *Uw[2] = {nullptr, nullptr};
  double *dinv_rm = nullptr;  // [K/2] row-major inverses of the level being processed
  double *work = nullptr;     // [K*64] right-hand side / solution, indexed by original block
  int *info = nullptr;
  unsigned int *barrier = nullptr;
  int fused_grid = 0; // 0 = fused kernel unavailable
};

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double band_entry(const double *__restrict__ bandm, uint32_t N, int band,
                                             long r, long i)
{
  if (r >= (long)N || i >= (long)N) return (r == i) ? 1.0 : 0.0; // identity padding
  const long k = i - r + band / 2 - 1;
  return (k >= 0 && k < band) ? bandm[(size_t)r * band + k] : 0.0;
}

// band rows -> explicit block arrays (row-major 64x64): L = block(k,k-1), D, U = block(k,k+1)
__global__ void __launch_bounds__(256)
  k_band_to_blocks(uint32_t N, uint32_t K, int band, const double *__restrict__ bandm,
                   double *__restrict__ L, double *__restrict__ D, double *__restrict__ U)
{
  const uint32_t k = blockIdx.x;
  const int which = blockIdx.y; // 0 L, 1 D, 2 U
  double *dst = (which == 0 ? L : which == 1 ? D : U) + (size_t)k * BS2;
  const long cb = (long)k + which - 1;
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x)
    {
      const int a = idx / BS, b = idx % BS;
      double v = 0.0;
      if (cb >= 0 && cb < (long)K) v = band_entry(bandm, N, band, (long)k * BS + a, cb * BS + b);
      dst[idx] = v;
    }
}

__device__ __forceinline__ void tile_load(double *t, const double *__restrict__ g)
{ // global row-major 64x64 -> padded smem
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x) t[(idx >> 6) * LDP + (idx & 63)] = g[idx];
}
__device__ __forceinline__ void tile_store(double *__restrict__ g, const double *t)
{
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x) g[idx] = t[(idx >> 6) * LDP + (idx & 63)];
}
__device__ __forceinline__ void tile_store_T(double *__restrict__ g, const double *t)
{ // g[b*64 + a] = t[a][b]
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x) g[idx] = t[(idx & 63) * LDP + (idx >> 6)];
}
__device__ __forceinline__ void tile_zero_store(double *__restrict__ g)
{
  for (int idx = threadIdx.x; idx < BS2; idx += blockDim.x) g[idx] = 0.0;
}

// C = sign * A * B  (+ D if D != nullptr), all padded smem tiles; 256 threads; thread (ty,tx)
// owns rows 4ty..4ty+3 and columns tx, tx+16, tx+32, tx+48 (conflict-free smem reads).
__device__ void gemm64(double *C, const double *A, const double *B, const double *D, double sign)
{
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int kk = 0; kk < BS; ++kk)
    {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = A[(4 * ty + i) * LDP + kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[kk * LDP + tx + 16 * j];
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
        const int o = (4 * ty + i) * LDP + tx + 16 * j;
        const double v = sign * acc[i][j];
        C[o] = D ? D[o] + v : v;
      }
}

// In-place Gauss-Jordan inverse with implicit partial pivoting.  256 threads; thread t keeps
// the 16 entries S[r][16q .. 16q+15] (r = t >> 2, q = t & 3) in registers for all 64 steps.
// Step c: the not-yet-used row with the largest |S[.][c]| becomes the pivot row; no rows are
// swapped -- the permutation is undone when the result is written:
//     inverse[col_of_row[r]][row_of_col[j]] = S[r][j].
__device__ void invert64(const double *Sin, double *Sout, int *info, int tag)
{
  __shared__ double colbuf[2][BS];
  __shared__ double prow[2][BS];
  __shared__ int row_of_col[BS];
  const int t = threadIdx.x, lane = t & 31;
  const int r = t >> 2, q = t & 3;
  double a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = Sin[r * LDP + 16 * q + j];
  bool used = false;
  int my_col = 0;
#pragma unroll
  for (int c = 0; c < BS; ++c)
    {
      const int buf = c & 1;
      const double acol = a[c & 15];
      if (q == (c >> 4)) colbuf[buf][r] = used ? -1.0 : fabs(acol);
      __syncthreads();
      // every warp finds the pivot row redundantly (no second barrier needed)
      double best = colbuf[buf][lane];
      int bi = lane;
      {
        const double v1 = colbuf[buf][lane + 32];
        if (v1 > best)
          {
            best = v1;
            bi = lane + 32;
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
      const int p = bi;
      if (t == 0)
        {
          row_of_col[c] = p;
          if (!(best > 0.0)) atomicExch(info, tag * BS + c + 1);
        }
      // multiplier of my row: S[r][c], held by the lane of my 4-lane group with q == c >> 4
      const double f = __shfl_sync(0xffffffffu, acol, (lane & ~3) | (c >> 4));
      if (r == p)
        {
          const double pivinv = 1.0 / f;
          if (q == (c >> 4)) a[c & 15] = 1.0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            {
              a[j] *= pivinv;
              prow[buf][16 * q + j] = a[j];
            }
          used = true;
          my_col = c;
        }
      __syncthreads();
      if (r != p)
        {
          if (q == (c >> 4)) a[c & 15] = 0.0;
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = fma(-f, prow[buf][16 * q + j], a[j]);
        }
    }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; ++j) Sout[my_col * LDP + row_of_col[16 * q + j]] = a[j];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// factorisation kernels (one level)
// ---------------------------------------------------------------------------------------
// eliminated position p = 2m+1: Dinv = D_p^-1 -> dinv_rm[m] (row-major, for k_bcr_update) and
// DinvT (saved); L_p, U_p saved transposed.
__global__ void __launch_bounds__(256, 1)
  k_bcr_invert(uint32_t n_elim, const double *__restrict__ Lw, const double *__restrict__ Dw,
               const double *__restrict__ Uw, double *__restrict__ dinv_rm, double *__restrict__ dinvT,
               double *__restrict__ lT, double *__restrict__ uT, int *info, int tag_base, int single)
{
  extern __shared__ double sm[];
  double *S = sm, *R = sm + BS * LDP;
  const uint32_t m = blockIdx.x;
  const uint32_t p = single ? 0 : 2 * m + 1;
  tile_load(S, Dw + (size_t)p * BS2);
  __syncthreads();
  invert64(S, R, info, tag_base + (int)m);
  tile_store_T(dinvT + (size_t)m * BS2, R);
  if (!single)
    {
      tile_store(dinv_rm + (size_t)m * BS2, R);
      __syncthreads();
      tile_load(S, Lw + (size_t)p * BS2);
      __syncthreads();
      tile_store_T(lT + (size_t)m * BS2, S);
      __syncthreads();
      tile_load(S, Uw + (size_t)p * BS2);
      __syncthreads();
      tile_store_T(uT + (size_t)m * BS2, S);
    }
}

// kept position q = 2t -> position t of the next level
__global__ void __launch_bounds__(256, 1)
  k_bcr_update(uint32_t n, const double *__restrict__ Lw, const double *__restrict__ Dw,
               const double *__restrict__ Uw, const double *__restrict__ dinv_rm,
               double *__restrict__ Ln, double *__restrict__ Dn, double *__restrict__ Un,
               double *__restrict__ pmT, double *__restrict__ ppT)
{
  extern __shared__ double sm[];
  double *Dt = sm, *X = Dt + BS * LDP, *Y = X + BS * LDP, *Z = Y + BS * LDP;
  const uint32_t t = blockIdx.x, q = 2 * t;
  tile_load(Dt, Dw + (size_t)q * BS2);
  if (q >= 1)
    {
      tile_load(X, Lw + (size_t)q * BS2);
      tile_load(Y, dinv_rm + (size_t)(t - 1) * BS2);
      __syncthreads();
      gemm64(Z, X, Y, nullptr, 1.0); // P- = L_q Dinv_{q-1}
      __syncthreads();
      tile_store_T(pmT + (size_t)t * BS2, Z);
      tile_load(X, Uw + (size_t)(q - 1) * BS2);
      __syncthreads();
      gemm64(Dt, Z, X, Dt, -1.0); // D' -= P- U_{q-1}
      __syncthreads();
      tile_load(X, Lw + (size_t)(q - 1) * BS2);
      __syncthreads();
      gemm64(Y, Z, X, nullptr, -1.0); // L' = -P- L_{q-1}
      __syncthreads();
      tile_store(Ln + (size_t)t * BS2, Y);
      __syncthreads();
    }
  else
    {
      tile_zero_store(Ln + (size_t)t * BS2);
      tile_zero_store(pmT + (size_t)t * BS2);
    }
  if (q + 1 < n)
    {
      tile_load(X, Uw + (size_t)q * BS2);
      tile_load(Y, dinv_rm + (size_t)t * BS2);
      __syncthreads();
      gemm64(Z, X, Y, nullptr, 1.0); // P+ = U_q Dinv_{q+1}
      __syncthreads();
      tile_store_T(ppT + (size_t)t * BS2, Z);
      tile_load(X, Lw + (size_t)(q + 1) * BS2);
      __syncthreads();
      gemm64(Dt, Z, X, Dt, -1.0); // D' -= P+ L_{q+1}
      __syncthreads();
      tile_load(X, Uw + (size_t)(q + 1) * BS2);
      __syncthreads();
      gemm64(Y, Z, X, nullptr, -1.0); // U' = -P+ U_{q+1}
      __syncthreads();
      tile_store(Un + (size_t)t * BS2, Y);
    }
  else
    {
      tile_zero_store(Un + (size_t)t * BS2);
      tile_zero_store(ppT + (size_t)t * BS2);
    }
  __syncthreads();
  tile_store(Dn + (size_t)t * BS2, Dt);
}

// ---------------------------------------------------------------------------------------
// solve kernels.  w[K*64] is indexed by ORIGINAL block (position p of level l = block p << l),
// so the reduced right-hand sides and the solution all live in place.
// Blocks are column-major: Mt[j*64 + r] = M[r][j]  ->  thread r reads coalesced.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double gemv_half(const double *__restrict__ Mt, const double *v, int r, int j0)
{ // sum_{j=j0}^{j0+31} M[r][j] v[j]
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 8
  for (int j = j0; j < j0 + 32; j += 4)
    {
      a0 = fma(Mt[(j + 0) * BS + r], v[j + 0], a0);
      a1 = fma(Mt[(j + 1) * BS + r], v[j + 1], a1);
      a2 = fma(Mt[(j + 2) * BS + r], v[j + 2], a2);
      a3 = fma(Mt[(j + 3) * BS + r], v[j + 3], a3);
    }
  return (a0 + a1) + (a2 + a3);
}

// forward: kept q = 2t: w_q -= P-_q w_{q-1} + P+_q w_{q+1}.  256 threads = 64 rows x 4 parts
// (part 0,1: halves of P-, part 2,3: halves of P+).
__global__ void __launch_bounds__(256)
  k_bcr_forward(uint32_t n, int shift, const double *__restrict__ pmT, const double *__restrict__ ppT,
                double *w)
{
  __shared__ double vl[BS], vr[BS], part[4][BS];
  const uint32_t t = blockIdx.x, q = 2 * t;
  const int r = threadIdx.x & 63, part_id = threadIdx.x >> 6;
  const bool has_l = q >= 1, has_r = q + 1 < n;
  if (threadIdx.x < BS) vl[r] = has_l ? w[((size_t)(q - 1) << shift) * BS + r] : 0.0;
  else if (threadIdx.x < 2 * BS) vr[r] = has_r ? w[((size_t)(q + 1) << shift) * BS + r] : 0.0;
  __syncthreads();
  double s = 0.0;
  if (part_id < 2)
    {
      if (has_l) s = gemv_half(pmT + (size_t)t * BS2, vl, r, 32 * part_id);
    }
  else if (has_r)
    s = gemv_half(ppT + (size_t)t * BS2, vr, r, 32 * (part_id - 2));
  part[part_id][r] = s;
  __syncthreads();
  if (threadIdx.x < BS)
    {
      double *dst = w + ((size_t)q << shift) * BS + r;
      *dst = *dst - ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r]));
    }
}

// backward: eliminated p = 2m+1: w_p = Dinv_p (w_p - L_p w_{p-1} - U_p w_{p+1})
__global__ void __launch_bounds__(256)
  k_bcr_backward(uint32_t n, int shift, const double *__restrict__ dinvT, const double *__restrict__ lT,
                 const double *__restrict__ uT, double *w)
{
  __shared__ double vl[BS], vr[BS], rhs[BS], part[4][BS];
  const uint32_t m = blockIdx.x, p = 2 * m + 1;
  const int r = threadIdx.x & 63, part_id = threadIdx.x >> 6;
  const bool has_r = p + 1 < n;
  if (threadIdx.x < BS) vl[r] = w[((size_t)(p - 1) << shift) * BS + r];
  else if (threadIdx.x < 2 * BS) vr[r] = has_r ? w[((size_t)(p + 1) << shift) * BS + r] : 0.0;
  __syncthreads();
  double s = 0.0;
  if (part_id < 2) s = gemv_half(lT + (size_t)m * BS2, vl, r, 32 * part_id);
  else if (has_r) s = gemv_half(uT + (size_t)m * BS2, vr, r, 32 * (part_id - 2));
  part[part_id][r] = s;
  __syncthreads();
  double *dst = w + ((size_t)p << shift) * BS;
  if (threadIdx.x < BS) rhs[r] = dst[r] - ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r]));
  __syncthreads();
  if (part_id < 2) part[part_id][r] = gemv_half(dinvT + (size_t)m * BS2, rhs, r, 32 * part_id);
  __syncthreads();
  if (threadIdx.x < BS) dst[r] = part[0][r] + part[1][r];
}

// last block: w_0 = Dinv w_0 ; also used for K == 1
__global__ void __launch_bounds__(128) k_bcr_last(const double *__restrict__ dinvT, double *w)
{
  __shared__ double rhs[BS], part[2][BS];
  const int r = threadIdx.x & 63, part_id = threadIdx.x >> 6;
  if (threadIdx.x < BS) rhs[r] = w[r];
  __syncthreads();
  part[part_id][r] = gemv_half(dinvT, rhs, r, 32 * part_id);
  __syncthreads();
  if (threadIdx.x < BS) w[r] = part[0][r] + part[1][r];
}

__global__ void k_pad_copy_in(uint32_t N, uint32_t Np, const double *__restrict__ in, double *__restrict__ w)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) w[i] = i < N ? in[i] : 0.0;
}

// ---------------------------------------------------------------------------------------
// Fused application of M^-1: one cooperative kernel walks all levels, with a grid-wide
// barrier between levels instead of one launch per level (the per-level work is a few
// microseconds, so launch gaps dominated the un-fused version).
// ---------------------------------------------------------------------------------------
#define BCR_MAX_LEVELS 24
struct BcrSolveArgs
{
  uint32_t N, K, nlev;
  uint32_t n[BCR_MAX_LEVELS];
  const double *pmT[BCR_MAX_LEVELS], *ppT[BCR_MAX_LEVELS], *dinvT[BCR_MAX_LEVELS], *lT[BCR_MAX_LEVELS],
    *uT[BCR_MAX_LEVELS];
  const double *lastT;
  const double *in;
  double *out; // also the work array: must hold K*64 doubles
  unsigned int *barrier; // zeroed before every launch
};

__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch)
{
  __syncthreads();
  if (threadIdx.x == 0)
    {
      ++epoch;
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned int target = epoch * gridDim.x;
      unsigned int v;
      do
        {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        }
      while (v < target);
      __threadfence();
    }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) k_bcr_solve_fused(const BcrSolveArgs a)
{
  __shared__ double vl[BS], vr[BS], rhs[BS], part[4][BS];
  unsigned int epoch = 0;
  const int r = threadIdx.x & 63, part_id = threadIdx.x >> 6;
  // The output vector doubles as the work array (every caller's buffer holds K*64 doubles);
  // level 0 reads the input directly, so there is no separate copy-in pass or barrier.
  double *w = a.out;
  auto in_val = [&](size_t blk, int r) -> double {
    const size_t g = blk * BS + r;
    return g < a.N ? a.in[g] : 0.0;
  };
  if (a.nlev == 0)
    {
      if (blockIdx.x == 0 && threadIdx.x < BS) w[r] = in_val(0, r);
      __syncthreads();
    }
  // forward reduction of the right-hand side
  for (uint32_t l = 0; l < a.nlev; ++l)
    {
      const uint32_t n = a.n[l], n_kept = (n + 1) / 2;
      for (uint32_t t = blockIdx.x; t < n_kept; t += gridDim.x)
        {
          const uint32_t q = 2 * t;
          const bool has_l = q >= 1, has_r = q + 1 < n;
          double own = 0.0;
          if (l == 0)
            {
              if (threadIdx.x < BS) vl[r] = has_l ? in_val(q - 1, r) : 0.0;
              else if (threadIdx.x < 2 * BS)
                {
                  vr[r] = has_r ? in_val(q + 1, r) : 0.0;
                  if (has_r) w[(size_t)(q + 1) * BS + r] = vr[r]; // odd blocks: plain copy
                }
              if (threadIdx.x < BS) own = in_val(q, r);
            }
          else
            {
              if (threadIdx.x < BS) vl[r] = has_l ? w[((size_t)(q - 1) << l) * BS + r] : 0.0;
              else if (threadIdx.x < 2 * BS) vr[r] = has_r ? w[((size_t)(q + 1) << l) * BS + r] : 0.0;
              if (threadIdx.x < BS) own = w[((size_t)q << l) * BS + r];
            }
          __syncthreads();
          double s = 0.0;
          if (part_id < 2)
            {
              if (has_l) s = gemv_half(a.pmT[l] + (size_t)t * BS2, vl, r, 32 * part_id);
            }
          else if (has_r)
            s = gemv_half(a.ppT[l] + (size_t)t * BS2, vr, r, 32 * (part_id - 2));
          part[part_id][r] = s;
          __syncthreads();
          if (threadIdx.x < BS)
            w[((size_t)q << l) * BS + r] = own - ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r]));
          __syncthreads();
        }
      grid_barrier(a.barrier, epoch);
    }
  // last block
  if (blockIdx.x == 0)
    {
      if (threadIdx.x < BS) rhs[r] = w[r];
      __syncthreads();
      if (part_id < 2) part[part_id][r] = gemv_half(a.lastT, rhs, r, 32 * part_id);
      __syncthreads();
      if (threadIdx.x < BS) w[r] = part[0][r] + part[1][r];
    }
  grid_barrier(a.barrier, epoch);
  // back substitution
  for (uint32_t l = a.nlev; l-- > 0;)
    {
      const uint32_t n = a.n[l], n_elim = n / 2;
      for (uint32_t m = blockIdx.x; m < n_elim; m += gridDim.x)
        {
          const uint32_t p = 2 * m + 1;
          const bool has_r = p + 1 < n;
          if (threadIdx.x < BS) vl[r] = w[((size_t)(p - 1) << l) * BS + r];
          else if (threadIdx.x < 2 * BS) vr[r] = has_r ? w[((size_t)(p + 1) << l) * BS + r] : 0.0;
          __syncthreads();
          double s = 0.0;
          if (part_id < 2) s = gemv_half(a.lT[l] + (size_t)m * BS2, vl, r, 32 * part_id);
          else if (has_r) s = gemv_half(a.uT[l] + (size_t)m * BS2, vr, r, 32 * (part_id - 2));
          part[part_id][r] = s;
          __syncthreads();
          double *dst = w + ((size_t)p << l) * BS;
          if (threadIdx.x < BS) rhs[r] = dst[r] - ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r]));
          __syncthreads();
          if (part_id < 2) part[part_id][r] = gemv_half(a.dinvT[l] + (size_t)m * BS2, rhs, r, 32 * part_id);
          __syncthreads();
          if (threadIdx.x < BS) dst[r] = part[0][r] + part[1][r];
          __syncthreads();
        }
      if (l > 0) grid_barrier(a.barrier, epoch);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------


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
      size_t blocks = 0;
      for (uint32_t n = K; n > 1; n = (n + 1) / 2)
        {
          BcrLevel L;
          L.n = n;
          L.n_elim = n / 2;
          L.n_kept = (n + 1) / 2;
          L.off_dinvT = blocks;
          blocks += L.n_elim;
          L.off_lT = blocks;
          blocks += L.n_elim;
          L.off_uT = blocks;
          blocks += L.n_elim;
          L.off_pmT = blocks;
          blocks += L.n_kept;
          L.off_ppT = blocks;
          blocks += L.n_kept;
          dp->lev.push_back(L);
        }
      dp->off_last = blocks++;
      dp->pool_blocks = blocks;
      CUDA_OK(ctx, cudaMalloc((void **)&dp->pool, sizeof(double) * blocks * BS2));
      for (int i = 0; i < 2; ++i)
        {
          const size_t nb = i == 0 ? K : (K + 1) / 2;
          CUDA_OK(ctx, cudaMalloc((void **)&dp->Lw[i], sizeof(double) * nb * BS2));
          CUDA_OK(ctx, cudaMalloc((void **)&dp->Dw[i], sizeof(double) * nb * BS2));
          CUDA_OK(ctx, cudaMalloc((void **)&dp->Uw[i], sizeof(double) * nb * BS2));
        }
      CUDA_OK(ctx, cudaMalloc((void **)&dp->dinv_rm, sizeof(double) * (size_t)(K / 2 + 1) * BS2));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->work, sizeof(double) * (size_t)K * BS));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->info, sizeof(int)));
      CUDA_OK(ctx, cudaMalloc((void **)&dp->barrier, sizeof(unsigned int)));
      int coop = 0, n_sm = 0, per_sm = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->dev);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bcr_solve_fused, 256, 0);
      dp->fused_grid = (coop && per_sm >= 1 && dp->lev.size() <= BCR_MAX_LEVELS) ? n_sm : 0;
    }
  if (!ctx->bcr_attr_done)
    {
      CUDA_OK(ctx, cudaFuncSetAttribute(k_bcr_invert, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * 2 * BS * LDP)));
      CUDA_OK(ctx, cudaFuncSetAttribute(k_bcr_update, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * 4 * BS * LDP)));
      ctx->bcr_attr_done = true;
    }
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemsetAsync(dp->info, 0, sizeof(int), st));
  k_band_to_blocks<<<dim3(K, 3), 256, 0, st>>>(ctx->N, K, band, ctx->d_band, dp->Lw[0], dp->Dw[0], dp->Uw[0]);
  ctx->launches++;
  int cur = 0, tag = 0;
  // level 0 works on the K-block arrays; from level 1 on the two half-size arrays alternate
  double *Lc = dp->Lw[0], *Dc = dp->Dw[0], *Uc = dp->Uw[0];
  for (size_t l = 0; l < dp->lev.size(); ++l)
    {
      const BcrLevel &L = dp->lev[l];
      k_bcr_invert<<<L.n_elim, 256, sizeof(double) * 2 * BS * LDP, st>>>(
        L.n_elim, Lc, Dc, Uc, dp->dinv_rm, dp->pool + L.off_dinvT * BS2, dp->pool + L.off_lT * BS2,
        dp->pool + L.off_uT * BS2, dp->info, tag, 0);
      tag += L.n_elim;
      // next-level arrays: level 0 -> Lw[1]; afterwards alternate between Lw[1] and Lw[0]
      const int nxt = (l == 0) ? 1 : 1 - cur;
      k_bcr_update<<<L.n_kept, 256, sizeof(double) * 4 * BS * LDP, st>>>(
        L.n, Lc, Dc, Uc, dp->dinv_rm, dp->Lw[nxt], dp->Dw[nxt], dp->Uw[nxt], dp->pool + L.off_pmT * BS2,
        dp->pool + L.off_ppT * BS2);
      ctx->launches += 2;
      cur = nxt;
      Lc = dp->Lw[cur];
      Dc = dp->Dw[cur];
      Uc = dp->Uw[cur];
    }
  k_bcr_invert<<<1, 256, sizeof(double) * 2 * BS * LDP, st>>>(1, Lc, Dc, Uc, dp->dinv_rm,
                                                             dp->pool + dp->off_last * BS2, nullptr, nullptr,
                                                             dp->info, tag, 1);
  ctx->launches++;
  int info = 0;
  CUDA_OK(ctx, cudaMemcpyAsync(&info, dp->info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  CUDA_OK(ctx, cudaGetLastError());
  if (info) WBEM_FAIL(ctx, -6, "band preconditioner: singular diagonal block (code %d)", info);
  return 0;
}

int wbem_device_precond_solve(wbem_ctx *ctx, const double *d_in, double *d_out)
{
  DevPrecond *dp = reinterpret_cast<DevPrecond *>(ctx->dev_precond);
  if (!dp) WBEM_FAIL(ctx, -3, "device preconditioner not factorised");
  cudaStream_t st = ctx->stream;
  const uint32_t Np = dp->K * BS;
  if (dp->fused_grid > 0)
    {
      BcrSolveArgs a;
      a.N = ctx->N;
      a.K = dp->K;
      a.nlev = (uint32_t)dp->lev.size();
      for (size_t l = 0; l < dp->lev.size(); ++l)
        {
          const BcrLevel &L = dp->lev[l];
          a.n[l] = L.n;
          a.pmT[l] = dp->pool + L.off_pmT * BS2;
          a.ppT[l] = dp->pool + L.off_ppT * BS2;
          a.dinvT[l] = dp->pool + L.off_dinvT * BS2;
          a.lT[l] = dp->pool + L.off_lT * BS2;
          a.uT[l] = dp->pool + L.off_uT * BS2;
        }
      a.lastT = dp->pool + dp->off_last * BS2;
      a.in = d_in;
      a.out = d_out;
      a.barrier = dp->barrier;
      CUDA_OK(ctx, cudaMemsetAsync(dp->barrier, 0, sizeof(unsigned int), st));
      void *args[] = {(void *)&a};
      int grid = dp->fused_grid;
      const int most = (int)((dp->K + 1) / 2);
      if (grid > most) grid = most < 1 ? 1 : most;
      CUDA_OK(ctx, cudaLaunchCooperativeKernel((void *)k_bcr_solve_fused, dim3(grid), dim3(256), args, 0, st));
      ctx->launches++;
      return 0;
    }
  k_pad_copy_in<<<(Np + 255) / 256, 256, 0, st>>>(ctx->N, Np, d_in, dp->work);
  ctx->launches++;
  for (size_t l = 0; l < dp->lev.size(); ++l)
    {
      const BcrLevel &L = dp->lev[l];
      k_bcr_forward<<<L.n_kept, 256, 0, st>>>(L.n, (int)l, dp->pool + L.off_pmT * BS2,
                                             dp->pool + L.off_ppT * BS2, dp->work);
      ctx->launches++;
    }
  k_bcr_last<<<1, 128, 0, st>>>(dp->pool + dp->off_last * BS2, dp->work);
  ctx->launches++;
  for (size_t l = dp->lev.size(); l-- > 0;)
    {
      const BcrLevel &L = dp->lev[l];
      k_bcr_backward<<<L.n_elim, 256, 0, st>>>(L.n, (int)l, dp->pool + L.off_dinvT * BS2,
                                              dp->pool + L.off_lT * BS2, dp->pool + L.off_uT * BS2, dp->work);
      ctx->launches++;
    }
  CUDA_OK(ctx, cudaMemcpyAsync(d_out, dp->work, sizeof(double) * ctx->N, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

void wbem_device_precond_free(wbem_ctx *ctx)
{
  DevPrecond *dp = reinterpret_cast<DevPrecond *>(ctx->dev_precond);
  if (!dp) return;
  cudaFree(dp->pool);
  for (int i = 0; i < 2; ++i)
    {
      cudaFree(dp->Lw[i]);
      cudaFree(dp->Dw[i]);
      cudaFree(dp->Uw[i]);
    }
  cudaFree(dp->dinv_rm);
  cudaFree(dp->work);
  cudaFree(dp->info);
  cudaFree(dp->barrier);
  delete dp;
  ctx->dev_precond = nullptr;
}
