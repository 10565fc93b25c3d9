// assemble.cu -- sm_100a kernels for BEMProblem<3>::assemble_system
// (reference source/bem_problem.cc:106-590) and compute_alpha (:594-618).
//
//   k_cell_geometry      FEValues of the regular rule for every cell (:133-137, 192-196),
//                        folded into per-(cell,q) constants:  y_q, n_q JxW_q /(-4 pi), JxW_q/(4 pi)
//   k_assemble_tiled     regular (node, cell) pairs (:241-260, 531-537): one CTA per
//                        (256-row tile, cell cluster); panel data staged in shared memory by a
//                        TMA bulk copy; per-column accumulators in shared memory; deterministic
//                        STORE/ADD flush (see plan.cpp)
//   k_assemble_simple    same integrals, literal reference arithmetic, global atomics: the
//                        independent cross-check and the fallback for quadrature orders != 4
//   k_assemble_singular  pairs whose cell holds a dof of double_nodes_set[i] (:223-230,
//                        261-525): QGaussOneOverR rule, one warp per row, shuffle reduction
//   k_alpha_rowsum       alpha = -row sums of the Neumann matrix (:594-618)
#include <cstdio>
#include <cstring>
#include <vector>

#include "internal.h"
#include "q1map.cuh"

#define FOUR_PI 12.566370614359172953850573533118

struct DevTables
{
  int nq, ns, n1, pad;
  double g_u[WBEM_MAX_NQ], g_v[WBEM_MAX_NQ], g_w[WBEM_MAX_NQ];
  double g_shape[4][WBEM_MAX_NQ];
  double g1_x[8];
  double s_u[4][WBEM_MAX_NS], s_v[4][WBEM_MAX_NS], s_w[4][WBEM_MAX_NS];
};
// The tables live in global memory, one copy per context (two contexts with different
// quadrature orders may share a device); kernels get the pointer as an argument.

int wbem_upload_tables(wbem_ctx *ctx)
{
  std::vector<DevTables> tv(1); // large: keep off the stack
  DevTables &t = tv[0];
  const QuadTables &q = ctx->qt;
  t.nq = q.nq;
  t.ns = q.ns;
  t.n1 = q.n1;
  t.pad = 0;
  memcpy(t.g_u, q.g_u, sizeof(t.g_u));
  memcpy(t.g_v, q.g_v, sizeof(t.g_v));
  memcpy(t.g_w, q.g_w, sizeof(t.g_w));
  memcpy(t.g_shape, q.g_shape, sizeof(t.g_shape));
  memcpy(t.g1_x, q.g1_x, sizeof(t.g1_x));
  memcpy(t.s_u, q.s_u, sizeof(t.s_u));
  memcpy(t.s_v, q.s_v, sizeof(t.s_v));
  memcpy(t.s_w, q.s_w, sizeof(t.s_w));
  if (!ctx->d_tables) CUDA_OK(ctx, cudaMalloc(&ctx->d_tables, sizeof(DevTables)));
  CUDA_OK(ctx, cudaMemcpy(ctx->d_tables, &t, sizeof(t), cudaMemcpyHostToDevice));
  return 0;
}

// One thread per (cell position, q).  Output layout per cell: [7][nq] =
// y_x, y_y, y_z, nJ_x, nJ_y, nJ_z, wJ   with  nJ = n JxW / (-4 pi),  wJ = JxW / (4 pi).
// n JxW = +-(d_u x d_v) w_q exactly (no normalisation needed).
__global__ void k_cell_geometry(const DevTables *__restrict__ qt, uint32_t C, int nq, const double *__restrict__ xyz,
                                const uint32_t *__restrict__ cell_dofs,
                                const uint8_t *__restrict__ dir, double *__restrict__ geo)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c = t / nq;
  const int q = t - c * nq;
  if (c >= C) return;
  QuadVerts X;
  load_verts(xyz, cell_dofs + 4 * (size_t)c, X);
  double y[3], cr[3], phi[4];
  map_q1(X, qt->g_u[q], qt->g_v[q], y, cr, phi);
  const double w = qt->g_w[q];
  const double sgn = dir[c] ? 1.0 : -1.0;
  const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
  double *g = geo + (size_t)c * 7 * nq;
  g[0 * nq + q] = y[0];
  g[1 * nq + q] = y[1];
  g[2 * nq + q] = y[2];
  const double f = sgn * w * (-1.0 / FOUR_PI);
  g[3 * nq + q] = cr[0] * f;
  g[4 * nq + q] = cr[1] * f;
  g[5 * nq + q] = cr[2] * f;
  g[6 * nq + q] = cn * w * (1.0 / FOUR_PI);
}

// The literal FEValues of the regular rule handed over by the caller (reference :192-196:
// get_quadrature_points / get_normal_vectors / JxW), caller's cell order, folded into the same
// per-(cell,q) constants as k_cell_geometry.
__global__ void k_fevalues_to_geometry(uint32_t C, int nq, const uint32_t *__restrict__ cell_order,
                                       const double *__restrict__ qp, const double *__restrict__ nrm,
                                       const double *__restrict__ jxw, double *__restrict__ geo)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = t / nq; // processing position
  const int q = t - p * nq;
  if (p >= C) return;
  const size_t src = (size_t)cell_order[p] * nq + q;
  double *g = geo + (size_t)p * 7 * nq;
  const double w = jxw[src];
  g[0 * nq + q] = qp[3 * src + 0];
  g[1 * nq + q] = qp[3 * src + 1];
  g[2 * nq + q] = qp[3 * src + 2];
  g[3 * nq + q] = nrm[3 * src + 0] * w * (-1.0 / FOUR_PI);
  g[4 * nq + q] = nrm[3 * src + 1] * w * (-1.0 / FOUR_PI);
  g[5 * nq + q] = nrm[3 * src + 2] * w * (-1.0 / FOUR_PI);
  g[6 * nq + q] = w * (1.0 / FOUR_PI);
}

int wbem_upload_fevalues(wbem_ctx *ctx, const double *q_points, const double *normals, const double *JxW)
{
  const int nq = ctx->qt.nq;
  const size_t n = (size_t)ctx->C * nq;
  if (n == 0) return 0;
  double *d = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)&d, sizeof(double) * 7 * n));
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemcpyAsync(d, q_points, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(d + 3 * n, normals, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
  CUDA_OK(ctx, cudaMemcpyAsync(d + 6 * n, JxW, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  k_fevalues_to_geometry<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ctx->C, nq, ctx->d_cell_order, d, d + 3 * n,
                                                                     d + 6 * n, ctx->d_cellgeo);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int wbem_launch_geometry(wbem_ctx *ctx)
{
  const int nq = ctx->qt.nq;
  const uint32_t total = ctx->C * nq;
  if (total == 0) return 0;
  k_cell_geometry<<<(total + 255) / 256, 256, 0, ctx->stream>>>((const DevTables *)ctx->d_tables, ctx->C, nq, ctx->d_xyz,
                                                               ctx->d_cell_dofs, ctx->d_dir,
                                                               ctx->d_cellgeo);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------
// fast 1/sqrt(a): MUFU.RSQ64H seed (rel. error ~2^-22) + one third-order (Householder) step
//   e = 1 - a y^2 ;  y <- y + y e (1/2 + 3/8 e)        -> rel. error O(e^3) < 2^-60
// 5 FP64-pipe instructions instead of the ~10 + branch of the library rsqrt().
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_h3(double a)
{
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double t = a * y;
  const double e = fma(-t, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double ye = y * e;
  return fma(p, ye, y);
}

// a + b and a - b through the FMA datapath (bit-identical results: a*1 + b is rounded once).
// WBEM_FMA_ADDS selects it; see the opcode probe (wbem_issue_probe 100..103) for the reason.
__device__ __forceinline__ double add64(double a, double b)
{
#ifdef WBEM_FMA_ADDS
  double r;
  asm("fma.rn.f64 %0, %1, 0d3FF0000000000000, %2;" : "=d"(r) : "d"(a), "d"(b));
  return r;
#else
  return a + b;
#endif
}
__device__ __forceinline__ double sub64(double a, double b)
{
#ifdef WBEM_FMA_ADDS
  double r;
  asm("fma.rn.f64 %0, %1, 0dBFF0000000000000, %2;" : "=d"(r) : "d"(b), "d"(a));
  return r;
#else
  return a - b;
#endif
}

__global__ void k_rsqrt_selftest(const double *__restrict__ in, double *__restrict__ out, int n)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rsqrt_h3(in[i]);
}

// ---------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes,
                                              uint64_t *bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
      smem_u32(dst)),
    "l"(src), "r"(bytes), "r"(smem_u32(bar))
    : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t done;
  do
    {
      asm volatile("{\n\t.reg .pred p;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(bar)), "r"(parity)
                   : "memory");
    }
  while (!done);
}

// ---------------------------------------------------------------------------------------
// Tiled regular-pair kernel, Gauss 4x4.
//
// One CTA = (64-row tile) x (one cell cluster), 128 threads: TWO threads per collocation
// row, each integrating two of the four Gauss lines (8 of the 16 points) of every cell.
// They swap their partial moments with one shuffle; thread 0 of the pair finishes the
// Neumann sums, thread 1 the Dirichlet sums, and each adds its 4 values to per-column
// accumulators in shared memory (acc[matrix][slot][row]).  TILE_MIN_CTAS CTAs are resident
// per SM (one warp of every CTA on every scheduler) so one CTA's prologue / flush overlaps
// the others' FP64 loops.  Panel data arrives in chunks of TILE_CHUNK cells through a
// double-buffered TMA bulk copy; the LAST warp to finish a chunk (shared-memory counter)
// re-arms the barrier and issues the refill, so no warp ever spins and the integration loop
// has no CTA-wide barrier.
// ---------------------------------------------------------------------------------------
#define TILE_ROWS WBEM_TILE_ROWS          // 64
#define TILE_THREADS (2 * TILE_ROWS)      // 128
#define TILE_WARPS (TILE_THREADS / 32)
#define TILE_W WBEM_TILE_W
#define TILE_MAX_CELLS 64                 // bit mask of singular cells is 64 bits wide
#ifndef WBEM_TILE_CHUNK
#define WBEM_TILE_CHUNK 8
#endif
#ifndef WBEM_TILE_MIN_CTAS
#define WBEM_TILE_MIN_CTAS 3
#endif
#define TILE_CHUNK WBEM_TILE_CHUNK
#define ACC_STRIDE (TILE_ROWS + 1)
// offset between the two matrices' accumulators: = 8 (mod 16) doubles, so that the N-thread and
// the D-thread of a row hit disjoint shared-memory banks
#define ACC_MATOFF (TILE_W * ACC_STRIDE + ((8 - (TILE_W * ACC_STRIDE) % 16) + 16) % 16)

struct TiledArgs
{
  const double *xyz;         // [N][3]
  const double *geo;         // [C][7][16] processing order
  const uint8_t *cell_slots; // [C][4]
  const uint32_t *cl_cell_ptr, *cl_slot_ptr, *slot_col, *color_clusters;
  const uint32_t *sing_ptr, *sing_cellpos; // CSR by local row
  const uint8_t *tile_sing;                // [row tiles][clusters]: any singular pair inside?
  double *Nm, *Dm;
  double *alpha_part; // [n_clusters + 1][nloc]: per-cluster partial of sum_j N_ij (row sum = sum of moments S)
  uint32_t ld, row0, nloc, cluster_base, n_clusters;
  double g1_x[4]; // nodes of the 1-D Gauss rule (kernel parameters sit in the constant bank)
};

constexpr size_t tiled_smem_bytes()
{
  return sizeof(double) * (2 * ACC_MATOFF + 2 * TILE_CHUNK * 7 * 16) + 40 /*mbar + counters*/ + TILE_MAX_CELLS * 4 +
         ((TILE_W + 3) / 4) * 16;
}

#define TILE_BLOCK TILE_THREADS

// one column value of the flush: STORE for the first writer of a column, RED.ADD.F64 for later
// ones -- both predicated, so a warp with mixed lanes runs one instruction stream
__device__ __forceinline__ void flush_value(double *p, double v, uint32_t is_add, uint32_t is_store)
{
  asm volatile("{\n\t.reg .pred pa, ps;\n\t"
               "setp.ne.u32 pa, %2, 0;\n\t"
               "setp.ne.u32 ps, %3, 0;\n\t"
               "@pa red.global.add.f64 [%0], %1;\n\t"
               "@ps st.global.f64 [%0], %1;\n\t}" ::"l"(p),
               "d"(v), "r"(is_add), "r"(is_store)
               : "memory");
}

__global__ void __launch_bounds__(TILE_BLOCK, WBEM_TILE_MIN_CTAS) k_assemble_tiled(const TiledArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *acc = reinterpret_cast<double *>(smem_raw);              // [2][ACC_MATOFF]
  double *geo = acc + 2 * ACC_MATOFF;                              // [2][CHUNK][7][16]
  uint64_t *bar = reinterpret_cast<uint64_t *>(geo + 2 * TILE_CHUNK * 112); // [2] "full" + [2] "empty" barriers
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(bar + 4);         // [2] warps done with a buffer
  uint32_t *s_slots = reinterpret_cast<uint32_t *>(bar + 5);       // [cells] packed 4 x u8
  uint32_t *s_col = s_slots + TILE_MAX_CELLS;                      // [W]

  const int tid = threadIdx.x;
  const int row_l = tid >> 1, h = tid & 1;
  const uint32_t cluster = a.color_clusters[a.cluster_base + blockIdx.x];
  const uint32_t p0 = a.cl_cell_ptr[cluster], p1 = a.cl_cell_ptr[cluster + 1];
  const uint32_t s0 = a.cl_slot_ptr[cluster], s1 = a.cl_slot_ptr[cluster + 1];
  const int ncell = (int)(p1 - p0), nslot = (int)(s1 - s0);
  const int nchunk = (ncell + TILE_CHUNK - 1) / TILE_CHUNK;
  const uint32_t lrow_base = blockIdx.y * TILE_ROWS;
  const uint32_t lrow = lrow_base + row_l;
  const uint32_t lrow_c = lrow < a.nloc ? lrow : a.nloc - 1;

  auto issue_chunk = [&](int c) {
    const int nc = min(TILE_CHUNK, ncell - c * TILE_CHUNK);
    const uint32_t bytes = (uint32_t)nc * 112 * sizeof(double);
    mbar_expect_tx(&bar[c & 1], bytes);
    bulk_copy_g2s(geo + (c & 1) * TILE_CHUNK * 112, a.geo + ((size_t)p0 + (size_t)c * TILE_CHUNK) * 112, bytes,
                  &bar[c & 1]);
  };
  if (tid == 0)
    {
      mbar_init(&bar[0], 1); // "full": TMA bytes landed
      mbar_init(&bar[1], 1);
      mbar_init(&bar[2], TILE_THREADS); // "empty": every thread is done reading the buffer
      mbar_init(&bar[3], TILE_THREADS);
      s_cnt[0] = 0;
      s_cnt[1] = 0;
      issue_chunk(0);
      if (nchunk > 1) issue_chunk(1);
    }
  // singular cells of this row inside the cluster -> bit mask (they are integrated by
  // k_assemble_singular only, reference :241/:261).  Most (row tile, cluster) pairs hold no
  // singular pair at all: a host-built byte map lets them skip the list walk.
  unsigned long long smask = 0ull;
  if (a.tile_sing[(size_t)blockIdx.y * a.n_clusters + cluster])
    {
      const uint32_t b = a.sing_ptr[lrow_c], e = a.sing_ptr[lrow_c + 1];
      for (uint32_t k = b; k < e; k += 4)
        {
          uint32_t pos[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pos[i] = (k + i < e) ? a.sing_cellpos[k + i] : 0xffffffffu;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (pos[i] >= p0 && pos[i] < p1) smask |= 1ull << (pos[i] - p0);
        }
    }
  const double xi0 = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 0];
  const double xi1 = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 1];
  const double xi2 = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 2];
  for (int i = tid; i < ncell; i += TILE_THREADS)
    s_slots[i] = reinterpret_cast<const uint32_t *>(a.cell_slots)[p0 + i];
  for (int i = tid; i < nslot; i += TILE_THREADS) s_col[i] = a.slot_col[s0 + i];
  // zero the accumulators in use: thread (row, h) clears matrix h of its row
  double *accM = acc + h * ACC_MATOFF + row_l;
  for (int s = 0; s < nslot; ++s) accM[s * ACC_STRIDE] = 0.0;
  __syncthreads(); // barriers initialised, slot tables visible

  const double vq0 = a.g1_x[2 * h], vq1 = a.g1_x[2 * h + 1];
  double row_sum = 0.0; // sum over this cluster's regular cells of the zeroth moment (h = 0: Neumann)
  for (int c = 0; c < nchunk; ++c)
    {
      mbar_wait(&bar[c & 1], (c >> 1) & 1);
      const double *gc = geo + (c & 1) * TILE_CHUNK * 112 + 8 * h; // this thread's 8 points
      const int kend = min(TILE_CHUNK, ncell - c * TILE_CHUNK);
      for (int kk = 0; kk < kend; ++kk)
        {
          const int k = c * TILE_CHUNK + kk;
          const double *g = gc + kk * 112;
          // accumulator slots of this cell: loaded now, added to after the integration (the
          // host routes meshes whose cells repeat a dof to the simple kernel)
          const uint32_t sl = s_slots[k];
          double *const pa = accM + (sl & 0xff) * ACC_STRIDE, *const pb = accM + ((sl >> 8) & 0xff) * ACC_STRIDE,
                        *const pc = accM + ((sl >> 16) & 0xff) * ACC_STRIDE, *const pd = accM + (sl >> 24) * ACC_STRIDE;
          const double oa = *pa, ob = *pb, oc = *pc, od = *pd;
          double SN, SuN, SvN, SuvN, SD, SuD, SvD, SuvD;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            {
              double t0n = 0, t1n = 0, t0d = 0, t1d = 0;
#pragma unroll
              for (int qx = 0; qx < 4; ++qx)
                {
                  const int q = j * 4 + qx;
                  const double Rx = sub64(g[q], xi0);
                  const double Ry = sub64(g[16 + q], xi1);
                  const double Rz = sub64(g[32 + q], xi2);
                  const double r2 = fma(Rz, Rz, fma(Ry, Ry, Rx * Rx));
                  const double ri = rsqrt_h3(r2);
                  const double ri2 = ri * ri;
                  const double ri3 = ri2 * ri;
                  const double Rn = fma(Rz, g[80 + q], fma(Ry, g[64 + q], Rx * g[48 + q]));
                  const double av = Rn * ri3;       // (D . n) JxW
                  const double bv = g[96 + q] * ri; // d JxW
                  const double uq = a.g1_x[qx];
                  if (qx == 0)
                    {
                      t0n = av;
                      t1n = av * uq;
                      t0d = bv;
                      t1d = bv * uq;
                    }
                  else
                    {
                      t0n = add64(t0n, av);
                      t1n = fma(av, uq, t1n);
                      t0d = add64(t0d, bv);
                      t1d = fma(bv, uq, t1d);
                    }
                }
              if (j == 0)
                {
                  SN = t0n;
                  SuN = t1n;
                  SvN = vq0 * t0n;
                  SuvN = vq0 * t1n;
                  SD = t0d;
                  SuD = t1d;
                  SvD = vq0 * t0d;
                  SuvD = vq0 * t1d;
                }
              else
                {
                  SN = add64(SN, t0n);
                  SuN = add64(SuN, t1n);
                  SvN = fma(vq1, t0n, SvN);
                  SuvN = fma(vq1, t1n, SuvN);
                  SD = add64(SD, t0d);
                  SuD = add64(SuD, t1d);
                  SvD = fma(vq1, t0d, SvD);
                  SuvD = fma(vq1, t1d, SuvD);
                }
            }
          // swap with the other thread of the row: h = 0 keeps Neumann, h = 1 keeps Dirichlet
          const double o0 = __shfl_xor_sync(0xffffffffu, h ? SN : SD, 1);
          const double o1 = __shfl_xor_sync(0xffffffffu, h ? SuN : SuD, 1);
          const double o2 = __shfl_xor_sync(0xffffffffu, h ? SvN : SvD, 1);
          const double o3 = __shfl_xor_sync(0xffffffffu, h ? SuvN : SuvD, 1);
          const double m0 = add64(h ? SD : SN, o0), m1 = add64(h ? SuD : SuN, o1), m2 = add64(h ? SvD : SvN, o2),
                       m3 = add64(h ? SuvD : SuvN, o3);
          // moments -> the four Q1 shape-function sums
          const double v3 = m3, v1 = sub64(m1, m3), v2 = sub64(m2, m3), v0 = sub64(sub64(m0, m1), v2);
          if (!((smask >> k) & 1ull))
            {
              row_sum = add64(row_sum, m0); // sum_j phi_j = 1: the row sum of the cell's four entries is its S moment
              *pa = add64(oa, v0);
              *pb = add64(ob, v1);
              *pc = add64(oc, v2);
              *pd = add64(od, v3);
            }
        }
      if (c + 2 < nchunk)
        { // this warp is done with the buffer; the last of the CTA's warps refills it
          // every thread releases its reads on the "empty" barrier; the shared counter only
          // elects the warp that arrived last -- its wait on the completed phase returns at
          // once and acquires all the releases
          mbar_arrive(&bar[2 + (c & 1)]);
          __syncwarp();
          if ((tid & 31) == 0)
            {
              if (atomicAdd(&s_cnt[c & 1], 1u) == TILE_WARPS - 1)
                {
                  s_cnt[c & 1] = 0;
                  mbar_wait(&bar[2 + (c & 1)], (c >> 1) & 1);
                  // generic-proxy reads of the buffer are ordered before the async-proxy refill
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  issue_chunk(c + 2);
                }
            }
        }
    }
  if (h == 0 && lrow < a.nloc) a.alpha_part[(size_t)cluster * a.nloc + lrow] = row_sum;
  __syncwarp(); // a warp flushes exactly the 16 rows its own lanes accumulated

  // flush: warp w owns rows [16w, 16w+16) of the tile; lanes run over the cluster's column
  // slots (STORE slots first: one coalesced row segment), 8 rows per lane in flight.
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t row_w = lrow_base + warp * 16;
  if (row_w >= a.nloc) return;
  const int nrw = min(16, (int)(a.nloc - row_w));
  for (int s = lane; s < nslot; s += 32)
    {
      const uint32_t cc = s_col[s];
      const uint32_t is_add = cc >> 31, is_store = is_add ^ 1u;
      double *gN = a.Nm + (size_t)row_w * a.ld + (cc & 0x7fffffffu);
      double *gD = a.Dm + (size_t)row_w * a.ld + (cc & 0x7fffffffu);
      const double *an = acc + s * ACC_STRIDE + warp * 16;
      if (nrw == 16)
        {
#pragma unroll
          for (int rb = 0; rb < 16; rb += 8)
            {
              double vn[8], vd[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                {
                  vn[i] = an[rb + i];
                  vd[i] = an[ACC_MATOFF + rb + i];
                }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                {
                  flush_value(gN, vn[i], is_add, is_store);
                  flush_value(gD, vd[i], is_add, is_store);
                  gN += a.ld;
                  gD += a.ld;
                }
            }
        }
      else
        for (int i = 0; i < nrw; ++i)
          {
            flush_value(gN, an[i], is_add, is_store);
            flush_value(gD, an[ACC_MATOFF + i], is_add, is_store);
            gN += a.ld;
            gD += a.ld;
          }
    }
}

// ---------------------------------------------------------------------------------------
// Simple kernel: one thread per local row, literal reference arithmetic, atomics.
// grid = (row blocks, cell chunks); matrices must be zeroed first.
// ---------------------------------------------------------------------------------------
#define SIMPLE_ROWS 128
#define SIMPLE_TC 4

__global__ void __launch_bounds__(SIMPLE_ROWS)
  k_assemble_simple(const DevTables *__restrict__ qt, int nq, uint32_t C, uint32_t cells_per_chunk, const double *__restrict__ xyz,
                    const double *__restrict__ geo, const uint32_t *__restrict__ cell_dofs,
                    const uint32_t *__restrict__ colpos, const uint32_t *__restrict__ sing_ptr,
                    const uint32_t *__restrict__ sing_cellpos, double *Nm, double *Dm, uint32_t ld,
                    uint32_t row0, uint32_t nloc)
{
  extern __shared__ __align__(16) double sgeo[]; // [TC][7][nq]
  __shared__ uint32_t scol[SIMPLE_TC][4];
  const int tid = threadIdx.x;
  const uint32_t lrow = blockIdx.x * SIMPLE_ROWS + tid;
  const bool row_ok = lrow < nloc;
  const uint32_t lrow_c = row_ok ? lrow : nloc - 1;
  const uint32_t c_begin = blockIdx.y * cells_per_chunk;
  const uint32_t c_end = min(C, c_begin + cells_per_chunk);
  const double xi[3] = {xyz[3 * (size_t)(row0 + lrow_c)], xyz[3 * (size_t)(row0 + lrow_c) + 1],
                        xyz[3 * (size_t)(row0 + lrow_c) + 2]};
  uint32_t sp = sing_ptr[lrow_c];
  const uint32_t se = sing_ptr[lrow_c + 1];
  while (sp < se && sing_cellpos[sp] < c_begin) ++sp;
  uint32_t next_sing = sp < se ? sing_cellpos[sp] : 0xffffffffu;

  for (uint32_t cb = c_begin; cb < c_end; cb += SIMPLE_TC)
    {
      const int nc = min((uint32_t)SIMPLE_TC, c_end - cb);
      __syncthreads();
      for (int i = tid; i < nc * 7 * nq; i += SIMPLE_ROWS) sgeo[i] = geo[(size_t)cb * 7 * nq + i];
      if (tid < nc * 4) scol[tid / 4][tid % 4] = colpos[cell_dofs[4 * (size_t)cb + tid]];
      __syncthreads();
      for (int k = 0; k < nc; ++k)
        {
          const uint32_t cpos = cb + k;
          if (cpos == next_sing)
            {
              ++sp;
              next_sing = sp < se ? sing_cellpos[sp] : 0xffffffffu;
              continue;
            }
          const double *g = sgeo + k * 7 * nq;
          double ln[4] = {0, 0, 0, 0}, ldd[4] = {0, 0, 0, 0};
          for (int q = 0; q < nq; ++q)
            {
              const double Rx = g[q] - xi[0], Ry = g[nq + q] - xi[1], Rz = g[2 * nq + q] - xi[2];
              const double r = sqrt(Rx * Rx + Ry * Ry + Rz * Rz);
              const double r2 = r * r;
              // folded constants: nJ carries n JxW/(-4 pi), wJ carries JxW/(4 pi)
              const double Dn = (Rx * g[3 * nq + q] + Ry * g[4 * nq + q] + Rz * g[5 * nq + q]) / (r2 * r);
              const double s = g[6 * nq + q] / r;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                {
                  ln[j] += Dn * qt->g_shape[j][q];
                  ldd[j] += s * qt->g_shape[j][q];
                }
            }
          if (row_ok)
            {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                {
                  atomicAdd(Nm + (size_t)lrow * ld + scol[k][j], ln[j]);
                  atomicAdd(Dm + (size_t)lrow * ld + scol[k][j], ldd[j]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Singular pairs: one warp per local row; lanes run over the 2 n^2 polar points; literal
// reference arithmetic (LaplaceKernel::kernels, include/laplace_kernel.h:55-58).  The warp
// shuffle-reduces the 8 integrals of a pair and lanes 0..7 add them into row i (the regular
// kernels have finished on this stream; no other warp touches row i).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
  k_assemble_singular(const DevTables *__restrict__ qt, const double *__restrict__ xyz, const uint32_t *__restrict__ cell_dofs,
                      const uint8_t *__restrict__ dir, const uint32_t *__restrict__ colpos,
                      const uint32_t *__restrict__ sing_ptr,
                      const uint32_t *__restrict__ sing_cellpos,
                      const uint8_t *__restrict__ sing_idx, double *Nm, double *Dm, uint32_t ld,
                      uint32_t row0, uint32_t nloc, double *__restrict__ alpha_sing)
{
  const int lane = threadIdx.x & 31;
  const uint32_t lrow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (lrow >= nloc) return;
  double row_sum = 0.0;
  const double *px = xyz + 3 * (size_t)(row0 + lrow);
  const double xi[3] = {px[0], px[1], px[2]};
  const int ns = qt->ns;
  for (uint32_t k = sing_ptr[lrow]; k < sing_ptr[lrow + 1]; ++k)
    {
      const uint32_t cpos = sing_cellpos[k];
      const int sj = sing_idx[k];
      const uint32_t *dofs = cell_dofs + 4 * (size_t)cpos;
      QuadVerts X;
      load_verts(xyz, dofs, X);
      const double sgn = dir[cpos] ? 1.0 : -1.0;
      double v8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int q = lane; q < ns; q += 32)
        {
          double y[3], cr[3], phi[4];
          map_q1(X, qt->s_u[sj][q], qt->s_v[sj][q], y, cr, phi);
          const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
          const double jxw = cn * qt->s_w[sj][q];
          const double nx = sgn * cr[0] / cn, ny = sgn * cr[1] / cn, nz = sgn * cr[2] / cn;
          const double Rx = y[0] - xi[0], Ry = y[1] - xi[1], Rz = y[2] - xi[2];
          const double r = sqrt(Rx * Rx + Ry * Ry + Rz * Rz);
          const double r2 = r * r;
          const double s = 1.0 / (r * FOUR_PI);
          const double den = -FOUR_PI * r2 * r;
          const double Dn = (Rx / den) * nx + (Ry / den) * ny + (Rz / den) * nz;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            {
              v8[j] += Dn * phi[j] * jxw;
              v8[4 + j] += s * phi[j] * jxw;
            }
        }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v8[j] += __shfl_xor_sync(0xffffffffu, v8[j], off);
        }
      row_sum += (v8[0] + v8[1]) + (v8[2] + v8[3]);
      // lanes 0..3 -> Neumann, 4..7 -> Dirichlet; sequential over j for repeated dofs
      double mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane == j) mine = v8[j];
      if (lane < 8)
        {
          double *M = (lane < 4 ? Nm : Dm) + (size_t)lrow * ld;
          const uint32_t col = colpos[dofs[lane & 3]];
          // two local dofs of a degenerate cell may share a column: serialise
          for (int j = 0; j < 4; ++j)
            {
              if ((lane & 3) == j) M[col] += mine;
              __syncwarp(0xffu);
            }
        }
      __syncwarp();
    }
  if (lane == 0 && alpha_sing) alpha_sing[lrow] = row_sum;
}

// alpha_loc[r] = -(sum over clusters of the partial row sums + singular part), fixed order
__global__ void __launch_bounds__(256)
  k_alpha_from_parts(const double *__restrict__ part, uint32_t n_parts, uint32_t nloc, double *__restrict__ out)
{
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  uint32_t k = 0;
  for (; k + 4 <= n_parts; k += 4)
    {
      s0 += part[(size_t)(k + 0) * nloc + r];
      s1 += part[(size_t)(k + 1) * nloc + r];
      s2 += part[(size_t)(k + 2) * nloc + r];
      s3 += part[(size_t)(k + 3) * nloc + r];
    }
  for (; k < n_parts; ++k) s0 += part[(size_t)k * nloc + r];
  out[r] = -((s0 + s1) + (s2 + s3));
}

// alpha_loc[r] = - sum_j N[r][j]   (one warp per local row, 128-bit loads)
__global__ void __launch_bounds__(256)
  k_alpha_rowsum(const double *__restrict__ Nm, uint32_t ld, uint32_t nloc, double *__restrict__ out)
{
  const int lane = threadIdx.x & 31;
  const uint32_t lrow = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (lrow >= nloc) return;
  const double2 *row = reinterpret_cast<const double2 *>(Nm + (size_t)lrow * ld);
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const uint32_t n2 = ld / 2;
  uint32_t j = lane;
  for (; j + 96 < n2; j += 128)
    {
      const double2 a = row[j], b = row[j + 32], c = row[j + 64], d = row[j + 96];
      s0 += a.x + a.y;
      s1 += b.x + b.y;
      s2 += c.x + c.y;
      s3 += d.x + d.y;
    }
  for (; j < n2; j += 32)
    {
      const double2 a = row[j];
      s0 += a.x + a.y;
    }
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[lrow] = -s;
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
int wbem_launch_assemble(wbem_ctx *ctx)
{
  cudaStream_t st = ctx->stream;
  const int nq = ctx->qt.nq;
  if (ctx->nloc == 0 || ctx->C == 0) return 0;
  const bool tiled = (ctx->p.assemble_variant == 0) && (ctx->qt.n1 == 4) && !ctx->has_degenerate_cells;
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], st));
  if (tiled)
    {
      const AssemblyPlan &pl = ctx->plan;
      if (!ctx->tiled_attr_set)
        { // per device: a second context on another GPU needs its own opt-in
          CUDA_OK(ctx, cudaFuncSetAttribute(k_assemble_tiled,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)tiled_smem_bytes()));
          ctx->tiled_attr_set = true;
        }
      // columns no cell touches (none on deal.II meshes) stay zero
      if (pl.n_cols_written < ctx->ld)
        {
          const size_t w = (ctx->ld - pl.n_cols_written) * sizeof(double);
          CUDA_OK(ctx, cudaMemset2DAsync(ctx->d_Nm + pl.n_cols_written, ctx->ld * sizeof(double), 0,
                                         w, ctx->nloc, st));
          CUDA_OK(ctx, cudaMemset2DAsync(ctx->d_Dm + pl.n_cols_written, ctx->ld * sizeof(double), 0,
                                         w, ctx->nloc, st));
        }
      TiledArgs a;
      a.xyz = ctx->d_xyz;
      a.geo = ctx->d_cellgeo;
      a.cell_slots = ctx->d_cell_slots;
      a.cl_cell_ptr = ctx->d_cl_cell_ptr;
      a.cl_slot_ptr = ctx->d_cl_slot_ptr;
      a.slot_col = ctx->d_slot_col;
      a.color_clusters = ctx->d_color_clusters;
      a.sing_ptr = ctx->d_sing_ptr;
      a.sing_cellpos = ctx->d_sing_cellpos;
      a.tile_sing = ctx->d_tile_sing;
      a.n_clusters = pl.n_clusters;
      a.Nm = ctx->d_Nm;
      a.Dm = ctx->d_Dm;
      a.alpha_part = ctx->d_alpha_part;
      a.ld = ctx->ld;
      a.row0 = ctx->row0;
      a.nloc = ctx->nloc;
      for (int i = 0; i < 4; ++i) a.g1_x[i] = ctx->qt.g1_x[i];
      const uint32_t row_tiles = (ctx->nloc + TILE_ROWS - 1) / TILE_ROWS;
      for (uint32_t c = 0; c < pl.n_colors; ++c)
        {
          const uint32_t nclu = pl.color_ptr[c + 1] - pl.color_ptr[c];
          if (nclu == 0) continue;
          a.cluster_base = pl.color_ptr[c];
          dim3 grid(nclu, row_tiles);
          k_assemble_tiled<<<grid, TILE_BLOCK, tiled_smem_bytes(), st>>>(a);
          ctx->launches++;
        }
      CUDA_OK(ctx, cudaGetLastError());
    }
  else
    {
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_Nm, 0, sizeof(double) * (size_t)ctx->nloc * ctx->ld, st));
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_Dm, 0, sizeof(double) * (size_t)ctx->nloc * ctx->ld, st));
      const uint32_t row_blocks = (ctx->nloc + SIMPLE_ROWS - 1) / SIMPLE_ROWS;
      uint32_t chunks = (4 * 148 + row_blocks - 1) / row_blocks; // ~4 CTAs per SM in flight
      if (chunks < 1) chunks = 1;
      uint32_t per = (ctx->C + chunks - 1) / chunks;
      per = ((per + SIMPLE_TC - 1) / SIMPLE_TC) * SIMPLE_TC;
      chunks = (ctx->C + per - 1) / per;
      dim3 grid(row_blocks, chunks);
      const size_t sm = sizeof(double) * SIMPLE_TC * 7 * nq;
      k_assemble_simple<<<grid, SIMPLE_ROWS, sm, st>>>((const DevTables *)ctx->d_tables, nq, ctx->C, per, ctx->d_xyz, ctx->d_cellgeo,
                                                       ctx->d_cell_dofs, ctx->d_colpos,
                                                       ctx->d_sing_ptr, ctx->d_sing_cellpos,
                                                       ctx->d_Nm, ctx->d_Dm, ctx->ld, ctx->row0,
                                                       ctx->nloc);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], st));
  if (ctx->n_sing)
    {
      const uint32_t warps_per_block = 8;
      k_assemble_singular<<<(ctx->nloc + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(
        (const DevTables *)ctx->d_tables, ctx->d_xyz, ctx->d_cell_dofs, ctx->d_dir, ctx->d_colpos, ctx->d_sing_ptr,
        ctx->d_sing_cellpos, ctx->d_sing_idx, ctx->d_Nm, ctx->d_Dm, ctx->ld, ctx->row0, ctx->nloc,
        tiled ? ctx->d_alpha_part + (size_t)ctx->plan.n_clusters * ctx->nloc : nullptr);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  else if (tiled)
    CUDA_OK(ctx, cudaMemsetAsync(ctx->d_alpha_part + (size_t)ctx->plan.n_clusters * ctx->nloc, 0,
                                 sizeof(double) * ctx->nloc, st));
  ctx->alpha_parts_valid = tiled;
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], st));
  return 0;
}

int wbem_launch_alpha(wbem_ctx *ctx, bool from_matrix)
{
  cudaStream_t st = ctx->stream;
  if (ctx->nloc && ctx->alpha_parts_valid && !from_matrix)
    {
      k_alpha_from_parts<<<(ctx->nloc + 255) / 256, 256, 0, st>>>(ctx->d_alpha_part, ctx->plan.n_clusters + 1,
                                                                 ctx->nloc,
                                                                 ctx->d_yloc + (size_t)ctx->p.rank * ctx->chunk);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  else if (ctx->nloc)
    {
      k_alpha_rowsum<<<(ctx->nloc + 7) / 8, 256, 0, st>>>(ctx->d_Nm, ctx->ld, ctx->nloc,
                                                         ctx->d_yloc + (size_t)ctx->p.rank * ctx->chunk);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
    }
  int rc = wbem_allgather_rows(ctx, ctx->d_yloc);
  if (rc) return rc;
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_alpha, ctx->d_yloc, sizeof(double) * ctx->N,
                               cudaMemcpyDeviceToDevice, st));
  ctx->have_alpha = true;
  return 0;
}

// exported self test of the fast reciprocal square root (tests/test_gpu_kernels.py)
extern "C" int wbem_selftest_rsqrt(wbem_ctx *ctx, const double *in, double *out, int n)
{
  double *d_in = nullptr, *d_out = nullptr;
  CUDA_OK(ctx, cudaMalloc(&d_in, sizeof(double) * n));
  CUDA_OK(ctx, cudaMalloc(&d_out, sizeof(double) * n));
  CUDA_OK(ctx, cudaMemcpy(d_in, in, sizeof(double) * n, cudaMemcpyHostToDevice));
  k_rsqrt_selftest<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_in, d_out, n);
  ctx->launches++;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaMemcpy(out, d_out, sizeof(double) * n, cudaMemcpyDeviceToHost));
  cudaFree(d_in);
  cudaFree(d_out);
  return 0;
}
