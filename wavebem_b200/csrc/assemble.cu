// assemble.cu -- sm_100a kernels for BEMProblem<3>::assemble_system
// (reference source/bem_problem.cc:106-590) and compute_alpha (:594-618).
//
//   k_cell_geometry      FEValues of the regular rule for every cell (:133-137, 192-196), folded into
//                        per-(cell,q) records  y_q, n_q JxW_q /(-4 pi), JxW_q/(4 pi), JxW_q u_q/(4 pi)
//                        and, for the 4 x 4 rule, into the line records of the stream kernel
//   k_assemble_rows      regular (node, cell) pairs (:241-260, 531-537), the default: ONE persistent
//                        launch; work items (128-row tile, cell cluster) drawn from a ticket counter;
//                        line-wise polynomial arithmetic; panel chunks through a TMA / mbarrier pipeline;
//                        per-column accumulators in shared memory; STORE / ADD flush ordered by done
//                        flags between the clusters sharing a column (see plan.cpp): deterministic
//   k_assemble_colours   the same tiling with per-point arithmetic, one launch per colour
//                        (assemble_variant = 2; caller-supplied FEValues; plans the stream kernel's
//                        records cannot hold)
//   k_assemble_simple    same integrals, literal reference arithmetic, global atomics: the
//                        independent cross-check and the fallback for quadrature orders != 4
//   k_assemble_singular  pairs whose cell holds a dof of double_nodes_set[i] (:223-230,
//                        261-525): QGaussOneOverR rule, one warp per row, shuffle reduction
//   k_alpha_from_parts / k_alpha_rowsum   alpha = -row sums of the Neumann matrix (:594-618)
// Development switches (never set in the product build): -DWBEM_DBG_NOWAIT / _NOFLUSH / _WAITLOG time the
// kernel without its dependency waits / without the flush, or log the spins per work item.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"
#include "q1map.cuh"

#define FOUR_PI 12.566370614359172953850573533118
#define GEO_REC 8 // doubles per (cell, q): y[3], n JxW/(-4 pi)[3], JxW/(4 pi), JxW u_q/(4 pi)
// Line records of the 4 x 4 rule on a bilinear (Q1) cell, for the stream kernel.  Along a Gauss
// line v = v_j the panel point is y(u) = A + u B and d_u y x d_v y = P + u Q, with B . P = B . Q = 0, so
// for a collocation point x and D = A - x
//     r^2(u)             = D.D + u (2 B.D) + u^2 B.B          (quadratic in u)
//     (y - x) . n JxW(u) = w_u (D.P' + u D.Q')                 (linear in u; P', Q' carry w_v / (-4 pi) and the orientation)
// -- 15 FP64 operations per row and line instead of 9 per row and Gauss point.
// Layout per cell: 4 lines x [A(3), 2B(3), B.B, P'(3), Q'(3), pad] then 16 x (JxW/(4 pi), JxW u/(4 pi)).
#define LINE_REC 14
#define GEO2_REC (4 * LINE_REC + 32)

struct DevTables
{
  int nq, ns, n1, pad;
  double g_u[WBEM_MAX_NQ], g_v[WBEM_MAX_NQ], g_w[WBEM_MAX_NQ];
  double g_shape[4][WBEM_MAX_NQ];
  double g1_x[8], g1_w[8];
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
  memcpy(t.g1_w, q.g1_w, sizeof(t.g1_w));
  memcpy(t.s_u, q.s_u, sizeof(t.s_u));
  memcpy(t.s_v, q.s_v, sizeof(t.s_v));
  memcpy(t.s_w, q.s_w, sizeof(t.s_w));
  if (!ctx->d_tables) CUDA_OK(ctx, cudaMalloc(&ctx->d_tables, sizeof(DevTables)));
  CUDA_OK(ctx, cudaMemcpy(ctx->d_tables, &t, sizeof(t), cudaMemcpyHostToDevice));
  return 0;
}

// One thread per (cell position, q).  Output layout per cell: [GEO_REC = 8][nq] =
// y_x, y_y, y_z, nJ_x, nJ_y, nJ_z, wJ, wJ u_q   with  nJ = n JxW / (-4 pi),  wJ = JxW / (4 pi)
// (u_q = the point's first reference coordinate: the first-moment weight of the row kernel).
// n JxW = +-(d_u x d_v) w_q exactly (no normalisation needed).
__global__ void k_cell_geometry(const DevTables *__restrict__ qt, uint32_t C, int nq, const double *__restrict__ xyz,
                                const uint32_t *__restrict__ cell_dofs,
                                const uint8_t *__restrict__ dir, double *__restrict__ geo, double *__restrict__ geo2)
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
  double *g = geo + (size_t)c * GEO_REC * nq;
  g[0 * nq + q] = y[0];
  g[1 * nq + q] = y[1];
  g[2 * nq + q] = y[2];
  const double f = sgn * w * (-1.0 / FOUR_PI);
  g[3 * nq + q] = cr[0] * f;
  g[4 * nq + q] = cr[1] * f;
  g[5 * nq + q] = cr[2] * f;
  g[6 * nq + q] = cn * w * (1.0 / FOUR_PI);
  g[7 * nq + q] = (cn * w * (1.0 / FOUR_PI)) * qt->g_u[q];
  if (geo2 && nq == 16)
    {
      double *h = geo2 + (size_t)c * GEO2_REC;
      h[4 * LINE_REC + 2 * q] = cn * w * (1.0 / FOUR_PI);
      h[4 * LINE_REC + 2 * q + 1] = (cn * w * (1.0 / FOUR_PI)) * qt->g_u[q];
      if ((q & 3) == 0)
        { // first point of Gauss line j: the line's constants
          const int j = q >> 2;
          const double v = qt->g_v[q];
          const double fl = sgn * qt->g1_w[j] * (-1.0 / FOUR_PI);
          double A[3], B[3], cc[3], dd[3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
            {
              cc[d] = X.x[2][d] - X.x[0][d];
              dd[d] = (X.x[3][d] - X.x[2][d]) - (X.x[1][d] - X.x[0][d]);
              A[d] = X.x[0][d] + v * cc[d];
              B[d] = (X.x[1][d] - X.x[0][d]) + v * dd[d];
            }
          double *l = h + j * LINE_REC;
          l[0] = A[0], l[1] = A[1], l[2] = A[2];
          l[3] = 2.0 * B[0], l[4] = 2.0 * B[1], l[5] = 2.0 * B[2];
          l[6] = B[0] * B[0] + B[1] * B[1] + B[2] * B[2];
          l[7] = (B[1] * cc[2] - B[2] * cc[1]) * fl;
          l[8] = (B[2] * cc[0] - B[0] * cc[2]) * fl;
          l[9] = (B[0] * cc[1] - B[1] * cc[0]) * fl;
          l[10] = (B[1] * dd[2] - B[2] * dd[1]) * fl;
          l[11] = (B[2] * dd[0] - B[0] * dd[2]) * fl;
          l[12] = (B[0] * dd[1] - B[1] * dd[0]) * fl;
          l[13] = 0.0;
        }
    }
}

// The literal FEValues of the regular rule handed over by the caller (reference :192-196:
// get_quadrature_points / get_normal_vectors / JxW), caller's cell order, folded into the same
// per-(cell,q) constants as k_cell_geometry.
__global__ void k_fevalues_to_geometry(const DevTables *__restrict__ qt, uint32_t C, int nq,
                                       const uint32_t *__restrict__ cell_order,
                                       const double *__restrict__ qp, const double *__restrict__ nrm,
                                       const double *__restrict__ jxw, double *__restrict__ geo)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = t / nq; // processing position
  const int q = t - p * nq;
  if (p >= C) return;
  const size_t src = (size_t)cell_order[p] * nq + q;
  double *g = geo + (size_t)p * GEO_REC * nq;
  const double w = jxw[src];
  g[0 * nq + q] = qp[3 * src + 0];
  g[1 * nq + q] = qp[3 * src + 1];
  g[2 * nq + q] = qp[3 * src + 2];
  g[3 * nq + q] = nrm[3 * src + 0] * w * (-1.0 / FOUR_PI);
  g[4 * nq + q] = nrm[3 * src + 1] * w * (-1.0 / FOUR_PI);
  g[5 * nq + q] = nrm[3 * src + 2] * w * (-1.0 / FOUR_PI);
  g[6 * nq + q] = w * (1.0 / FOUR_PI);
  g[7 * nq + q] = (w * (1.0 / FOUR_PI)) * qt->g_u[q];
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
  k_fevalues_to_geometry<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const DevTables *)ctx->d_tables, ctx->C, nq,
                                                                     ctx->d_cell_order, d, d + 3 * n,
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
                                                               ctx->d_cellgeo, ctx->d_cellgeo2);
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

#define TILE_MAX_CELLS 64 // bit mask of singular cells is 64 bits wide

struct CtaDesc
{ // one (cluster) work item in launch order: everything the CTA needs from one 16-byte load
  uint32_t p0;      // first cell (processing position)
  uint32_t s0;      // first slot
  uint32_t counts;  // cells | slots << 8
  uint32_t cluster; // cluster id (alpha partials, singular-pair map)
};

struct TiledArgs
{
  const double *xyz;         // [N][3]
  const double *geo;         // [C][GEO_REC][16] processing order
  const uint8_t *cell_slots; // [C][4]
  const CtaDesc *desc;       // [clusters] in launch (colour) order
  const uint32_t *slot_col;
  const uint32_t *sing_ptr, *sing_cellpos; // CSR by local row
  const uint8_t *tile_sing;                // [row tiles][clusters]: any singular pair inside?
  double *Nm, *Dm;
  double *alpha_part; // [n_clusters + 1][nloc]: per-cluster partial of sum_j N_ij (row sum = sum of moments S)
  uint32_t ld, row0, nloc, cluster_base, n_clusters;
  double g1_x[4]; // nodes of the 1-D Gauss rule (kernel parameters sit in the constant bank)
};

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

// ---------------------------------------------------------------------------------------
// Row kernel: the same tiling, ONE thread per collocation row (T1_RPT rows per thread).
//
// A thread integrates all 16 Gauss points of a cell for its row(s) and both kernels, so there is
// no partner thread, no shuffle and no select in the loop: per (cell, row) 16 evaluations cost
// ~390 FP64 instructions and ~90 others (56 broadcast LDS.128 of the panel record -- shared by
// the T1_RPT rows of a thread --, 16 MUFU, 4 + 4 128-bit accumulator updates).  The accumulators
// are acc[slot][row] pairs (N, D) in shared memory: consecutive lanes touch consecutive 16-byte
// words (no bank conflicts in the loop); the slot stride of ROWS + 1 pairs keeps the flush
// (lanes over slots) at the minimum of four wavefronts per 128-bit request.
// ---------------------------------------------------------------------------------------
#ifndef WBEM_T1_THREADS
#define WBEM_T1_THREADS 128
#endif
#ifndef WBEM_T1_RPT
#define WBEM_T1_RPT 1
#endif
#ifndef WBEM_T1_W
#define WBEM_T1_W 32
#endif
#ifndef WBEM_T1_MINCTAS
#define WBEM_T1_MINCTAS 3
#endif
#ifndef WBEM_T1_CHUNK
#define WBEM_T1_CHUNK 4
#endif
#define T1_THREADS WBEM_T1_THREADS
#define T1_RPT WBEM_T1_RPT
#define T1_ROWS (T1_THREADS * T1_RPT)
#define T1_W WBEM_T1_W
#define T1_CHUNK WBEM_T1_CHUNK
#define T1_WARPS (T1_THREADS / 32)
#define T1_STRIDE (T1_ROWS + 1) // double2 units between consecutive slots
#define T1_REC (GEO_REC * 16)   // doubles per cell record

// Moments of one cell for the T1_RPT rows of a thread.  g = the cell's panel record in shared memory
// ([GEO_REC][16], warp-uniform addresses: broadcast loads); u = the 1-D Gauss nodes.  The tensor
// structure of the 4 x 4 rule turns the four shape-function sums into the moments
// S = sum a, Su = sum a u, Sv = sum a v, Suv = sum a u v  (a = kernel value x weight):
// 2.75 instead of 4 FMAs per evaluation and matrix.
struct CellMoments
{
  double SN[T1_RPT], SuN[T1_RPT], SvN[T1_RPT], SuvN[T1_RPT], SD[T1_RPT], SuD[T1_RPT], SvD[T1_RPT], SuvD[T1_RPT];
};

__device__ __forceinline__ void integrate_cell(const double *__restrict__ g, const double (&xi0)[T1_RPT],
                                               const double (&xi1)[T1_RPT], const double (&xi2)[T1_RPT], const double u0,
                                               const double u1, const double u2, const double u3, CellMoments &m)
{
  double(&SN)[T1_RPT] = m.SN, (&SuN)[T1_RPT] = m.SuN, (&SvN)[T1_RPT] = m.SvN, (&SuvN)[T1_RPT] = m.SuvN;
  double(&SD)[T1_RPT] = m.SD, (&SuD)[T1_RPT] = m.SuD, (&SvD)[T1_RPT] = m.SvD, (&SuvD)[T1_RPT] = m.SuvD;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    {
      const double vq = j == 0 ? u0 : j == 1 ? u1 : j == 2 ? u2 : u3;
      double t0n[T1_RPT], t1n[T1_RPT], t0d[T1_RPT], t1d[T1_RPT];
#pragma unroll
      for (int qx = 0; qx < 4; ++qx)
        {
          const int q = j * 4 + qx;
          const double uq = qx == 0 ? u0 : qx == 1 ? u1 : qx == 2 ? u2 : u3;
          const double y0 = g[q], y1 = g[16 + q], y2 = g[32 + q];
          const double n0 = g[48 + q], n1 = g[64 + q], n2 = g[80 + q], wj = g[96 + q], wju = g[112 + q];
#pragma unroll
          for (int r = 0; r < T1_RPT; ++r)
            {
              const double Rx = y0 - xi0[r], Ry = y1 - xi1[r], Rz = y2 - xi2[r];
              const double r2 = fma(Rz, Rz, fma(Ry, Ry, Rx * Rx));
              const double ri = rsqrt_h3(r2);
              const double ri2 = ri * ri;
              const double ri3 = ri2 * ri;
              const double Rn = fma(Rz, n2, fma(Ry, n1, Rx * n0));
              const double av = Rn * ri3; // (D . n) JxW; the single-layer d JxW = wj ri goes
              if (qx == 0)                // straight into its two line moments (wju = wj u_q)
                {
                  t0n[r] = av;
                  t1n[r] = av * uq;
                  t0d[r] = wj * ri;
                  t1d[r] = wju * ri;
                }
              else
                {
                  t0n[r] += av;
                  t1n[r] = fma(av, uq, t1n[r]);
                  t0d[r] = fma(wj, ri, t0d[r]);
                  t1d[r] = fma(wju, ri, t1d[r]);
                }
            }
        }
#pragma unroll
      for (int r = 0; r < T1_RPT; ++r)
        if (j == 0)
          {
            SN[r] = t0n[r];
            SuN[r] = t1n[r];
            SvN[r] = vq * t0n[r];
            SuvN[r] = vq * t1n[r];
            SD[r] = t0d[r];
            SuD[r] = t1d[r];
            SvD[r] = vq * t0d[r];
            SuvD[r] = vq * t1d[r];
          }
        else
          {
            SN[r] += t0n[r];
            SuN[r] += t1n[r];
            SvN[r] = fma(vq, t0n[r], SvN[r]);
            SuvN[r] = fma(vq, t1n[r], SuvN[r]);
            SD[r] += t0d[r];
            SuD[r] += t1d[r];
            SvD[r] = fma(vq, t0d[r], SvD[r]);
            SuvD[r] = fma(vq, t1d[r], SuvD[r]);
          }
    }
}

// The same moments from the line records (GEO2_REC doubles per cell, see the top of the file), one
// Gauss line (v = un[J]) at a time.  wq / wuq: the 1-D Gauss weights and weights x nodes (kernel
// parameters: constant-bank operands).
template <int J>
__device__ __forceinline__ void line_moments(const double *__restrict__ g, const double (&xi0)[T1_RPT],
                                             const double (&xi1)[T1_RPT], const double (&xi2)[T1_RPT], const double (&un)[4],
                                             const double (&wq)[4], const double (&wuq)[4], CellMoments &m)
{
  const double2 *wj2 = reinterpret_cast<const double2 *>(g + 4 * LINE_REC) + J * 4;
  const double2 *L = reinterpret_cast<const double2 *>(g + J * LINE_REC);
  const double2 l0 = L[0], l1 = L[1], l2 = L[2], l3 = L[3], l4 = L[4], l5 = L[5], l6 = L[6];
  const double vq = un[J], bb = l3.x;
  double c0[T1_RPT], c1[T1_RPT], e0[T1_RPT], e1[T1_RPT];
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      const double Dx = l0.x - xi0[r], Dy = l0.y - xi1[r], Dz = l1.x - xi2[r];
      c0[r] = fma(Dz, Dz, fma(Dy, Dy, Dx * Dx));
      c1[r] = fma(Dz, l2.y, fma(Dy, l2.x, Dx * l1.y));
      e0[r] = fma(Dz, l4.y, fma(Dy, l4.x, Dx * l3.y));
      e1[r] = fma(Dz, l6.x, fma(Dy, l5.y, Dx * l5.x));
    }
  double t0n[T1_RPT], t1n[T1_RPT], t0d[T1_RPT], t1d[T1_RPT];
#pragma unroll
  for (int qx = 0; qx < 4; ++qx)
    {
      const double u = un[qx];
      const double2 wj = wj2[qx];
#pragma unroll
      for (int r = 0; r < T1_RPT; ++r)
        {
          const double r2 = fma(fma(bb, u, c1[r]), u, c0[r]);
          const double ri = rsqrt_h3(r2);
          const double ri2 = ri * ri;
          const double Rn = fma(e1[r], u, e0[r]);
          const double av = (Rn * ri) * ri2; // (two products side by side: one level less than r^-3 first)
          if (qx == 0)
            {
              t0n[r] = wq[0] * av;
              t1n[r] = wuq[0] * av;
              t0d[r] = wj.x * ri;
              t1d[r] = wj.y * ri;
            }
          else
            {
              t0n[r] = fma(wq[qx], av, t0n[r]);
              t1n[r] = fma(wuq[qx], av, t1n[r]);
              t0d[r] = fma(wj.x, ri, t0d[r]);
              t1d[r] = fma(wj.y, ri, t1d[r]);
            }
        }
    }
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    if (J == 0)
      {
        m.SN[r] = t0n[r];
        m.SuN[r] = t1n[r];
        m.SvN[r] = vq * t0n[r];
        m.SuvN[r] = vq * t1n[r];
        m.SD[r] = t0d[r];
        m.SuD[r] = t1d[r];
        m.SvD[r] = vq * t0d[r];
        m.SuvD[r] = vq * t1d[r];
      }
    else
      {
        m.SN[r] += t0n[r];
        m.SuN[r] += t1n[r];
        m.SvN[r] = fma(vq, t0n[r], m.SvN[r]);
        m.SuvN[r] = fma(vq, t1n[r], m.SuvN[r]);
        m.SD[r] += t0d[r];
        m.SuD[r] += t1d[r];
        m.SvD[r] = fma(vq, t0d[r], m.SvD[r]);
        m.SuvD[r] = fma(vq, t1d[r], m.SuvD[r]);
      }
}

// NL Gauss lines at once, written stage by stage over their 4 NL points: more independent dependency
// chains in flight per warp (the per-point chain r^2 -> rsqrt -> r^-3 -> moments is ~12 FP64 latencies deep).
template <int J, int NL>
__device__ __forceinline__ void line_group_moments(const double *__restrict__ g, const double (&xi0)[T1_RPT],
                                                   const double (&xi1)[T1_RPT], const double (&xi2)[T1_RPT], const double (&un)[4],
                                                   const double (&wq)[4], const double (&wuq)[4], CellMoments &m)
{
  static_assert(T1_RPT == 1, "written for one row per thread");
  constexpr int NP = 4 * NL;
  double c0[NL], c1[NL], e0[NL], e1[NL], bb[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l)
    {
      const double2 *L = reinterpret_cast<const double2 *>(g + (J + l) * LINE_REC);
      const double2 l0 = L[0], l1 = L[1], l2 = L[2], l3 = L[3], l4 = L[4], l5 = L[5], l6 = L[6];
      bb[l] = l3.x;
      const double Dx = l0.x - xi0[0], Dy = l0.y - xi1[0], Dz = l1.x - xi2[0];
      c0[l] = fma(Dz, Dz, fma(Dy, Dy, Dx * Dx));
      c1[l] = fma(Dz, l2.y, fma(Dy, l2.x, Dx * l1.y));
      e0[l] = fma(Dz, l4.y, fma(Dy, l4.x, Dx * l3.y));
      e1[l] = fma(Dz, l6.x, fma(Dy, l5.y, Dx * l5.x));
    }
  double r2[NP], y[NP], t[NP], e[NP], p[NP], ye[NP], ri[NP], rr[NP], av[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) r2[i] = fma(fma(bb[i >> 2], un[i & 3], c1[i >> 2]), un[i & 3], c0[i >> 2]);
#pragma unroll
  for (int i = 0; i < NP; ++i) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(r2[i]));
#pragma unroll
  for (int i = 0; i < NP; ++i) t[i] = r2[i] * y[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) e[i] = fma(-t[i], y[i], 1.0);
#pragma unroll
  for (int i = 0; i < NP; ++i) p[i] = fma(0.375, e[i], 0.5), ye[i] = y[i] * e[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) ri[i] = fma(p[i], ye[i], y[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) rr[i] = ri[i] * ri[i]; // r^-2
#pragma unroll
  for (int i = 0; i < NP; ++i) av[i] = (fma(e1[i >> 2], un[i & 3], e0[i >> 2]) * ri[i]) * rr[i];
  const double2 *wj2 = reinterpret_cast<const double2 *>(g + 4 * LINE_REC) + J * 4;
#pragma unroll
  for (int l = 0; l < NL; ++l)
    {
      double t0n = wq[0] * av[4 * l], t1n = wuq[0] * av[4 * l];
      const double2 w0 = wj2[4 * l];
      double t0d = w0.x * ri[4 * l], t1d = w0.y * ri[4 * l];
#pragma unroll
      for (int qx = 1; qx < 4; ++qx)
        {
          const double2 wj = wj2[4 * l + qx];
          t0n = fma(wq[qx], av[4 * l + qx], t0n);
          t1n = fma(wuq[qx], av[4 * l + qx], t1n);
          t0d = fma(wj.x, ri[4 * l + qx], t0d);
          t1d = fma(wj.y, ri[4 * l + qx], t1d);
        }
      const double vq = un[J + l];
      if (J + l == 0)
        {
          m.SN[0] = t0n, m.SuN[0] = t1n, m.SvN[0] = vq * t0n, m.SuvN[0] = vq * t1n;
          m.SD[0] = t0d, m.SuD[0] = t1d, m.SvD[0] = vq * t0d, m.SuvD[0] = vq * t1d;
        }
      else
        {
          m.SN[0] += t0n, m.SuN[0] += t1n, m.SvN[0] = fma(vq, t0n, m.SvN[0]), m.SuvN[0] = fma(vq, t1n, m.SuvN[0]);
          m.SD[0] += t0d, m.SuD[0] += t1d, m.SvD[0] = fma(vq, t0d, m.SvD[0]), m.SuvD[0] = fma(vq, t1d, m.SuvD[0]);
        }
    }
}

// moments -> the four Q1 shape-function sums, added to the cell's four column slots of the thread's
// rows.  The four slots of a cell are distinct (cells with a repeated dof take the simple kernel),
// so the four accumulator loads go out together: one shared-memory round trip per cell, not four.
__device__ __forceinline__ void scatter_cell_distinct(double2 *accT, const uint32_t sl, const CellMoments &cm,
                                                      const unsigned long long (&smask)[T1_RPT], const int k,
                                                      double (&row_sum)[T1_RPT])
{
  double2 *const pa = accT + (sl & 0xff) * T1_STRIDE, *const pb = accT + ((sl >> 8) & 0xff) * T1_STRIDE,
                *const pc = accT + ((sl >> 16) & 0xff) * T1_STRIDE, *const pd = accT + (sl >> 24) * T1_STRIDE;
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      if ((smask[r] >> k) & 1ull) continue; // singular pair: k_assemble_singular integrates it
      double2 va = pa[r * T1_THREADS], vb = pb[r * T1_THREADS], vc = pc[r * T1_THREADS], vd = pd[r * T1_THREADS];
      const double n3 = cm.SuvN[r], n1 = cm.SuN[r] - n3, n2 = cm.SvN[r] - n3, n0 = (cm.SN[r] - cm.SuN[r]) - n2;
      const double d3 = cm.SuvD[r], d1 = cm.SuD[r] - d3, d2 = cm.SvD[r] - d3, d0 = (cm.SD[r] - cm.SuD[r]) - d2;
      row_sum[r] += cm.SN[r];
      va.x += n0, va.y += d0;
      vb.x += n1, vb.y += d1;
      vc.x += n2, vc.y += d2;
      vd.x += n3, vd.y += d3;
      pa[r * T1_THREADS] = va;
      pb[r * T1_THREADS] = vb;
      pc[r * T1_THREADS] = vc;
      pd[r * T1_THREADS] = vd;
    }
}

// moments -> the four Q1 shape-function sums, added to the cell's four column slots of the thread's
// rows (accT = the thread's first accumulator; sl = the cell's four slot numbers, one byte each)
__device__ __forceinline__ void scatter_cell(double2 *accT, const uint32_t sl, const CellMoments &cm,
                                             const unsigned long long (&smask)[T1_RPT], const int k, double (&row_sum)[T1_RPT])
{
  double2 *const pa = accT + (sl & 0xff) * T1_STRIDE, *const pb = accT + ((sl >> 8) & 0xff) * T1_STRIDE,
                *const pc = accT + ((sl >> 16) & 0xff) * T1_STRIDE, *const pd = accT + (sl >> 24) * T1_STRIDE;
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      if ((smask[r] >> k) & 1ull) continue; // singular pair: k_assemble_singular integrates it
      const double n3 = cm.SuvN[r], n1 = cm.SuN[r] - n3, n2 = cm.SvN[r] - n3, n0 = (cm.SN[r] - cm.SuN[r]) - n2;
      const double d3 = cm.SuvD[r], d1 = cm.SuD[r] - d3, d2 = cm.SvD[r] - d3, d0 = (cm.SD[r] - cm.SuD[r]) - d2;
      row_sum[r] += cm.SN[r];
      double2 v;
      v = pa[r * T1_THREADS];
      v.x += n0;
      v.y += d0;
      pa[r * T1_THREADS] = v;
      v = pb[r * T1_THREADS];
      v.x += n1;
      v.y += d1;
      pb[r * T1_THREADS] = v;
      v = pc[r * T1_THREADS];
      v.x += n2;
      v.y += d2;
      pc[r * T1_THREADS] = v;
      v = pd[r * T1_THREADS];
      v.x += n3;
      v.y += d3;
      pd[r * T1_THREADS] = v;
    }
}

// flush: warp w owns rows r * T1_THREADS + [32 w, 32 w + 32) of the tile; lanes run over the
// cluster's column slots (STORE slots first: one coalesced row segment per row)
__device__ __forceinline__ void flush_rows(const double2 *acc, const uint32_t *s_col, const int nslot, const uint32_t lrow_base,
                                           const uint32_t nloc, const uint32_t ld, double *Nm, double *Dm, const int warp,
                                           const int lane)
{
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      const uint32_t row_w = lrow_base + r * T1_THREADS + warp * 32;
      if (row_w >= nloc) continue;
      const int nrw = min(32, (int)(nloc - row_w));
      for (int s = lane; s < nslot; s += 32)
        {
          const uint32_t cc = s_col[s];
          const uint32_t is_add = cc >> 31, is_store = is_add ^ 1u;
          double *gN = Nm + (size_t)row_w * ld + (cc & WBEM_SLOT_COL_MASK);
          double *gD = Dm + (size_t)row_w * ld + (cc & WBEM_SLOT_COL_MASK);
          const double2 *an = acc + s * T1_STRIDE + r * T1_THREADS + warp * 32;
          if (nrw == 32)
            {
#pragma unroll
              for (int rb = 0; rb < 32; rb += 8)
                {
                  double2 v[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = an[rb + i];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    {
                      flush_value(gN, v[i].x, is_add, is_store);
                      flush_value(gD, v[i].y, is_add, is_store);
                      gN += ld;
                      gD += ld;
                    }
                }
            }
          else
            for (int i = 0; i < nrw; ++i)
              {
                const double2 v = an[i];
                flush_value(gN, v.x, is_add, is_store);
                flush_value(gD, v.y, is_add, is_store);
                gN += ld;
                gD += ld;
              }
        }
    }
}

// One pass of the stream kernel's flush over n consecutive slots [s_begin, s_begin + n) of the cluster,
// ADD = false: first writers (plain stores, one contiguous row segment per row), ADD = true: RED.ADD.F64.
// Lanes = slots; when n <= 16 the spare lanes take further rows (P rows per step).
template <bool ADD>
__device__ __forceinline__ void flush_pass(const double2 *acc, const uint32_t *s_col, const int s_begin, const int n,
                                           const uint32_t lrow_base, const uint32_t nloc, const uint32_t ld, double *Nm, double *Dm,
                                           const int warp, const int lane)
{
  if (n <= 0) return;
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      const uint32_t row_w = lrow_base + r * T1_THREADS + warp * 32;
      if (row_w >= nloc) continue;
      const int nrw = min(32, (int)(nloc - row_w));
      for (int sb = 0; sb < n; sb += 32)
        {
          const int nn = min(32, n - sb);
          const int P = 32 / nn; // rows per step
          const int part = lane / nn, j = lane - part * nn;
          if (part < P)
            {
              const int s = s_begin + sb + j;
              const uint32_t cc = s_col[s] & WBEM_SLOT_COL_MASK;
              double *gN = Nm + (size_t)(row_w + part) * ld + cc;
              double *gD = Dm + (size_t)(row_w + part) * ld + cc;
              const double2 *an = acc + s * T1_STRIDE + r * T1_THREADS + warp * 32 + part;
              const size_t step = (size_t)P * ld;
#pragma unroll 4
              for (int i = part; i < nrw; i += P)
                {
                  const double2 v = *an;
                  if (ADD)
                    {
                      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(gN), "d"(v.x) : "memory");
                      asm volatile("red.global.add.f64 [%0], %1;" ::"l"(gD), "d"(v.y) : "memory");
                    }
                  else
                    {
                      *gN = v.x;
                      *gD = v.y;
                    }
                  an += P;
                  gN += step;
                  gD += step;
                }
            }
        }
    }
}

constexpr size_t rows_smem_bytes()
{
  return sizeof(double) * (2 * (size_t)T1_W * T1_STRIDE + 2 * T1_CHUNK * T1_REC) + 40 + TILE_MAX_CELLS * 4 +
         ((T1_W + 3) / 4) * 16;
}

__global__ void __launch_bounds__(T1_THREADS, WBEM_T1_MINCTAS) k_assemble_colours(const TiledArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *acc = reinterpret_cast<double2 *>(smem_raw);                    // [W][T1_STRIDE] (N, D)
  double *geo = reinterpret_cast<double *>(acc + (size_t)T1_W * T1_STRIDE); // [2][CHUNK][8][16]
  uint64_t *bar = reinterpret_cast<uint64_t *>(geo + 2 * T1_CHUNK * T1_REC); // [2] full + [2] empty
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(bar + 4);
  uint32_t *s_slots = reinterpret_cast<uint32_t *>(bar + 5);
  uint32_t *s_col = s_slots + TILE_MAX_CELLS;

  const int tid = threadIdx.x;
  // one 16-byte load tells the CTA its work item: the first panel chunk is on its way one
  // global-memory latency after the CTA starts
  const uint4 dsc = *reinterpret_cast<const uint4 *>(a.desc + a.cluster_base + blockIdx.x);
  const uint32_t p0 = dsc.x, s0 = dsc.y, cluster = dsc.w;
  const int ncell = (int)(dsc.z & 0xffu), nslot = (int)(dsc.z >> 8);
  const uint32_t p1 = p0 + ncell;
  const int nchunk = (ncell + T1_CHUNK - 1) / T1_CHUNK;
  const uint32_t lrow_base = blockIdx.y * T1_ROWS;

  auto issue_chunk = [&](int c) {
    const int nc = min(T1_CHUNK, ncell - c * T1_CHUNK);
    const uint32_t bytes = (uint32_t)nc * T1_REC * sizeof(double);
    mbar_expect_tx(&bar[c & 1], bytes);
    bulk_copy_g2s(geo + (c & 1) * T1_CHUNK * T1_REC, a.geo + ((size_t)p0 + (size_t)c * T1_CHUNK) * T1_REC, bytes,
                  &bar[c & 1]);
  };
  if (tid == 0)
    {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_init(&bar[2], T1_THREADS);
      mbar_init(&bar[3], T1_THREADS);
      s_cnt[0] = 0;
      s_cnt[1] = 0;
      issue_chunk(0);
      if (nchunk > 1) issue_chunk(1);
    }
  // thread t owns the tile's rows t, t + T1_THREADS, ... (lanes stay on consecutive rows)
  unsigned long long smask[T1_RPT];
  double xi0[T1_RPT], xi1[T1_RPT], xi2[T1_RPT], row_sum[T1_RPT];
  const bool any_sing = a.tile_sing[(size_t)blockIdx.y * a.n_clusters + cluster] != 0;
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      const uint32_t lrow = lrow_base + r * T1_THREADS + tid;
      const uint32_t lrow_c = lrow < a.nloc ? lrow : a.nloc - 1;
      smask[r] = 0ull;
      if (any_sing)
        for (uint32_t k = a.sing_ptr[lrow_c]; k < a.sing_ptr[lrow_c + 1]; ++k)
          {
            const uint32_t pos = a.sing_cellpos[k];
            if (pos >= p0 && pos < p1) smask[r] |= 1ull << (pos - p0);
          }
      xi0[r] = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 0];
      xi1[r] = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 1];
      xi2[r] = a.xyz[3 * (size_t)(a.row0 + lrow_c) + 2];
      row_sum[r] = 0.0;
    }
  for (int i = tid; i < ncell; i += T1_THREADS) s_slots[i] = reinterpret_cast<const uint32_t *>(a.cell_slots)[p0 + i];
  for (int i = tid; i < nslot; i += T1_THREADS) s_col[i] = a.slot_col[s0 + i];
  double2 *accT = acc + tid;
  for (int s = 0; s < nslot; ++s)
#pragma unroll
    for (int r = 0; r < T1_RPT; ++r) accT[s * T1_STRIDE + r * T1_THREADS] = make_double2(0.0, 0.0);
  __syncthreads();

  const double u0 = a.g1_x[0], u1 = a.g1_x[1], u2 = a.g1_x[2], u3 = a.g1_x[3];
  for (int c = 0; c < nchunk; ++c)
    {
      mbar_wait(&bar[c & 1], (c >> 1) & 1);
      const double *gc = geo + (c & 1) * T1_CHUNK * T1_REC;
      const int kend = min(T1_CHUNK, ncell - c * T1_CHUNK);
      for (int kk = 0; kk < kend; ++kk)
        {
          const int k = c * T1_CHUNK + kk;
          const double *g = gc + kk * T1_REC;
          const uint32_t sl = s_slots[k];
          CellMoments cm;
          integrate_cell(g, xi0, xi1, xi2, u0, u1, u2, u3, cm);
          scatter_cell(accT, sl, cm, smask, k, row_sum);
        }
      if (c + 2 < nchunk)
        {
          mbar_arrive(&bar[2 + (c & 1)]);
          __syncwarp();
          if ((tid & 31) == 0)
            {
              if (atomicAdd(&s_cnt[c & 1], 1u) == T1_WARPS - 1)
                {
                  s_cnt[c & 1] = 0;
                  mbar_wait(&bar[2 + (c & 1)], (c >> 1) & 1);
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  issue_chunk(c + 2);
                }
            }
        }
    }
#pragma unroll
  for (int r = 0; r < T1_RPT; ++r)
    {
      const uint32_t lrow = lrow_base + r * T1_THREADS + tid;
      if (lrow < a.nloc) a.alpha_part[(size_t)cluster * a.nloc + lrow] = row_sum[r];
    }
  __syncwarp(); // a warp flushes exactly the rows its own lanes accumulated

  flush_rows(acc, s_col, nslot, lrow_base, a.nloc, a.ld, a.Nm, a.Dm, tid >> 5, tid & 31);
}

// ---------------------------------------------------------------------------------------
// Stream kernel: ONE launch for the whole assembly.  Persistent CTAs (3 per SM) draw (row tile, cluster)
// work items from a ticket counter.  Item order: groups of G consecutive 128-row tiles; inside a
// group colour after colour, inside a colour tile after tile, cluster after cluster.  So
//   * the STORE of a column, the ADDs of its other clusters and the neighbouring columns of the same
//     32-byte sector land within tens of microseconds: they meet in L2 and (almost) every sector
//     of the two matrices goes to HBM once.  The colour-per-launch order re-read and re-wrote the ADD
//     columns milliseconds later: 3.7 x the matrix bytes of DRAM traffic; here 1.19 x at G = 1,
//     1.65 x at G = 2 (N = 20 k; DESIGN.md has the table);
//   * there is no CTA launch / prologue bubble: the panel-chunk pipeline (TMA bulk copies,
//     full/empty mbarriers) runs on across work items, the next item's record arrives with its
//     first chunk, and the rows' coordinates of the next tile are fetched behind the last chunk.
// Order of the flushes: an item first stores the columns it is the first writer of, then waits until
// the lower-colour clusters it shares a column with have flushed the same row tile (done flags,
// release / acquire at gpu scope), then adds -- STORE before ADD and ADDs in colour order for any
// item order and CTA scheduling, so the matrices are bitwise reproducible.  G tiles per group put
// G x (clusters of a colour) tickets between an item and the ones it waits for; with fewer tickets
// than CTAs in flight (G = 1) some items do wait.  The next item's ticket is drawn as late as the
// chunk pipeline allows (NBUF + 2 chunks before the current item ends; thread 0, one chunk of
// integration between the dependent steps ticket -> descriptor -> hand-over, so it never waits for
// an answer): a CTA that falls behind holds back one drawn ticket for a fraction of an item, not
// two tickets for two items (which made lateness spread to every item waiting on them).
// Tickets are handed out in item order and an item only waits for lower tickets, which running
// CTAs own: no deadlock for any CTA scheduling; the wait is bounded anyway (error flag).
// No CTA-wide barrier in steady state: a warp accumulates and flushes its own 32 rows; the last
// warp to finish publishes the item's flag (the others order their stores at CTA scope only).
// ---------------------------------------------------------------------------------------
#define META_PRED 16  // longest predecessor list the record holds (else: colour-per-launch kernel)
#define META_TILES 12 // row tiles with singular pairs listed per cluster (more: "look it up")
#define STREAM_MAX_COLORS 16

struct ItemMeta
{ // per cluster, launch order; one bulk copy brings it into shared memory
  uint32_t p0;     // first cell (processing position)
  uint32_t counts; // cells | slots << 8 | predecessors << 16 | listed row tiles << 24 (0xff: not listed)
  uint32_t cluster;
  uint32_t pad;
  uint32_t pred[META_PRED];
  uint32_t sing_tiles[META_TILES];
  uint32_t cell_slots[TILE_MAX_CELLS];
  uint32_t slot_col[T1_W];
};
static_assert(sizeof(ItemMeta) % 16 == 0, "bulk copies move multiples of 16 bytes");

struct StreamArgs
{
  const double *xyz;
  const double *geo; // line records [C][GEO2_REC], processing order
  const ItemMeta *meta;                    // [clusters] launch order
  const uint2 *desc;                       // [clusters] launch order: first cell, cells
  const uint32_t *sing_ptr, *sing_cellpos; // CSR by local row
  double *Nm, *Dm, *alpha_part;
  unsigned int *ticket;     // zeroed before the launch
  unsigned int *done;       // [row tiles][clusters]: epoch of the last assembly that flushed the item
  unsigned int *error_flag; // set when a dependency wait gave up
  uint32_t epoch, n_items, n_clusters, ld, row0, nloc;
  uint32_t row_tiles, group_tiles, n_colors;
  uint32_t color_ptr[STREAM_MAX_COLORS + 1];
  double g1_x[4], g1_w[4], g1_wx[4]; // 1-D Gauss nodes, weights, weights x nodes
};

#ifndef WBEM_T1_NBUF
#define WBEM_T1_NBUF 2
#endif
#define T1_NBUF WBEM_T1_NBUF // panel-chunk buffers of the stream kernel (every item has at least this many chunks, possibly empty)

constexpr size_t stream_smem_bytes()
{
  return sizeof(double2) * (size_t)T1_W * T1_STRIDE + sizeof(double) * T1_NBUF * T1_CHUNK * GEO2_REC + 2 * sizeof(ItemMeta) +
         (2 * T1_NBUF + 2) * sizeof(uint64_t) + (T1_NBUF + 2 + (T1_NBUF & 1)) * sizeof(uint32_t) + 2 * sizeof(uint2) + 2 * sizeof(uint2);
}

__device__ __forceinline__ void st_relaxed_gpu(unsigned int *p, unsigned int v)
{
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// A warp (its lane 0) reports that its rows of an item are flushed; the last of the CTA's warps
// publishes the item's flag.  The others only order their stores at CTA scope (cheap); the gpu-scope
// fence of the last one is cumulative over what it has observed through the counter.
__device__ __forceinline__ void signal_flushed(unsigned int *flag, uint32_t *counter, const unsigned int epoch)
{
  asm volatile("fence.acq_rel.cta;" ::: "memory");
  if (atomicAdd(counter, 1u) == T1_WARPS - 1)
    {
      *counter = 0;
      fence_acq_rel_gpu();
      st_relaxed_gpu(flag, epoch);
    }
}

// ticket -> (row tile, launch position); tile = 0xffffffff past the end.  Order: groups of group_tiles
// row tiles; inside a group colour after colour, inside a colour tile after tile, cluster after cluster.
__host__ __device__ __forceinline__ uint2 decode_ticket_order(const uint32_t t, const uint32_t n_items, const uint32_t n_clusters,
                                                              const uint32_t row_tiles, const uint32_t group_tiles,
                                                              const uint32_t n_colors, const uint32_t *color_ptr)
{
  if (t >= n_items) return make_uint2(0xffffffffu, 0u);
  const uint32_t per_group = group_tiles * n_clusters;
  const uint32_t g = t / per_group;
  uint32_t r = t - g * per_group;
  const uint32_t left = row_tiles - g * group_tiles;
  const uint32_t G = group_tiles < left ? group_tiles : left;
  uint32_t c = 0;
  while (c + 1 < n_colors && r >= G * color_ptr[c + 1]) ++c;
  r -= G * color_ptr[c];
  const uint32_t nc = color_ptr[c + 1] - color_ptr[c];
  const uint32_t tg = r / nc;
  return make_uint2(g * group_tiles + tg, color_ptr[c] + (r - tg * nc));
}
__device__ __forceinline__ uint2 decode_ticket(const uint32_t t, const StreamArgs &a)
{
  return decode_ticket_order(t, a.n_items, a.n_clusters, a.row_tiles, a.group_tiles, a.n_colors, a.color_ptr);
}

__global__ void __launch_bounds__(T1_THREADS, WBEM_T1_MINCTAS) k_assemble_rows(const StreamArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2 *acc = reinterpret_cast<double2 *>(smem_raw);                              // [W][T1_STRIDE] (N, D)
  double *geo = reinterpret_cast<double *>(acc + (size_t)T1_W * T1_STRIDE);          // [NBUF][CHUNK][GEO2_REC]
  ItemMeta *meta = reinterpret_cast<ItemMeta *>(geo + T1_NBUF * T1_CHUNK * GEO2_REC); // [2] by item parity
  uint64_t *bar = reinterpret_cast<uint64_t *>(meta + 2);                            // [NBUF] full, [NBUF] empty, [2] next item known
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(bar + 2 * T1_NBUF + 2);             // [NBUF] refill counters
  uint32_t *s_fin = s_cnt + T1_NBUF;                                                 // [2] warps that reported an item flushed
  volatile uint2 *s_item = reinterpret_cast<volatile uint2 *>(s_fin + 2 + (T1_NBUF & 1)); // [2] by item parity: tile, launch position
  volatile uint2 *s_desc = s_item + 2;                                               // [2] by item parity: first cell, cells
  uint64_t *const bar_full = bar, *const bar_empty = bar + T1_NBUF, *const bar_next = bar + 2 * T1_NBUF;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t ncl = a.n_clusters;

  // stream chunk -> buffer / barrier; msrc: the item's record rides on its first chunk
  auto issue = [&](uint32_t p0, int ncell, int cidx, int buf, const ItemMeta *msrc, ItemMeta *mdst) {
    int nc = ncell - cidx * T1_CHUNK;
    nc = nc < 0 ? 0 : (nc > T1_CHUNK ? T1_CHUNK : nc);
    const uint32_t gbytes = (uint32_t)nc * GEO2_REC * sizeof(double);
    mbar_expect_tx(&bar_full[buf], gbytes + (msrc ? (uint32_t)sizeof(ItemMeta) : 0u));
    if (gbytes)
      bulk_copy_g2s(geo + buf * T1_CHUNK * GEO2_REC, a.geo + ((size_t)p0 + (size_t)cidx * T1_CHUNK) * GEO2_REC, gbytes,
                    &bar_full[buf]);
    if (msrc) bulk_copy_g2s(mdst, msrc, (uint32_t)sizeof(ItemMeta), &bar_full[buf]);
  };

  if (tid == 0)
    {
      for (int b = 0; b < T1_NBUF; ++b)
        {
          mbar_init(&bar_full[b], 1);
          mbar_init(&bar_empty[b], T1_THREADS);
          s_cnt[b] = 0;
        }
      mbar_init(&bar_next[0], 1);
      mbar_init(&bar_next[1], 1);
      mbar_arrive(&bar_next[0]); // item 0 is handed over right here: use number n of barrier b announces item 2 n + b
      s_fin[0] = s_fin[1] = 0;
      const uint2 i0 = decode_ticket(atomicAdd(a.ticket, 1u), a);
      s_item[0].x = i0.x;
      s_item[0].y = i0.y;
      if (i0.x != 0xffffffffu)
        {
          const uint2 d0 = a.desc[i0.y];
          s_desc[0].x = d0.x;
          s_desc[0].y = d0.y;
          for (int b = 0; b < T1_NBUF; ++b) issue(d0.x, (int)d0.y, b, b, b == 0 ? a.meta + i0.y : nullptr, meta);
        }
    }
  __syncthreads();

  const double un[4] = {a.g1_x[0], a.g1_x[1], a.g1_x[2], a.g1_x[3]};
  const double wq[4] = {a.g1_w[0], a.g1_w[1], a.g1_w[2], a.g1_w[3]};
  const double wuq[4] = {a.g1_wx[0], a.g1_wx[1], a.g1_wx[2], a.g1_wx[3]};
  double xi0[T1_RPT], xi1[T1_RPT], xi2[T1_RPT]; // this item's rows
  double xn0[T1_RPT], xn1[T1_RPT], xn2[T1_RPT]; // prefetched for the next row tile
  uint32_t cur_tile = 0xffffffffu;
  uint32_t gbuf = 0, gph = 0; // buffer and phase of the next chunk of this CTA's stream
  double2 *accT = acc + tid;

  auto load_rows = [&](uint32_t tile, double(&o0)[T1_RPT], double(&o1)[T1_RPT], double(&o2)[T1_RPT]) {
#pragma unroll
    for (int r = 0; r < T1_RPT; ++r)
      {
        const uint32_t lrow = tile * T1_ROWS + r * T1_THREADS + tid;
        const uint32_t lrow_c = lrow < a.nloc ? lrow : a.nloc - 1;
        const double *px = a.xyz + 3 * (size_t)(a.row0 + lrow_c);
        o0[r] = px[0];
        o1[r] = px[1];
        o2[r] = px[2];
      }
  };

  for (uint32_t it = 0;; ++it)
    {
      const uint32_t tile = s_item[it & 1].x, kpos = s_item[it & 1].y;
      if (tile == 0xffffffffu) break;
      if (tile != cur_tile)
        {
          if (it == 0)
            load_rows(tile, xi0, xi1, xi2);
          else
            {
#pragma unroll
              for (int r = 0; r < T1_RPT; ++r) xi0[r] = xn0[r], xi1[r] = xn1[r], xi2[r] = xn2[r];
            }
          cur_tile = tile;
        }
      // the item's record arrives with its first chunk
      mbar_wait(&bar_full[gbuf], gph);
      const ItemMeta *mt = meta + (it & 1);
      const uint32_t p0 = mt->p0, counts = mt->counts, cluster = mt->cluster;
      const int ncell = (int)(counts & 0xffu), nslot = (int)((counts >> 8) & 0xffu);
      const int npred = (int)((counts >> 16) & 0xffu), ntl = (int)(counts >> 24);
      // every item has at least NBUF + 2 chunks (possibly empty ones): the next item is drawn as late
      // as the pipeline allows -- NBUF + 2 chunks before this one ends
      const int nchunk = max(T1_NBUF + 2, (ncell + T1_CHUNK - 1) / T1_CHUNK);
      const int c_draw = nchunk - T1_NBUF - 2;
      const uint32_t lrow_base = tile * T1_ROWS;

      bool any_sing = ntl == 0xff;
      for (int i = 0; i < ntl && i < META_TILES; ++i) any_sing |= (mt->sing_tiles[i] == tile);
      unsigned long long smask[T1_RPT];
      double row_sum[T1_RPT];
#pragma unroll
      for (int r = 0; r < T1_RPT; ++r)
        {
          smask[r] = 0ull;
          row_sum[r] = 0.0;
          if (any_sing)
            {
              const uint32_t lrow = lrow_base + r * T1_THREADS + tid;
              const uint32_t lrow_c = lrow < a.nloc ? lrow : a.nloc - 1;
              for (uint32_t k = a.sing_ptr[lrow_c]; k < a.sing_ptr[lrow_c + 1]; ++k)
                {
                  const uint32_t pos = a.sing_cellpos[k];
                  if (pos >= p0 && pos < p0 + ncell) smask[r] |= 1ull << (pos - p0);
                }
            }
        }
      for (int s = 0; s < nslot; ++s)
#pragma unroll
        for (int r = 0; r < T1_RPT; ++r) accT[s * T1_STRIDE + r * T1_THREADS] = make_double2(0.0, 0.0);

      unsigned int *const done_tile = a.done + (size_t)tile * ncl;
      uint32_t tkn = 0xffffffffu;          // thread 0: the next item's ticket ...
      uint2 itn = make_uint2(0xffffffffu, 0u), dn = make_uint2(0u, 0u); // ... decoded, and its descriptor
      for (int c = 0; c < nchunk; ++c)
        {
          const int buf = gbuf;
          const uint32_t ph = gph;
          if (++gbuf == T1_NBUF) gbuf = 0, gph ^= 1u;
          mbar_wait(&bar_full[buf], ph);
          // thread 0 draws the next item: ticket, decode + descriptor, hand-over -- one chunk of
          // integration between the steps, so that it never waits for the answers
          if (tid == 0)
            {
              if (c == c_draw) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(tkn) : "l"(a.ticket) : "memory");
              if (c == c_draw + 1)
                {
                  itn = decode_ticket(tkn, a);
                  if (itn.x != 0xffffffffu)
                    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(dn.x), "=r"(dn.y) : "l"(a.desc + itn.y));
                }
              if (c == c_draw + 2)
                {
                  s_item[(it + 1) & 1].x = itn.x;
                  s_item[(it + 1) & 1].y = itn.y;
                  s_desc[(it + 1) & 1].x = dn.x;
                  s_desc[(it + 1) & 1].y = dn.y;
                  mbar_arrive(&bar_next[(it + 1) & 1]);
                }
            }
          if (c == nchunk - 1)
            { // everybody: the next item is known by now; its rows' coordinates are fetched behind this chunk
              mbar_wait(&bar_next[(it + 1) & 1], ((it + 1) >> 1) & 1);
              const uint32_t tile_next = s_item[(it + 1) & 1].x;
              if (tile_next != 0xffffffffu && tile_next != tile) load_rows(tile_next, xn0, xn1, xn2);
            }
          const double *gc = geo + buf * T1_CHUNK * GEO2_REC;
          const int kend = min(T1_CHUNK, ncell - c * T1_CHUNK);
          for (int kk = 0; kk < kend; ++kk)
            {
              const int k = c * T1_CHUNK + kk;
              const double *g = gc + kk * GEO2_REC;
              const uint32_t sl = mt->cell_slots[k];
              CellMoments cm;
#if !defined(WBEM_NO_LINE_PAIRS) && WBEM_T1_RPT == 1
              // (8 points in flight; all 16 at once, 156 registers, measured 1 % slower)
              line_group_moments<0, 2>(g, xi0, xi1, xi2, un, wq, wuq, cm);
              line_group_moments<2, 2>(g, xi0, xi1, xi2, un, wq, wuq, cm);
#else
              line_moments<0>(g, xi0, xi1, xi2, un, wq, wuq, cm);
              line_moments<1>(g, xi0, xi1, xi2, un, wq, wuq, cm);
              line_moments<2>(g, xi0, xi1, xi2, un, wq, wuq, cm);
              line_moments<3>(g, xi0, xi1, xi2, un, wq, wuq, cm);
#endif
              scatter_cell_distinct(accT, sl, cm, smask, k, row_sum);
            }
          // buffer consumed: the warp that arrives last refills it with stream chunk + NBUF
          mbar_arrive(&bar_empty[buf]);
          __syncwarp();
          if (lane == 0)
            {
              if (atomicAdd(&s_cnt[buf], 1u) == T1_WARPS - 1)
                {
                  s_cnt[buf] = 0;
                  mbar_wait(&bar_empty[buf], ph);
                  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                  const int target = c + T1_NBUF;
                  if (target < nchunk)
                    issue(p0, ncell, target, buf, nullptr, nullptr);
                  else
                    { // the next item's first chunks (thread 0 has handed it over before its own arrival)
                      const uint32_t kn = s_item[(it + 1) & 1].y;
                      if (s_item[(it + 1) & 1].x != 0xffffffffu)
                        issue(s_desc[(it + 1) & 1].x, (int)s_desc[(it + 1) & 1].y, target - nchunk, buf,
                              target == nchunk ? a.meta + kn : nullptr, meta + ((it + 1) & 1));
                    }
                }
            }
        }
      __syncwarp(); // a warp flushes exactly the rows its own lanes accumulated
      // first writers first (nobody to wait for), then the columns other clusters have written before
      int nstore = 0;
      for (int sb = 0; sb < nslot; sb += 32)
        nstore += __popc(__ballot_sync(0xffffffffu, sb + lane < nslot && (mt->slot_col[sb + lane] & WBEM_SLOT_ADD) == 0u));
#ifndef WBEM_DBG_NOFLUSH
      flush_pass<false>(acc, mt->slot_col, 0, nstore, lrow_base, a.nloc, a.ld, a.Nm, a.Dm, warp, lane);
#endif
      // the lower-colour clusters sharing a column must have flushed this row tile (acquire loads of their flags)
#ifndef WBEM_DBG_NOWAIT
      if (npred)
        {
          bool ok = true;
#ifdef WBEM_DBG_WAITLOG
          unsigned int dbg_spins = 0;
#endif
          if (lane < npred)
            {
              const unsigned int *f = done_tile + mt->pred[lane];
              uint32_t spins = 0;
              while (ld_acquire_gpu(f) != a.epoch)
                {
                  __nanosleep(200);
                  if (++spins > (1u << 22))
                    {
                      ok = false;
                      break;
                    }
                }
#ifdef WBEM_DBG_WAITLOG
              dbg_spins = spins;
#endif
            }
          if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicExch(a.error_flag, 1u);
#ifdef WBEM_DBG_WAITLOG
          {
            unsigned int sp = 0;
            if (lane < npred) sp = dbg_spins;
            for (int off = 16; off > 0; off >>= 1) sp = max(sp, __shfl_xor_sync(0xffffffffu, sp, off));
            if (lane == 0 && warp == 0) a.done[(size_t)a.row_tiles * ncl + (size_t)tile * ncl + kpos] = sp | (blockIdx.x << 20);
          }
#endif
        }
#endif
#ifndef WBEM_DBG_NOFLUSH
      flush_pass<true>(acc, mt->slot_col, nstore, nslot - nstore, lrow_base, a.nloc, a.ld, a.Nm, a.Dm, warp, lane);
#endif
#pragma unroll
      for (int r = 0; r < T1_RPT; ++r)
        {
          const uint32_t lrow = lrow_base + r * T1_THREADS + tid;
          if (lrow < a.nloc) a.alpha_part[(size_t)cluster * a.nloc + lrow] = row_sum[r];
        }
      __syncwarp();
      if (lane == 0) signal_flushed(done_tile + kpos, &s_fin[it & 1], a.epoch);
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
  extern __shared__ __align__(16) double sgeo[]; // [TC][GEO_REC][nq]
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
      for (int i = tid; i < nc * GEO_REC * nq; i += SIMPLE_ROWS) sgeo[i] = geo[(size_t)cb * GEO_REC * nq + i];
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
          const double *g = sgeo + k * GEO_REC * nq;
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

// Per-cluster records of the stream kernel (launch order) and its synchronisation words.
int wbem_upload_stream_tables(wbem_ctx *ctx, const std::vector<uint32_t> &sing_ptr, const std::vector<uint32_t> &sing_pos,
                              const std::vector<uint32_t> &cluster_of_pos)
{
  const AssemblyPlan &pl = ctx->plan;
  const uint32_t ncl = pl.n_clusters;
  const uint32_t row_tiles = (ctx->nloc + T1_ROWS - 1) / T1_ROWS;
  ctx->stream_ok = ncl > 0 && pl.max_pred <= META_PRED && pl.max_cells <= TILE_MAX_CELLS && pl.W <= T1_W &&
                   pl.n_colors <= STREAM_MAX_COLORS && (uint64_t)row_tiles * ncl < 0xfffffff0ull;
  if (!ctx->stream_ok) return 0;
  // row tiles in which a cluster has singular pairs
  std::vector<std::vector<uint32_t>> tiles_of(ncl);
  for (uint32_t r = 0; r < ctx->nloc; ++r)
    for (uint32_t k = sing_ptr[r]; k < sing_ptr[r + 1]; ++k)
      {
        std::vector<uint32_t> &t = tiles_of[cluster_of_pos[sing_pos[k]]];
        if (t.empty() || t.back() != r / T1_ROWS) t.push_back(r / T1_ROWS); // rows ascend
      }
  std::vector<ItemMeta> meta(ncl);
  std::vector<uint2> desc(ncl);
  for (uint32_t k = 0; k < ncl; ++k)
    {
      const uint32_t cl = pl.color_clusters[k];
      ItemMeta &m = meta[k];
      memset(&m, 0, sizeof(m));
      const uint32_t nc = pl.cl_cell_ptr[cl + 1] - pl.cl_cell_ptr[cl], nsl = pl.cl_slot_ptr[cl + 1] - pl.cl_slot_ptr[cl];
      const uint32_t npred = pl.pred_ptr[k + 1] - pl.pred_ptr[k];
      const std::vector<uint32_t> &tl = tiles_of[cl];
      const uint32_t ntl = tl.size() <= META_TILES ? (uint32_t)tl.size() : 0xffu;
      m.p0 = pl.cl_cell_ptr[cl];
      m.counts = nc | (nsl << 8) | (npred << 16) | (ntl << 24);
      m.cluster = cl;
      for (uint32_t i = 0; i < npred; ++i) m.pred[i] = pl.pred[pl.pred_ptr[k] + i];
      if (ntl != 0xffu)
        for (uint32_t i = 0; i < ntl; ++i) m.sing_tiles[i] = tl[i];
      for (uint32_t i = 0; i < nc; ++i)
        memcpy(&m.cell_slots[i], &pl.cell_slots[4 * (size_t)(m.p0 + i)], 4);
      for (uint32_t i = 0; i < nsl; ++i) m.slot_col[i] = pl.slot_col[pl.cl_slot_ptr[cl] + i];
      desc[k] = make_uint2(m.p0, nc);
    }
  if (ctx->d_item_meta) cudaFree(ctx->d_item_meta);
  if (ctx->d_item_desc) cudaFree(ctx->d_item_desc);
  if (ctx->d_asm_sync) cudaFree(ctx->d_asm_sync);
  ctx->d_item_meta = ctx->d_item_desc = nullptr;
  ctx->d_asm_sync = nullptr;
  const size_t nsync = 4 + 2 * (size_t)std::max(1u, row_tiles) * ncl; // (second half: wait log of debug builds)
  CUDA_OK(ctx, cudaMalloc(&ctx->d_item_meta, sizeof(ItemMeta) * ncl));
  CUDA_OK(ctx, cudaMalloc(&ctx->d_item_desc, sizeof(uint2) * ncl));
  CUDA_OK(ctx, cudaMalloc((void **)&ctx->d_asm_sync, sizeof(unsigned int) * nsync));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_item_meta, meta.data(), sizeof(ItemMeta) * ncl, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_item_desc, desc.data(), sizeof(uint2) * ncl, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(ctx, cudaMemsetAsync(ctx->d_asm_sync, 0, sizeof(unsigned int) * nsync, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream)); // the host vectors go out of scope
  ctx->asm_epoch = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
int wbem_launch_assemble(wbem_ctx *ctx)
{
  cudaStream_t st = ctx->stream;
  const int nq = ctx->qt.nq;
  if (ctx->nloc == 0 || ctx->C == 0) return 0;
  const bool tiled = (ctx->p.assemble_variant == 0 || ctx->p.assemble_variant == 2) && (ctx->qt.n1 == 4) && !ctx->has_degenerate_cells;
  // the line records need the cells' vertices: caller-supplied FEValues (any mapping) take the per-point kernel
  const bool stream = tiled && ctx->p.assemble_variant == 0 && ctx->stream_ok && !ctx->fevalues_given && ctx->d_cellgeo2;
  CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], st));
  if (tiled)
    {
      const AssemblyPlan &pl = ctx->plan;
      if (!ctx->tiled_attr_set)
        { // per device: a second context on another GPU needs its own opt-in
          CUDA_OK(ctx, cudaFuncSetAttribute(k_assemble_colours, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)rows_smem_bytes()));
          CUDA_OK(ctx, cudaFuncSetAttribute(k_assemble_rows, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)stream_smem_bytes()));
          CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->stream_ctas_per_sm, k_assemble_rows, T1_THREADS,
                                                                     stream_smem_bytes()));
          if (ctx->n_sm == 0) CUDA_OK(ctx, cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, ctx->dev));
          if (const char *e = getenv("WBEM_ASM_GROUP")) ctx->asm_group_tiles = atoi(e); // tuning experiments
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
    }
  if (stream && ctx->stream_ctas_per_sm > 0)
    {
      const AssemblyPlan &pl = ctx->plan;
      const uint32_t row_tiles = (ctx->nloc + T1_ROWS - 1) / T1_ROWS;
      StreamArgs a;
      a.xyz = ctx->d_xyz;
      a.geo = ctx->d_cellgeo2;
      a.meta = (const ItemMeta *)ctx->d_item_meta;
      a.desc = (const uint2 *)ctx->d_item_desc;
      a.sing_ptr = ctx->d_sing_ptr;
      a.sing_cellpos = ctx->d_sing_cellpos;
      a.Nm = ctx->d_Nm;
      a.Dm = ctx->d_Dm;
      a.alpha_part = ctx->d_alpha_part;
      a.ticket = ctx->d_asm_sync;
      a.error_flag = ctx->d_asm_sync + 1;
      a.done = ctx->d_asm_sync + 4;
      if (++ctx->asm_epoch == 0)
        { // the flags hold epochs: start over after a wrap
          CUDA_OK(ctx, cudaMemsetAsync(ctx->d_asm_sync, 0, sizeof(unsigned int) * (4 + (size_t)row_tiles * pl.n_clusters), st));
          ctx->asm_epoch = 1;
        }
      a.epoch = ctx->asm_epoch;
      a.n_items = row_tiles * pl.n_clusters;
      a.n_clusters = pl.n_clusters;
      a.ld = ctx->ld;
      a.row0 = ctx->row0;
      a.nloc = ctx->nloc;
      a.row_tiles = row_tiles;
      a.group_tiles = std::max(1, std::min(ctx->asm_group_tiles, (int)row_tiles));
      a.n_colors = pl.n_colors;
      for (uint32_t c = 0; c <= pl.n_colors; ++c) a.color_ptr[c] = pl.color_ptr[c];
      for (int i = 0; i < 4; ++i)
        {
          a.g1_x[i] = ctx->qt.g1_x[i];
          a.g1_w[i] = ctx->qt.g1_w[i];
          a.g1_wx[i] = ctx->qt.g1_w[i] * ctx->qt.g1_x[i];
        }
      CUDA_OK(ctx, cudaMemsetAsync(ctx->d_asm_sync, 0, sizeof(unsigned int), st)); // ticket
      const uint32_t want = (uint32_t)(ctx->n_sm * ctx->stream_ctas_per_sm);
      const uint32_t grid = std::min(want, a.n_items);
      k_assemble_rows<<<std::max(1u, grid), T1_THREADS, stream_smem_bytes(), st>>>(a);
      ctx->launches++;
      CUDA_OK(ctx, cudaGetLastError());
      if (!ctx->h_asm_flag) CUDA_OK(ctx, cudaMallocHost((void **)&ctx->h_asm_flag, sizeof(unsigned int)));
      CUDA_OK(ctx, cudaMemcpyAsync(ctx->h_asm_flag, ctx->d_asm_sync + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    }
  else if (tiled)
    {
      const AssemblyPlan &pl = ctx->plan;
      TiledArgs a;
      a.xyz = ctx->d_xyz;
      a.geo = ctx->d_cellgeo;
      a.cell_slots = ctx->d_cell_slots;
      a.desc = (const CtaDesc *)ctx->d_cta_desc;
      a.slot_col = ctx->d_slot_col;
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
      const uint32_t tile_rows = T1_ROWS;
      const uint32_t row_tiles = (ctx->nloc + tile_rows - 1) / tile_rows;
      for (uint32_t c = 0; c < pl.n_colors; ++c)
        {
          const uint32_t nclu = pl.color_ptr[c + 1] - pl.color_ptr[c];
          if (nclu == 0) continue;
          a.cluster_base = pl.color_ptr[c];
          dim3 grid(nclu, row_tiles);
          k_assemble_colours<<<grid, T1_THREADS, rows_smem_bytes(), st>>>(a);
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
      const size_t sm = sizeof(double) * SIMPLE_TC * GEO_REC * nq;
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

// Host-only check of the stream kernel's item order (no GPU; tests/test_plan_and_host.py): every (row tile,
// cluster) gets exactly one ticket and every cluster's predecessors on the same tile hold lower tickets.
// stats[0..3] = items, smallest / mean ticket distance between an item and a predecessor, row tiles.
extern "C" int wbem_stream_order_check(uint32_t n_dofs, uint32_t n_cells, const uint32_t *cell_dofs, uint32_t n_rows,
                                       uint32_t group_tiles, double *stats4)
{
  AssemblyPlan pl;
  if (wbem_build_plan(n_dofs, n_cells, cell_dofs, T1_W, TILE_MAX_CELLS, &pl)) return 100;
  if (pl.n_colors > STREAM_MAX_COLORS || pl.max_pred > META_PRED) return 101; // the colour kernel's case
  const uint32_t ncl = pl.n_clusters, row_tiles = (n_rows + T1_ROWS - 1) / T1_ROWS;
  const uint32_t n_items = row_tiles * ncl;
  group_tiles = std::max(1u, std::min(group_tiles, row_tiles));
  std::vector<uint32_t> ticket_of((size_t)n_items, 0xffffffffu);
  for (uint32_t t = 0; t < n_items; ++t)
    {
      const uint2 it = decode_ticket_order(t, n_items, ncl, row_tiles, group_tiles, pl.n_colors, pl.color_ptr.data());
      if (it.x >= row_tiles || it.y >= ncl) return 1;
      uint32_t &slot = ticket_of[(size_t)it.x * ncl + it.y];
      if (slot != 0xffffffffu) return 2; // drawn twice
      slot = t;
    }
  if (decode_ticket_order(n_items, n_items, ncl, row_tiles, group_tiles, pl.n_colors, pl.color_ptr.data()).x != 0xffffffffu) return 3;
  double dmin = 1e300, dsum = 0, dn = 0;
  for (uint32_t tile = 0; tile < row_tiles; ++tile)
    for (uint32_t k = 0; k < ncl; ++k)
      for (uint32_t q = pl.pred_ptr[k]; q < pl.pred_ptr[k + 1]; ++q)
        {
          const uint32_t a = ticket_of[(size_t)tile * ncl + k], b = ticket_of[(size_t)tile * ncl + pl.pred[q]];
          if (b >= a) return 4; // an item would wait for a ticket nobody has drawn yet
          dmin = std::min(dmin, (double)(a - b));
          dsum += a - b;
          dn += 1;
        }
  if (stats4)
    {
      stats4[0] = n_items;
      stats4[1] = dn ? dmin : 0;
      stats4[2] = dn ? dsum / dn : 0;
      stats4[3] = row_tiles;
    }
  return 0;
}

// development: raw copy of the stream kernel's synchronisation words (ticket, error flag, done flags, wait log)
extern "C" int wbem_debug_read_asm_sync(wbem_ctx *ctx, size_t offset, size_t n, unsigned int *out)
{
  if (!ctx || !ctx->d_asm_sync) return -1;
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_OK(ctx, cudaMemcpy(out, ctx->d_asm_sync + offset, n * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  return 0;
}

// tile geometry of the regular-pair kernel, for the host-side maps built in wbem_set_topology
uint32_t wbem_tile_rows(void) { return T1_ROWS; }
uint32_t wbem_tile_width(void) { return T1_W; }
uint32_t wbem_line_record_doubles(void) { return GEO2_REC; }
