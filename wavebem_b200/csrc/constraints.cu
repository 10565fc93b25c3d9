// constraints.cu -- BEMProblem<3>::compute_constraints (reference source/bem_problem.cc:990-1105)
// and the two L2 projections it runs on EVERY solve_system call:
//
//   ComputationalDomain::compute_normals      (source/computational_domain.cc:1525-1620)
//       M n_d = int phi_i n_d dS ,  nodes_normals = n / |n|
//   BEMProblem::compute_surface_gradients     (source/bem_problem.cc:1153-1293)
//       M g_d = int phi_i (grad_s phi)_d dS ,  phi = tmp_rhs o surface_nodes
//
// with M the Q1 mass matrix of the surface mesh (one scalar matrix shared by the three
// components of the reference's FESystem(FE_Q(1),3); dofs of different patches never couple
// because patch edges carry double nodes).  The reference factorises the 3N x 3N matrices with
// UMFPACK twice per solve_system; here
//   * the mass matrix and the right-hand sides are gathered node by node from the adjacent cells
//     (fixed order: bitwise reproducible, no atomics),
//   * the three systems are solved together by a Jacobi-preconditioned conjugate gradient that
//     runs inside ONE cooperative kernel (grid barriers, fixed-order reductions) to a relative
//     residual of 1e-15 -- the mass matrix is well conditioned (diagonally scaled spectrum inside
//     about [1/4, 9/4]), so this takes ~40 iterations and is the same answer as the sparse LU to
//     rounding,
//   * the constraint lines are then produced on the host (an O(N) walk over the double-node
//     sets, reference :1002-1101) and installed like wbem_set_constraints would.
// Normals are cached per geometry; they and the gradients are only computed when some double-node
// set holds two Dirichlet dofs (the only place the reference consumes them, :1043-1075).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "internal.h"
#include "q1map.cuh"

#define CON_NONE 0xffffffffu
#define CG_THREADS 256
#define CG_MAX_ITERS 300
#define CG_RTOL 1e-15

struct GaussTable
{
  int nq, pad;
  double u[WBEM_MAX_NQ], v[WBEM_MAX_NQ], w[WBEM_MAX_NQ];
};

struct ConState
{
  uint32_t N = 0, C = 0, MW = 0;
  uint32_t *d_cells = nullptr; // [C][4] caller's cell order
  uint8_t *d_dir = nullptr;    // [C]
  uint32_t *d_nc_ptr = nullptr, *d_nc_cell = nullptr; // node -> adjacent cells (ascending cell index)
  uint8_t *d_nc_local = nullptr;                      // local index of the node in that cell
  uint8_t *d_nc_pos = nullptr;                        // [adjacency][4]: ELL slot of the cell's dof j
  uint32_t *d_mcol = nullptr;                         // [N][MW] columns of the mass matrix rows
  double *d_mval = nullptr;                           // [N][MW]
  double *d_diag = nullptr;                           // [N]
  double *d_b = nullptr, *d_x = nullptr, *d_r = nullptr, *d_p = nullptr, *d_ap = nullptr; // [3][N]
  double *d_part = nullptr;                           // [2][grid][3] reduction partials
  unsigned int *d_barrier = nullptr;
  int *d_iters = nullptr;
  double *d_normals = nullptr, *d_grads = nullptr;    // [N][3]
  double *d_phi = nullptr;                            // [N] tmp_rhs o surface_nodes
  int grid = 0;
  bool ready = false;
  uint64_t mass_geom_version = ~0ull, normals_geom_version = ~0ull;
  std::vector<double> h_normals, h_grads;
  // hanging-node lines supplied by the caller (DoFTools::make_hanging_node_constraints, :1000)
  std::vector<uint32_t> base_lines, base_ptr{0}, base_col;
  std::vector<double> base_val;
  // the lines produced by the last wbem_compute_constraints
  std::vector<uint32_t> out_lines, out_ptr{0}, out_col;
  std::vector<double> out_val, out_inhom;
  int last_cg_iters = 0;
  // host walk: only the dofs whose double-node set has more than one member matter (a few % of N)
  std::vector<uint32_t> multi;
  std::vector<uint8_t> kind;       // per dof: 0 none, 1 one entry (master, 1.0), 2 inhomogeneity only
  std::vector<uint32_t> master;
  std::vector<double> inhom;
  std::vector<double> h_rhs;       // pinned-size staging of tmp_rhs
  std::vector<int32_t> line_of_tmp; // chain resolution scratch
  // hanging-node lines inside the two projections (the reference condenses
  // DoFTools::make_hanging_node_constraints(vector_dh) into both mass systems,
  // source/computational_domain.cc:1535-1538, source/bem_problem.cc:1167-1170)
  uint32_t nh = 0, nm = 0;
  uint32_t *d_h_dof = nullptr, *d_hm_ptr = nullptr, *d_hm_col = nullptr; // hanging dof -> masters
  double *d_hm_w = nullptr;
  uint32_t *d_m_dof = nullptr, *d_mh_ptr = nullptr, *d_mh_h = nullptr;    // master -> hanging dofs
  double *d_mh_w = nullptr;
  uint8_t *d_hflag = nullptr;                                             // [N] 1 on hanging dofs
  bool hang_uploaded = false;
};

static ConState *con_state(wbem_ctx *ctx)
{
  if (!ctx->con) ctx->con = new ConState();
  return reinterpret_cast<ConState *>(ctx->con);
}

void wbem_constraints_free(wbem_ctx *ctx)
{
  ConState *s = reinterpret_cast<ConState *>(ctx->con);
  if (!s) return;
  void *ptrs[] = {s->d_cells, s->d_dir,  s->d_nc_ptr, s->d_nc_cell, s->d_nc_local, s->d_nc_pos, s->d_mcol,
                  s->d_mval,  s->d_diag, s->d_b,      s->d_x,       s->d_r,        s->d_p,      s->d_ap,
                  s->d_part,  s->d_barrier, s->d_iters, s->d_normals, s->d_grads,  s->d_phi,
                  s->d_h_dof, s->d_hm_ptr, s->d_hm_col, s->d_hm_w, s->d_m_dof, s->d_mh_ptr, s->d_mh_h, s->d_mh_w,
                  s->d_hflag};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete s;
  ctx->con = nullptr;
}

int wbem_constraints_upload_tables(wbem_ctx *ctx)
{
  GaussTable t;
  t.nq = ctx->qt.nq;
  t.pad = 0;
  memcpy(t.u, ctx->qt.g_u, sizeof(t.u));
  memcpy(t.v, ctx->qt.g_v, sizeof(t.v));
  memcpy(t.w, ctx->qt.g_w, sizeof(t.w));
  // global memory, one copy per context (contexts with different orders may share a device)
  if (!ctx->d_gauss) CUDA_OK(ctx, cudaMalloc(&ctx->d_gauss, sizeof(GaussTable)));
  CUDA_OK(ctx, cudaMemcpy(ctx->d_gauss, &t, sizeof(t), cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// One thread per node i: row i of the mass matrix and the normal right-hand side, gathered from
// the adjacent cells in ascending cell order (the reference's cell loop, :1566-1600: local
// matrix summed over q, then added to the global entries).
__global__ void __launch_bounds__(128)
  k_mass_rows(const GaussTable *__restrict__ gq, uint32_t N, uint32_t MW, const double *__restrict__ xyz, const uint32_t *__restrict__ cells,
              const uint8_t *__restrict__ dir, const uint32_t *__restrict__ nc_ptr,
              const uint32_t *__restrict__ nc_cell, const uint8_t *__restrict__ nc_local,
              const uint8_t *__restrict__ nc_pos, double *__restrict__ mval, double *__restrict__ diag,
              double *__restrict__ b /* [3][N] */)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double *row = mval + (size_t)i * MW;
  for (uint32_t k = 0; k < MW; ++k) row[k] = 0.0;
  double bn[3] = {0, 0, 0}, dg = 0.0;
  const int nq = gq->nq;
  for (uint32_t a = nc_ptr[i]; a < nc_ptr[i + 1]; ++a)
    {
      const uint32_t c = nc_cell[a];
      const int li = nc_local[a];
      QuadVerts X;
      load_verts(xyz, cells + 4 * (size_t)c, X);
      const double sgn = dir[c] ? 1.0 : -1.0;
      double m[4] = {0, 0, 0, 0}, r[3] = {0, 0, 0};
      for (int q = 0; q < nq; ++q)
        {
          double y[3], cr[3], phi[4];
          map_q1(X, gq->u[q], gq->v[q], y, cr, phi);
          const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
          const double jxw = cn * gq->w[q];
#pragma unroll
          for (int j = 0; j < 4; ++j) m[j] += phi[li] * phi[j] * jxw;
          // normal = sgn (d_u x d_v)/|.| (cell->direction_flag()), times JxW
#pragma unroll
          for (int d = 0; d < 3; ++d) r[d] += phi[li] * (sgn * cr[d] / cn) * jxw;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) row[nc_pos[4 * (size_t)a + j]] += m[j];
      dg += m[li];
#pragma unroll
      for (int d = 0; d < 3; ++d) bn[d] += r[d];
    }
  if (nc_ptr[i] == nc_ptr[i + 1])
    { // a dof no cell touches: unit row, so that the solve leaves it at zero
      dg = 1.0;
      row[0] = 1.0;
    }
  diag[i] = dg;
#pragma unroll
  for (int d = 0; d < 3; ++d) b[(size_t)d * N + i] = bn[d];
}

// right-hand side of the surface-gradient projection (:1239-1262):
//   b_d[i] = sum_cells sum_q phi_i(q) (grad_s phi)_d(q) JxW(q)
// grad_s phi = [t_u t_v] G^-1 [d_u phi, d_v phi]^T with G the first fundamental form (the
// covariant transformation deal.II applies to the reference gradients in codimension one).
__global__ void __launch_bounds__(128)
  k_gradient_rhs(const GaussTable *__restrict__ gq, uint32_t N, const double *__restrict__ xyz, const uint32_t *__restrict__ cells,
                 const uint32_t *__restrict__ nc_ptr, const uint32_t *__restrict__ nc_cell,
                 const uint8_t *__restrict__ nc_local, const double *__restrict__ phi_nodes,
                 double *__restrict__ b /* [3][N] */)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double bg[3] = {0, 0, 0};
  const int nq = gq->nq;
  for (uint32_t a = nc_ptr[i]; a < nc_ptr[i + 1]; ++a)
    {
      const uint32_t c = nc_cell[a];
      const int li = nc_local[a];
      const uint32_t *dofs = cells + 4 * (size_t)c;
      QuadVerts X;
      load_verts(xyz, dofs, X);
      const double f0 = phi_nodes[dofs[0]], f1 = phi_nodes[dofs[1]], f2 = phi_nodes[dofs[2]], f3 = phi_nodes[dofs[3]];
      double r[3] = {0, 0, 0};
      for (int q = 0; q < nq; ++q)
        {
          const double u = gq->u[q], v = gq->v[q];
          double tu[3], tv[3];
          q1_tangents(X, u, v, tu, tv);
          const double E = tu[0] * tu[0] + tu[1] * tu[1] + tu[2] * tu[2];
          const double F = tu[0] * tv[0] + tu[1] * tv[1] + tu[2] * tv[2];
          const double G = tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2];
          const double det = E * G - F * F;
          const double du = (1 - v) * (f1 - f0) + v * (f3 - f2);
          const double dv = (1 - u) * (f2 - f0) + u * (f3 - f1);
          const double ca = (G * du - F * dv) / det, cb = (E * dv - F * du) / det;
          const double jxw = sqrt(det) * gq->w[q];
          const double ph = li == 0 ? (1 - u) * (1 - v) : li == 1 ? u * (1 - v) : li == 2 ? (1 - u) * v : u * v;
#pragma unroll
          for (int d = 0; d < 3; ++d) r[d] += ph * (tu[d] * ca + tv[d] * cb) * jxw;
        }
#pragma unroll
      for (int d = 0; d < 3; ++d) bg[d] += r[d];
    }
#pragma unroll
  for (int d = 0; d < 3; ++d) b[(size_t)d * N + i] = bg[d];
}

__global__ void k_mask_product(uint32_t N, const double *__restrict__ a, const double *__restrict__ m, double *__restrict__ out)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = a[i] * m[i];
}

__device__ __forceinline__ void con_grid_barrier(unsigned int *counter, unsigned int &epoch)
{
  __syncthreads();
  if (threadIdx.x == 0)
    {
      ++epoch;
      // release (the CTA's writes, ordered before this thread by the barrier above) / acquire: no further fences
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      const unsigned int target = epoch * gridDim.x;
      unsigned int v;
      do
        {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        }
      while (v < target);
    }
  __syncthreads();
}

// CTA-wide sums of three values, fixed order; result in out[0..2] of thread 0 only
__device__ __forceinline__ void cta_sum3(double v[3], double (*red)[3])
{
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[d] += __shfl_xor_sync(0xffffffffu, v[d], off);
  if ((threadIdx.x & 31) == 0)
    for (int d = 0; d < 3; ++d) red[threadIdx.x >> 5][d] = v[d];
  __syncthreads();
  if (threadIdx.x == 0)
    for (int d = 0; d < 3; ++d)
      {
        double t = 0;
        for (int w = 0; w < CG_THREADS / 32; ++w) t += red[w][d];
        v[d] = t;
      }
  __syncthreads();
}

struct HangArgs
{ // hanging-node lines x_h = sum_k w_k x_m(k) and their transpose (nh = 0: none)
  uint32_t nh, nm;
  const uint32_t *h_dof, *hm_ptr, *hm_col;
  const double *hm_w;
  const uint32_t *m_dof, *mh_ptr, *mh_h;
  const double *mh_w;
  const uint8_t *hflag;
};

// b_m += sum_h w_hm b_h, b_h = 0 (the condensed right-hand side C^T b); with_diag: the same fold of the
// Jacobi diagonal (w^2 weights; the preconditioner need not be exact) and a unit diagonal on hanging dofs
__global__ void k_con_fold(uint32_t N, HangArgs H, double *__restrict__ b, double *__restrict__ diag, int with_diag)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= H.nm) return;
  const uint32_t m = H.m_dof[k];
  double a0 = 0, a1 = 0, a2 = 0, dg = 0;
  for (uint32_t e = H.mh_ptr[k]; e < H.mh_ptr[k + 1]; ++e)
    {
      const uint32_t h = H.mh_h[e];
      const double w = H.mh_w[e];
      a0 += w * b[h];
      a1 += w * b[(size_t)N + h];
      a2 += w * b[2 * (size_t)N + h];
      dg += w * w * diag[h];
    }
  b[m] += a0;
  b[(size_t)N + m] += a1;
  b[2 * (size_t)N + m] += a2;
  if (with_diag) diag[m] += dg;
}
__global__ void k_con_zero_hanging(uint32_t N, HangArgs H, double *__restrict__ b, double *__restrict__ diag, int with_diag)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= H.nh) return;
  const uint32_t h = H.h_dof[k];
  b[h] = b[(size_t)N + h] = b[2 * (size_t)N + h] = 0.0;
  if (with_diag) diag[h] = 1.0;
}

// Jacobi-preconditioned CG on M x_d = b_d, d = 0..2, in lock step (cooperative launch).  With hanging
// nodes the operator is the condensed C^T M C: p is expanded to the hanging dofs before the product and
// the product's hanging rows are folded into their masters after it; x then holds C x_free, i.e. the
// distributed solution (ConstraintMatrix::distribute), without a separate pass.
__global__ void __launch_bounds__(CG_THREADS, 1)
  k_mass_cg(uint32_t N, uint32_t MW, const uint32_t *__restrict__ mcol, const double *__restrict__ mval,
            const double *__restrict__ diag, const double *__restrict__ b, double *__restrict__ x,
            double *__restrict__ r, double *p, double *ap, double *part,
            unsigned int *barrier, int *iters_out, double rtol, int max_iters, const HangArgs H)
{
  __shared__ double red[CG_THREADS / 32][3];
  __shared__ double s_tot[3];
  unsigned int epoch = 0;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  double *part0 = part, *part1 = part + (size_t)gridDim.x * 3;

  auto total = [&](const double *pp, double out[3]) { // same order in every CTA: identical results
    if (threadIdx.x < 96)
      { // one warp per component: lanes stride over the CTAs' partials (L2), then a fixed shuffle tree
        const uint32_t d = threadIdx.x >> 5, lane = threadIdx.x & 31;
        double t = 0;
        for (uint32_t c = lane; c < gridDim.x; c += 32) t += __ldcg(pp + (size_t)c * 3 + d);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (lane == 0) s_tot[d] = t;
      }
    __syncthreads();
    for (int d = 0; d < 3; ++d) out[d] = s_tot[d];
    __syncthreads();
  };

  // x = 0, r = b, z = r/diag, p = z, rz = r.z
  double loc[3] = {0, 0, 0};
  for (uint32_t i = t0; i < N; i += nthreads)
    {
      const double di = 1.0 / diag[i];
      for (int d = 0; d < 3; ++d)
        {
          const double bi = b[(size_t)d * N + i];
          x[(size_t)d * N + i] = 0.0;
          r[(size_t)d * N + i] = bi;
          p[(size_t)d * N + i] = bi * di;
          loc[d] += bi * bi * di;
        }
    }
  cta_sum3(loc, red);
  if (threadIdx.x == 0)
    for (int d = 0; d < 3; ++d) part0[(size_t)blockIdx.x * 3 + d] = loc[d];
  con_grid_barrier(barrier, epoch);
  double rz[3], rz0[3];
  total(part0, rz);
  for (int d = 0; d < 3; ++d) rz0[d] = rz[d];
  int it = 0;
  for (; it < max_iters; ++it)
    {
      bool done = true;
      for (int d = 0; d < 3; ++d) done &= !(rz[d] > rtol * rtol * rz0[d]);
      if (done) break; // uniform: every CTA holds the same rz
      if (H.nh)
        { // expand: p_h = sum w p_m
          for (uint32_t k = t0; k < H.nh; k += nthreads)
            {
              double e[3] = {0, 0, 0};
              for (uint32_t q = H.hm_ptr[k]; q < H.hm_ptr[k + 1]; ++q)
                for (int d = 0; d < 3; ++d) e[d] = fma(H.hm_w[q], __ldcg(p + (size_t)d * N + H.hm_col[q]), e[d]);
              for (int d = 0; d < 3; ++d) p[(size_t)d * N + H.h_dof[k]] = e[d];
            }
          con_grid_barrier(barrier, epoch);
        }
      // ap = M p, pAp
      double l2[3] = {0, 0, 0};
      for (uint32_t i = t0; i < N; i += nthreads)
        {
          double s[3] = {0, 0, 0};
          for (uint32_t k = 0; k < MW; ++k)
            {
              const uint32_t c = mcol[(size_t)i * MW + k];
              if (c == CON_NONE) break;
              const double m = mval[(size_t)i * MW + k];
              for (int d = 0; d < 3; ++d) s[d] = fma(m, __ldcg(p + (size_t)d * N + c), s[d]); // neighbours' p: L2
            }
          for (int d = 0; d < 3; ++d)
            {
              ap[(size_t)d * N + i] = s[d];
              l2[d] = fma(s[d], p[(size_t)d * N + i], l2[d]);
            }
        }
      if (H.nh)
        { // fold the hanging rows into their masters, then p.(C^T M C p) over the free dofs
          con_grid_barrier(barrier, epoch);
          for (uint32_t k = t0; k < H.nm; k += nthreads)
            {
              double e[3] = {0, 0, 0};
              for (uint32_t q = H.mh_ptr[k]; q < H.mh_ptr[k + 1]; ++q)
                for (int d = 0; d < 3; ++d) e[d] = fma(H.mh_w[q], __ldcg(ap + (size_t)d * N + H.mh_h[q]), e[d]);
              for (int d = 0; d < 3; ++d) ap[(size_t)d * N + H.m_dof[k]] += e[d];
            }
          con_grid_barrier(barrier, epoch);
          for (int d = 0; d < 3; ++d) l2[d] = 0.0;
          for (uint32_t i = t0; i < N; i += nthreads)
            if (!H.hflag[i])
              for (int d = 0; d < 3; ++d) l2[d] = fma(__ldcg(ap + (size_t)d * N + i), p[(size_t)d * N + i], l2[d]);
        }
      cta_sum3(l2, red);
      if (threadIdx.x == 0)
        for (int d = 0; d < 3; ++d) part1[(size_t)blockIdx.x * 3 + d] = l2[d];
      con_grid_barrier(barrier, epoch);
      double pap[3], alpha[3];
      total(part1, pap);
      for (int d = 0; d < 3; ++d) alpha[d] = (pap[d] > 0.0 && rz[d] > rtol * rtol * rz0[d]) ? rz[d] / pap[d] : 0.0;
      // x += alpha p, r -= alpha ap, new r.z
      double l3[3] = {0, 0, 0};
      for (uint32_t i = t0; i < N; i += nthreads)
        {
          const double di = 1.0 / diag[i];
          for (int d = 0; d < 3; ++d)
            {
              const size_t o = (size_t)d * N + i;
              x[o] = fma(alpha[d], p[o], x[o]);
              const double rn = (H.nh && H.hflag[i]) ? 0.0 : fma(-alpha[d], __ldcg(ap + o), r[o]);
              r[o] = rn;
              l3[d] = fma(rn * di, rn, l3[d]);
            }
        }
      cta_sum3(l3, red);
      if (threadIdx.x == 0)
        for (int d = 0; d < 3; ++d) part0[(size_t)blockIdx.x * 3 + d] = l3[d];
      con_grid_barrier(barrier, epoch);
      double rzn[3];
      total(part0, rzn);
      // p = z + beta p (frozen components keep alpha = 0 from now on)
      for (uint32_t i = t0; i < N; i += nthreads)
        {
          const double di = 1.0 / diag[i];
          for (int d = 0; d < 3; ++d)
            {
              const size_t o = (size_t)d * N + i;
              const double beta = rz[d] > 0.0 ? rzn[d] / rz[d] : 0.0;
              if (alpha[d] != 0.0) p[o] = fma(beta, p[o], r[o] * di);
            }
        }
      for (int d = 0; d < 3; ++d)
        if (alpha[d] != 0.0) rz[d] = rzn[d];
      con_grid_barrier(barrier, epoch); // p visible before the next product
    }
  if (t0 == 0) *iters_out = it;
}

__global__ void k_interleave3(uint32_t N, const double *__restrict__ x, double *__restrict__ out, int normalise)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double a = x[i], b = x[(size_t)N + i], c = x[2 * (size_t)N + i];
  if (normalise)
    { // nodes_normals[i] /= |nodes_normals[i]| (computational_domain.cc:1613)
      const double n = sqrt(a * a + b * b + c * c);
      a /= n;
      b /= n;
      c /= n;
    }
  out[3 * (size_t)i] = a;
  out[3 * (size_t)i + 1] = b;
  out[3 * (size_t)i + 2] = c;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <typename T>
static int con_upload(wbem_ctx *ctx, T **p, const std::vector<T> &v)
{
  if (*p) cudaFree(*p);
  *p = nullptr;
  CUDA_OK(ctx, cudaMalloc((void **)p, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty())
    CUDA_OK(ctx, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

static int con_prepare(wbem_ctx *ctx)
{
  ConState *s = con_state(ctx);
  if (s->ready) return 0;
  if (!ctx->N) WBEM_FAIL(ctx, -3, "constraints before wbem_set_topology");
  const uint32_t N = ctx->N, C = ctx->C;
  const uint32_t *cd = ctx->h_cell_dofs.data();
  s->N = N;
  s->C = C;
  std::vector<uint32_t> nptr(N + 1, 0), ncell(4 * (size_t)C);
  std::vector<uint8_t> nlocal(4 * (size_t)C);
  for (size_t k = 0; k < 4 * (size_t)C; ++k) nptr[cd[k] + 1]++;
  for (uint32_t i = 0; i < N; ++i) nptr[i + 1] += nptr[i];
  {
    std::vector<uint32_t> fill(nptr.begin(), nptr.end() - 1);
    for (uint32_t c = 0; c < C; ++c)
      for (int j = 0; j < 4; ++j)
        {
          const uint32_t pos = fill[cd[4 * (size_t)c + j]]++;
          ncell[pos] = c;
          nlocal[pos] = (uint8_t)j;
        }
  }
  // ELL rows: the dofs of the adjacent cells, sorted
  std::vector<std::vector<uint32_t>> rows(N);
  uint32_t MW = 1;
  for (uint32_t i = 0; i < N; ++i)
    {
      std::vector<uint32_t> &r = rows[i];
      r.push_back(i); // a dof no cell touches still gets a unit diagonal slot
      for (uint32_t a = nptr[i]; a < nptr[i + 1]; ++a)
        for (int j = 0; j < 4; ++j) r.push_back(cd[4 * (size_t)ncell[a] + j]);
      std::sort(r.begin(), r.end());
      r.erase(std::unique(r.begin(), r.end()), r.end());
      MW = std::max<uint32_t>(MW, (uint32_t)r.size());
    }
  if (MW > 255) WBEM_FAIL(ctx, -1, "a dof is shared by too many cells for the mass-matrix rows (%u columns)", MW);
  s->MW = MW;
  std::vector<uint32_t> mcol((size_t)N * MW, CON_NONE);
  for (uint32_t i = 0; i < N; ++i) std::copy(rows[i].begin(), rows[i].end(), mcol.begin() + (size_t)i * MW);
  std::vector<uint8_t> npos(4 * ncell.size());
  for (uint32_t i = 0; i < N; ++i)
    for (uint32_t a = nptr[i]; a < nptr[i + 1]; ++a)
      for (int j = 0; j < 4; ++j)
        {
          const uint32_t col = cd[4 * (size_t)ncell[a] + j];
          npos[4 * (size_t)a + j] =
            (uint8_t)(std::lower_bound(rows[i].begin(), rows[i].end(), col) - rows[i].begin());
        }
  std::vector<uint32_t> cells(cd, cd + 4 * (size_t)C);
  int rc;
  if ((rc = con_upload(ctx, &s->d_cells, cells))) return rc;
  if ((rc = con_upload(ctx, &s->d_dir, ctx->h_dir))) return rc;
  if ((rc = con_upload(ctx, &s->d_nc_ptr, nptr))) return rc;
  if ((rc = con_upload(ctx, &s->d_nc_cell, ncell))) return rc;
  if ((rc = con_upload(ctx, &s->d_nc_local, nlocal))) return rc;
  if ((rc = con_upload(ctx, &s->d_nc_pos, npos))) return rc;
  if ((rc = con_upload(ctx, &s->d_mcol, mcol))) return rc;
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_mval, sizeof(double) * (size_t)N * MW));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_diag, sizeof(double) * N));
  for (double **p : {&s->d_b, &s->d_x, &s->d_r, &s->d_p, &s->d_ap, &s->d_normals, &s->d_grads})
    CUDA_OK(ctx, cudaMalloc((void **)p, sizeof(double) * 3 * (size_t)N));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_phi, sizeof(double) * N));
  int n_sm = 0, per_sm = 0, coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mass_cg, CG_THREADS, 0);
  if (!coop || per_sm < 1) WBEM_FAIL(ctx, -2, "cooperative launch unavailable for the mass-matrix solver");
  s->grid = std::max(1, std::min<int>(n_sm, (int)((N + CG_THREADS - 1) / CG_THREADS)));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_part, sizeof(double) * 2 * 3 * (size_t)s->grid));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_barrier, sizeof(unsigned int)));
  CUDA_OK(ctx, cudaMalloc((void **)&s->d_iters, sizeof(int)));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  // dofs with a non-trivial double-node set, and (should a caller's sets not be symmetric) their members
  s->multi.clear();
  {
    std::vector<uint8_t> in(N, 0);
    for (uint32_t i = 0; i < N; ++i)
      if (ctx->h_dn_ptr[i + 1] - ctx->h_dn_ptr[i] >= 2)
        {
          in[i] = 1;
          for (uint32_t k = ctx->h_dn_ptr[i]; k < ctx->h_dn_ptr[i + 1]; ++k) in[ctx->h_dn_idx[k]] = 1;
        }
    for (uint32_t i = 0; i < N; ++i)
      if (in[i]) s->multi.push_back(i);
  }
  s->kind.assign(N, 0);
  s->master.assign(N, 0);
  s->inhom.assign(N, 0.0);
  s->ready = true;
  return 0;
}

static HangArgs con_hang_args(const ConState *s)
{
  HangArgs H;
  H.nh = s->nh;
  H.nm = s->nm;
  H.h_dof = s->d_h_dof;
  H.hm_ptr = s->d_hm_ptr;
  H.hm_col = s->d_hm_col;
  H.hm_w = s->d_hm_w;
  H.m_dof = s->d_m_dof;
  H.mh_ptr = s->d_mh_ptr;
  H.mh_h = s->d_mh_h;
  H.mh_w = s->d_mh_w;
  H.hflag = s->d_hflag;
  return H;
}

// the caller's hanging-node lines (wbem_set_hanging_constraints), flattened for the projections
static int con_upload_hanging(wbem_ctx *ctx)
{
  ConState *s = con_state(ctx);
  if (s->hang_uploaded) return 0;
  const uint32_t N = ctx->N;
  const uint32_t nh = (uint32_t)s->base_lines.size();
  s->nh = nh;
  s->nm = 0;
  if (nh)
    {
      std::vector<uint8_t> flag(N, 0);
      for (uint32_t h : s->base_lines) flag[h] = 1;
      for (uint32_t c : s->base_col)
        if (flag[c]) WBEM_FAIL(ctx, -1, "hanging-node lines refer to hanging nodes (more than one refinement level across an edge)");
      std::vector<std::vector<std::pair<uint32_t, double>>> tr(N);
      for (uint32_t k = 0; k < nh; ++k)
        for (uint32_t e = s->base_ptr[k]; e < s->base_ptr[k + 1]; ++e) tr[s->base_col[e]].emplace_back(s->base_lines[k], s->base_val[e]);
      std::vector<uint32_t> m_dof, mh_ptr(1, 0), mh_h;
      std::vector<double> mh_w;
      for (uint32_t m = 0; m < N; ++m)
        if (!tr[m].empty())
          {
            m_dof.push_back(m);
            for (auto &pr : tr[m])
              {
                mh_h.push_back(pr.first);
                mh_w.push_back(pr.second);
              }
            mh_ptr.push_back((uint32_t)mh_h.size());
          }
      s->nm = (uint32_t)m_dof.size();
      int rc;
      if ((rc = con_upload(ctx, &s->d_h_dof, s->base_lines))) return rc;
      if ((rc = con_upload(ctx, &s->d_hm_ptr, s->base_ptr))) return rc;
      if ((rc = con_upload(ctx, &s->d_hm_col, s->base_col))) return rc;
      if ((rc = con_upload(ctx, &s->d_hm_w, s->base_val))) return rc;
      if ((rc = con_upload(ctx, &s->d_m_dof, m_dof))) return rc;
      if ((rc = con_upload(ctx, &s->d_mh_ptr, mh_ptr))) return rc;
      if ((rc = con_upload(ctx, &s->d_mh_h, mh_h))) return rc;
      if ((rc = con_upload(ctx, &s->d_mh_w, mh_w))) return rc;
      if ((rc = con_upload(ctx, &s->d_hflag, flag))) return rc;
      CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
  s->hang_uploaded = true;
  s->mass_geom_version = s->normals_geom_version = ~0ull; // the condensed systems changed
  return 0;
}

// condense the right-hand sides in d_b (and, after a new mass matrix, the Jacobi diagonal)
static int con_condense(wbem_ctx *ctx, int with_diag)
{
  ConState *s = con_state(ctx);
  if (!s->nh) return 0;
  const HangArgs H = con_hang_args(s);
  k_con_fold<<<(s->nm + 127) / 128, 128, 0, ctx->stream>>>(s->N, H, s->d_b, s->d_diag, with_diag);
  k_con_zero_hanging<<<(s->nh + 127) / 128, 128, 0, ctx->stream>>>(s->N, H, s->d_b, s->d_diag, with_diag);
  ctx->launches += 2;
  CUDA_OK(ctx, cudaGetLastError());
  return 0;
}

static int con_mass(wbem_ctx *ctx)
{ // mass matrix + normal rhs for the current geometry (left in d_b)
  ConState *s = con_state(ctx);
  if (!ctx->have_geometry) WBEM_FAIL(ctx, -3, "normals / surface gradients need the geometry (wbem_set_geometry)");
  k_mass_rows<<<(s->N + 127) / 128, 128, 0, ctx->stream>>>((const GaussTable *)ctx->d_gauss, s->N, s->MW, ctx->d_xyz, s->d_cells, s->d_dir, s->d_nc_ptr,
                                                          s->d_nc_cell, s->d_nc_local, s->d_nc_pos, s->d_mval,
                                                          s->d_diag, s->d_b);
  ctx->launches++;
  CUDA_OK(ctx, cudaGetLastError());
  s->mass_geom_version = ctx->geom_version;
  return con_condense(ctx, 1);
}

static int con_solve3(wbem_ctx *ctx)
{ // M x = b for the three right-hand sides in d_b
  ConState *s = con_state(ctx);
  cudaStream_t st = ctx->stream;
  CUDA_OK(ctx, cudaMemsetAsync(s->d_barrier, 0, sizeof(unsigned int), st));
  uint32_t N = s->N, MW = s->MW;
  double rtol = CG_RTOL;
  int max_iters = CG_MAX_ITERS;
  HangArgs H = con_hang_args(s);
  void *args[] = {&N,       &MW,      &s->d_mcol, &s->d_mval, &s->d_diag,    &s->d_b,     &s->d_x,  &s->d_r,
                  &s->d_p,  &s->d_ap, &s->d_part, &s->d_barrier, &s->d_iters, &rtol,      &max_iters, &H};
  CUDA_OK(ctx, cudaLaunchCooperativeKernel((void *)k_mass_cg, dim3(s->grid), dim3(CG_THREADS), args, 0, st));
  ctx->launches++;
  return 0;
}

static int con_normals(wbem_ctx *ctx)
{
  int rc = con_prepare(ctx);
  if (rc) return rc;
  if ((rc = con_upload_hanging(ctx))) return rc;
  ConState *s = con_state(ctx);
  if (s->normals_geom_version == ctx->geom_version) return 0;
  if ((rc = con_mass(ctx))) return rc;
  if ((rc = con_solve3(ctx))) return rc;
  k_interleave3<<<(s->N + 255) / 256, 256, 0, ctx->stream>>>(s->N, s->d_x, s->d_normals, 1);
  ctx->launches++;
  s->h_normals.resize(3 * (size_t)s->N);
  CUDA_OK(ctx, cudaMemcpyAsync(s->h_normals.data(), s->d_normals, sizeof(double) * 3 * s->N, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaMemcpyAsync(&s->last_cg_iters, s->d_iters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (s->last_cg_iters >= CG_MAX_ITERS) WBEM_FAIL(ctx, -6, "mass-matrix CG (normals) did not converge");
  s->normals_geom_version = ctx->geom_version;
  return 0;
}

static int con_gradients(wbem_ctx *ctx, const double *d_tmp_rhs)
{
  int rc = con_prepare(ctx);
  if (rc) return rc;
  if ((rc = con_upload_hanging(ctx))) return rc;
  ConState *s = con_state(ctx);
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "surface gradients need the masks (wbem_set_masks)");
  cudaStream_t st = ctx->stream;
  if (s->mass_geom_version != ctx->geom_version)
    { // the mass matrix follows the geometry; the normal rhs it leaves in d_b is overwritten below
      if ((rc = con_mass(ctx))) return rc;
    }
  k_mask_product<<<(s->N + 255) / 256, 256, 0, st>>>(s->N, d_tmp_rhs, ctx->d_surf, s->d_phi);
  k_gradient_rhs<<<(s->N + 127) / 128, 128, 0, st>>>((const GaussTable *)ctx->d_gauss, s->N, ctx->d_xyz, s->d_cells, s->d_nc_ptr, s->d_nc_cell,
                                                    s->d_nc_local, s->d_phi, s->d_b);
  ctx->launches += 2;
  CUDA_OK(ctx, cudaGetLastError());
  if ((rc = con_condense(ctx, 0))) return rc;
  if ((rc = con_solve3(ctx))) return rc;
  k_interleave3<<<(s->N + 255) / 256, 256, 0, st>>>(s->N, s->d_x, s->d_grads, 0);
  ctx->launches++;
  s->h_grads.resize(3 * (size_t)s->N);
  CUDA_OK(ctx, cudaMemcpyAsync(s->h_grads.data(), s->d_grads, sizeof(double) * 3 * s->N, cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaMemcpyAsync(&s->last_cg_iters, s->d_iters, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_OK(ctx, cudaStreamSynchronize(st));
  if (s->last_cg_iters >= CG_MAX_ITERS) WBEM_FAIL(ctx, -6, "mass-matrix CG (surface gradients) did not converge");
  return 0;
}

// compute_constraints with tmp_rhs in device memory (what solve_system calls, :845)
int wbem_compute_constraints_device(wbem_ctx *ctx, const double *d_tmp_rhs)
{
  if (!ctx->have_masks) WBEM_FAIL(ctx, -3, "compute_constraints before wbem_set_masks");
  int rc = con_prepare(ctx);
  if (rc) return rc;
  ConState *s = con_state(ctx);
  const uint32_t N = ctx->N;
  const uint32_t *dn_ptr = ctx->h_dn_ptr.data(), *dn_idx = ctx->h_dn_idx.data();
  const double *surf = ctx->h_surf.data();
  // which inputs does the walk need?  normals: a set with two Dirichlet dofs; gradients: such a
  // pair with different normals; tmp_rhs values: a Dirichlet first with a Neumann double
  bool need_normals = false, need_rhs = false;
  for (uint32_t i : s->multi)
    {
      const uint32_t b = dn_ptr[i], e = dn_ptr[i + 1];
      int nd = 0;
      for (uint32_t k = b; k < e; ++k) nd += surf[dn_idx[k]] == 1;
      if (nd >= 2) need_normals = true;
      if (nd >= 1 && nd < (int)(e - b)) need_rhs = true;
      if (need_normals && need_rhs) break;
    }
  if (need_rhs)
    { // start the copy now: it overlaps the normals' kernels
      s->h_rhs.resize(N);
      CUDA_OK(ctx, cudaMemcpyAsync(s->h_rhs.data(), d_tmp_rhs, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
    }
  if (need_normals && (rc = con_normals(ctx))) return rc;
  if (need_rhs) CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  const double *h_rhs = s->h_rhs.data();
  const double *nrm = s->h_normals.data();
  auto ndist = [&](uint32_t a, uint32_t b) {
    double d2 = 0;
    for (int d = 0; d < 3; ++d) d2 += (nrm[3 * (size_t)a + d] - nrm[3 * (size_t)b + d]) * (nrm[3 * (size_t)a + d] - nrm[3 * (size_t)b + d]);
    return std::sqrt(d2);
  };
  bool need_grads = false;
  if (need_normals)
    for (uint32_t i : s->multi)
      {
        if (surf[i] != 1) continue;
        for (uint32_t k = dn_ptr[i]; k < dn_ptr[i + 1]; ++k)
          if (dn_idx[k] != i && surf[dn_idx[k]] == 1 && !(ndist(dn_idx[k], i) < 1e-4)) need_grads = true;
        if (need_grads) break;
      }
  if (need_grads && (rc = con_gradients(ctx, d_tmp_rhs))) return rc;
  const double *grd = s->h_grads.data();

  // the walk of :1002-1101 over the non-trivial sets (ascending dof order, like the reference loop)
  std::vector<uint8_t> &kind = s->kind;
  std::vector<uint32_t> &master = s->master;
  std::vector<double> &inhom = s->inhom;
  for (uint32_t i : s->multi) kind[i] = 0;
  for (uint32_t i : s->multi)
    {
      const uint32_t b = dn_ptr[i], e = dn_ptr[i + 1];
      uint32_t first = dn_idx[b];
      for (uint32_t k = b; k < e; ++k)
        if (surf[dn_idx[k]] == 1)
          {
            first = dn_idx[k];
            break;
          }
      if (i != first) continue;
      for (uint32_t k = b; k < e; ++k)
        {
          const uint32_t d = dn_idx[k];
          if (d == i) continue;
          if (surf[i] == 1)
            {
              if (surf[d] == 1)
                {
                  if (ndist(d, i) < 1e-4)
                    { // flat edge between Dirichlet faces: equal normal derivatives
                      kind[d] = 1;
                      master[d] = i;
                      inhom[d] = 0.0;
                    }
                  else
                    { // sharp edge: both normal derivatives follow from the surface gradients
                      double c = 0, gd_ni = 0, gi_nd = 0;
                      for (int q = 0; q < 3; ++q)
                        {
                          c += nrm[3 * (size_t)i + q] * nrm[3 * (size_t)d + q];
                          gd_ni += grd[3 * (size_t)d + q] * nrm[3 * (size_t)i + q];
                          gi_nd += grd[3 * (size_t)i + q] * nrm[3 * (size_t)d + q];
                        }
                      const double f = 1.0 / (1.0 - c * c);
                      kind[i] = 2;
                      inhom[i] = f * (gd_ni + gi_nd * c);
                      kind[d] = 2;
                      inhom[d] = f * (gi_nd + gd_ni * c);
                    }
                }
              else
                { // Neumann double of a Dirichlet node: its potential is the imposed one
                  kind[d] = 2;
                  inhom[d] = h_rhs[i];
                }
            }
          else
            { // all Neumann: the potentials of the doubles are equal
              kind[d] = 1;
              master[d] = i;
              inhom[d] = 0.0;
            }
        }
    }
  // merge with the caller's hanging-node lines, sorted by constrained dof (ConstraintMatrix::close):
  // two sorted lists -- the dofs of the non-trivial sets and the (sorted) hanging-node lines
  std::vector<uint32_t> base_order(s->base_lines.size());
  for (size_t k = 0; k < base_order.size(); ++k) base_order[k] = (uint32_t)k;
  std::sort(base_order.begin(), base_order.end(),
            [&](uint32_t x, uint32_t y) { return s->base_lines[x] < s->base_lines[y]; });
  s->out_lines.clear();
  s->out_ptr.assign(1, 0);
  s->out_col.clear();
  s->out_val.clear();
  s->out_inhom.clear();
  size_t im = 0, ib = 0;
  while (im < s->multi.size() || ib < base_order.size())
    {
      const uint32_t dm = im < s->multi.size() ? s->multi[im] : CON_NONE;
      const uint32_t db = ib < base_order.size() ? s->base_lines[base_order[ib]] : CON_NONE;
      const uint32_t i = std::min(dm, db);
      const bool has_base = db == i, has_set = dm == i && kind[i] != 0;
      if (has_base || has_set)
        {
          s->out_lines.push_back(i);
          if (has_base)
            for (uint32_t k = s->base_ptr[base_order[ib]]; k < s->base_ptr[base_order[ib] + 1]; ++k)
              {
                s->out_col.push_back(s->base_col[k]);
                s->out_val.push_back(s->base_val[k]);
              }
          if (has_set && kind[i] == 1)
            {
              s->out_col.push_back(master[i]);
              s->out_val.push_back(1.0);
            }
          s->out_ptr.push_back((uint32_t)s->out_col.size());
          s->out_inhom.push_back(has_set ? inhom[i] : 0.0);
        }
      if (dm == i) ++im;
      if (db == i) ++ib;
    }
  // ConstraintMatrix::close() (:1104): entries that refer to a constrained dof are replaced by that
  // dof's line until none is left (a flat double of a dof that a sharp edge made inhomogeneous; a hanging
  // node whose master is a double node).  Lines that need no substitution keep their entries as they are.
  {
    std::vector<int32_t> &line_of = s->line_of_tmp;
    line_of.assign(N, -1);
    const size_t nl = s->out_lines.size();
    for (size_t k = 0; k < nl; ++k) line_of[s->out_lines[k]] = (int32_t)k;
    bool any = false;
    for (size_t k = 0; k < s->out_col.size() && !any; ++k) any = line_of[s->out_col[k]] >= 0;
    for (int pass = 0; any && pass < 16; ++pass)
      {
        any = false;
        std::vector<uint32_t> nptr(1, 0), ncol;
        std::vector<double> nval, ninh(s->out_inhom);
        std::vector<std::pair<uint32_t, double>> acc;
        for (size_t k = 0; k < nl; ++k)
          {
            bool sub = false;
            for (uint32_t e = s->out_ptr[k]; e < s->out_ptr[k + 1]; ++e)
              sub |= line_of[s->out_col[e]] >= 0 && s->out_col[e] != s->out_lines[k];
            if (!sub)
              {
                ncol.insert(ncol.end(), s->out_col.begin() + s->out_ptr[k], s->out_col.begin() + s->out_ptr[k + 1]);
                nval.insert(nval.end(), s->out_val.begin() + s->out_ptr[k], s->out_val.begin() + s->out_ptr[k + 1]);
              }
            else
              {
                any = true;
                acc.clear();
                for (uint32_t e = s->out_ptr[k]; e < s->out_ptr[k + 1]; ++e)
                  {
                    const uint32_t c = s->out_col[e];
                    const double v = s->out_val[e];
                    const int32_t L = c != s->out_lines[k] ? line_of[c] : -1;
                    if (L < 0)
                      acc.emplace_back(c, v);
                    else
                      {
                        for (uint32_t e2 = s->out_ptr[L]; e2 < s->out_ptr[L + 1]; ++e2)
                          acc.emplace_back(s->out_col[e2], v * s->out_val[e2]);
                        ninh[k] += v * s->out_inhom[L];
                      }
                  }
                std::stable_sort(acc.begin(), acc.end(), [](const std::pair<uint32_t, double> &x, const std::pair<uint32_t, double> &y) { return x.first < y.first; });
                for (size_t a2 = 0; a2 < acc.size();)
                  {
                    double v = 0;
                    size_t b2 = a2;
                    for (; b2 < acc.size() && acc[b2].first == acc[a2].first; ++b2) v += acc[b2].second;
                    ncol.push_back(acc[a2].first);
                    nval.push_back(v);
                    a2 = b2;
                  }
              }
            nptr.push_back((uint32_t)ncol.size());
          }
        s->out_ptr.swap(nptr);
        s->out_col.swap(ncol);
        s->out_val.swap(nval);
        s->out_inhom.swap(ninh);
        if (pass == 15 && any) WBEM_FAIL(ctx, -1, "constraint lines refer to each other in a cycle");
      }
  }
  return wbem_set_constraints(ctx, (uint32_t)s->out_lines.size(), s->out_lines.data(), s->out_ptr.data(),
                              s->out_col.data(), s->out_val.data(), s->out_inhom.data());
}

extern "C" {

int wbem_compute_normals(wbem_ctx *ctx, double *normals)
{
  if (!ctx) return -1;
  GROUP_FORWARD(ctx, wbem_compute_normals(s, wbem_is_root(s) ? normals : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  int rc = con_normals(ctx);
  if (rc) return rc;
  if (normals) memcpy(normals, con_state(ctx)->h_normals.data(), sizeof(double) * 3 * (size_t)ctx->N);
  return 0;
}

int wbem_compute_surface_gradients(wbem_ctx *ctx, const double *tmp_rhs, double *gradients)
{
  if (!ctx || !tmp_rhs) return -1;
  GROUP_FORWARD(ctx, wbem_compute_surface_gradients(s, tmp_rhs, wbem_is_root(s) ? gradients : nullptr));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_compute_surface_gradients before wbem_set_topology");
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[1], tmp_rhs, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  int rc = con_gradients(ctx, ctx->d_tmp[1]);
  if (rc) return rc;
  if (gradients) memcpy(gradients, con_state(ctx)->h_grads.data(), sizeof(double) * 3 * (size_t)ctx->N);
  return 0;
}

int wbem_set_hanging_constraints(wbem_ctx *ctx, uint32_t n_lines, const uint32_t *lines, const uint32_t *ptr,
                                 const uint32_t *col, const double *val)
{
  if (!ctx) return -1;
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_set_hanging_constraints before wbem_set_topology");
  GROUP_FORWARD(ctx, wbem_set_hanging_constraints(s, n_lines, lines, ptr, col, val));
  ConState *s = con_state(ctx);
  for (uint32_t k = 0; k < n_lines; ++k)
    if (lines[k] >= ctx->N) WBEM_FAIL(ctx, -1, "hanging-node line out of range");
  const uint32_t nnz = n_lines ? ptr[n_lines] : 0;
  for (uint32_t k = 0; k < nnz; ++k)
    if (col[k] >= ctx->N) WBEM_FAIL(ctx, -1, "hanging-node entry out of range");
  s->base_lines.assign(lines, lines + n_lines);
  s->base_ptr.assign(1, 0);
  if (n_lines) s->base_ptr.assign(ptr, ptr + n_lines + 1);
  s->base_col.assign(col, col + nnz);
  s->base_val.assign(val, val + nnz);
  s->hang_uploaded = false; // the projections condense these lines into their mass systems
  return 0;
}

int wbem_compute_constraints(wbem_ctx *ctx, const double *tmp_rhs)
{
  if (!ctx || !tmp_rhs) return -1;
  GROUP_FORWARD(ctx, wbem_compute_constraints(s, tmp_rhs));
  CUDA_OK(ctx, cudaSetDevice(ctx->dev));
  if (!ctx->N) WBEM_FAIL(ctx, -3, "wbem_compute_constraints before wbem_set_topology");
  CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tmp[1], tmp_rhs, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
  return wbem_compute_constraints_device(ctx, ctx->d_tmp[1]);
}

int wbem_get_constraints(wbem_ctx *ctx, uint32_t *n_lines, uint32_t *nnz, uint32_t *lines, uint32_t *ptr,
                         uint32_t *col, double *val, double *inhom)
{
  if (!ctx) return -1;
  ConState *s = con_state(ctx);
  if (n_lines) *n_lines = (uint32_t)s->out_lines.size();
  if (nnz) *nnz = (uint32_t)s->out_col.size();
  if (lines) std::copy(s->out_lines.begin(), s->out_lines.end(), lines);
  if (ptr) std::copy(s->out_ptr.begin(), s->out_ptr.end(), ptr);
  if (col) std::copy(s->out_col.begin(), s->out_col.end(), col);
  if (val) std::copy(s->out_val.begin(), s->out_val.end(), val);
  if (inhom) std::copy(s->out_inhom.begin(), s->out_inhom.end(), inhom);
  return 0;
}

int wbem_mass_cg_iterations(wbem_ctx *ctx) { return ctx && ctx->con ? con_state(ctx)->last_cg_iters : -1; }

} // extern "C"
